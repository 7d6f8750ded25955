// tcgen05 / TMEM / TMA implicit-GEMM 3x3 convolution for sm_100a.
//
// Replaces torch.nn.Conv2d(k=3) -> cuDNN/TensorRT on the reference's hot path
// (src/upscale/model/realesrgan/factory.py:44-67 SRVGG convs, basicsr RRDBNet convs reached from
//  factory.py:113-125, src/upscale/model/bsvd/model.py:22-53,231-323 BSVD convs).
//
// Design:
//   * activations NHWC 16-bit in HBM; GEMM M = 128 consecutive pixels of one image row (TMEM lanes),
//     N = output channels (TMEM columns), K = 9 taps x input channels in 64-channel blocks.
//   * one persistent CTA per SM; a tile is R output rows x 128 pixels with R accumulators in TMEM,
//     double buffered (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * each input halo row (130 pixels x 64 channels = one swizzle-128B K-major slab) is loaded ONCE by
//     TMA (zero fill outside the image == conv padding) and consumed by up to 9 tap-MMAs: the 3
//     horizontal taps are the same slab addressed with a 0/1/2-row shifted start address, the 3
//     vertical taps feed 3 different accumulators.  No im2col, no 9x re-read of activations.
//   * weights for a whole 64-channel K block (all taps) stay in shared memory, double buffered
//     (or resident for the whole kernel when K fits one block).
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2..5 = epilogue
//     (TMEM -> registers -> bias / PReLU / ReLU6 / scaled residual adds -> 16-byte NHWC stores,
//      or PixelShuffle / NCHW / uint8 / temporal-shift scatter stores).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_params.h"

namespace ss4k {

namespace {

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a pipeline bug must not hang the GPU.  After ~4 s the kernel records where it
// was stuck and traps (the host then reports SS4K_E_CUDA with the diagnostic).
__device__ __noinline__ void mbar_timeout(int32_t* err, int tag, uint32_t parity) {
  if (err != nullptr) {
    err[0] = tag;
    err[1] = static_cast<int32_t>(blockIdx.x);
    err[2] = static_cast<int32_t>(threadIdx.x);
    err[3] = static_cast<int32_t>(parity);
    __threadfence_system();
  }
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int32_t* err, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (globaltimer_ns() - t0 > 4000000000ull) mbar_timeout(err, tag, parity);
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                            int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], M=128, K=16, 16-bit operands, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, swizzle-128B shared memory matrix descriptor (sm_100 "version 1"):
//   rows are 128 bytes (64 x 16-bit), 8-row groups are 1024 bytes apart (SBO), LBO unused.
//   base_offset = row phase of the start address inside the 1024-byte swizzle pattern.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t base_offset) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(1024u >> 4) << 32) |
         (1ull << 46) | (static_cast<uint64_t>(base_offset & 7u) << 49) | (2ull << 61);
}

// ---------------------------------------------------------------- 16-bit pack / unpack
__device__ __forceinline__ uint32_t pack2(float a, float b, bool bf16) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t u, bool bf16) {
  if (bf16) {
    __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(h);
  }
  __half2 h = *reinterpret_cast<__half2*>(&u);
  return __half22float2(h);
}
__device__ __forceinline__ float round16(float a, bool bf16) {
  return bf16 ? __bfloat162float(__float2bfloat16_rn(a)) : __half2float(__float2half_rn(a));
}

__device__ __forceinline__ void add_residual16(float (&v)[16], const void* res, size_t elem_off,
                                               float beta, bool bf16) {
  const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(res) + elem_off);
  // plain (coherent) loads: the RRDB tail conv updates its residual buffer in place
  const uint4 q0 = *p;
  const uint4 q1 = *(p + 1);
  const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 f = unpack2(w[i], bf16);
    v[2 * i] = fmaf(beta, f.x, v[2 * i]);
    v[2 * i + 1] = fmaf(beta, f.y, v[2 * i + 1]);
  }
}

__device__ __forceinline__ void store8(void* base, size_t elem_off, const float* v, bool bf16) {
  uint4 q;
  q.x = pack2(v[0], v[1], bf16);
  q.y = pack2(v[2], v[3], bf16);
  q.z = pack2(v[4], v[5], bf16);
  q.w = pack2(v[6], v[7], bf16);
  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(base) + elem_off) = q;
}
__device__ __forceinline__ void store8_lo(void* base, size_t elem_off, const float* v, bool bf16) {
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = v[i] - round16(v[i], bf16);
  store8(base, elem_off, r, bf16);
}

// One 16-channel chunk of one output pixel: bias, activation, residuals, store.
//   ch0: channel index of v[0] within the conv's (padded) output channels
//   (ay, ax): pixel in A space; sub: accumulator phase (kModeUp2)
__device__ __forceinline__ void epilogue_chunk(const ConvParams& P, int n, int ay, int ax, int sub,
                                               int ch0, float (&v)[16]) {
  const Epilogue& E = P.ep;
  const bool bf16 = E.is_bf16 != 0;
  // ---- bias + activation
  {
    const float4* bp = reinterpret_cast<const float4*>(E.bias + ch0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = __ldg(bp + q);
      v[4 * q + 0] += b.x;
      v[4 * q + 1] += b.y;
      v[4 * q + 2] += b.z;
      v[4 * q + 3] += b.w;
    }
    if (E.act == kActPRelu) {
      const float4* sp = reinterpret_cast<const float4*>(E.slope + ch0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 s = __ldg(sp + q);
        v[4 * q + 0] = v[4 * q + 0] >= 0.f ? v[4 * q + 0] : v[4 * q + 0] * s.x;
        v[4 * q + 1] = v[4 * q + 1] >= 0.f ? v[4 * q + 1] : v[4 * q + 1] * s.y;
        v[4 * q + 2] = v[4 * q + 2] >= 0.f ? v[4 * q + 2] : v[4 * q + 2] * s.z;
        v[4 * q + 3] = v[4 * q + 3] >= 0.f ? v[4 * q + 3] : v[4 * q + 3] * s.w;
      }
    } else if (E.act == kActRelu6) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fminf(fmaxf(v[i], 0.f), 6.f);
    }
    if (E.alpha != 1.0f) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] *= E.alpha;
    }
  }
  // ---- output pixel / channel
  int oy = ay, ox = ax, oc = ch0;
  if (P.mode == kModeUp2) {
    oy = 2 * ay + (sub >> 1);
    ox = 2 * ax + (sub & 1);
  }
  if (E.out_mode == kOutPS2NHWC) {
    const int cq = E.cout >> 2;  // channels of the shuffled output
    const int ab = ch0 / cq;
    oc = ch0 - ab * cq;
    oy = 2 * oy + (ab >> 1);
    ox = 2 * ox + (ab & 1);
  }
  const size_t pix = (static_cast<size_t>(n) * E.out_h + oy) * E.out_w + ox;
  // ---- residuals (indexed at the output pixel, NHWC)
  if (E.res1 != nullptr)
    add_residual16(v, E.res1, pix * E.res1_pitch + E.res1_coff + oc, E.beta1, bf16);
  if (E.res2 != nullptr)
    add_residual16(v, E.res2, pix * E.res2_pitch + E.res2_coff + oc, E.beta2, bf16);
  // ---- store
  switch (E.out_mode) {
    case kOutNHWC:
    case kOutPS2NHWC: {
      const size_t off = pix * E.out_pitch + E.out_coff + oc;
      store8(E.out, off, v, bf16);
      store8(E.out, off + 8, v + 8, bf16);
      if (E.out_lo != nullptr) {
        store8_lo(E.out_lo, off, v, bf16);
        store8_lo(E.out_lo, off + 8, v + 8, bf16);
      }
    } break;
    case kOutScatterNHWC: {
      const size_t off = pix * E.out_pitch + E.out_coff + oc;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = oc + 8 * h;
        void* dst = c < E.fold ? E.out2 : (c < 2 * E.fold ? E.out3 : E.out);
        store8(dst, off + 8 * h, v + 8 * h, bf16);
      }
    } break;
    case kOutNCHWF32: {
      float* o = reinterpret_cast<float*>(E.out);
      const size_t plane = static_cast<size_t>(E.out_h) * E.out_w;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = oc + i;
        if (c < E.cout) o[(static_cast<size_t>(n) * E.cout + c) * plane + static_cast<size_t>(oy) * E.out_w + ox] = v[i];
      }
    } break;
    case kOutNCHWF16: {
      __half* o = reinterpret_cast<__half*>(E.out);
      const size_t plane = static_cast<size_t>(E.out_h) * E.out_w;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = oc + i;
        if (c < E.cout)
          o[(static_cast<size_t>(n) * E.cout + c) * plane + static_cast<size_t>(oy) * E.out_w + ox] = __float2half_rn(v[i]);
      }
    } break;
    case kOutU8NHWC: {
      uint8_t* o = reinterpret_cast<uint8_t*>(E.out) + pix * E.cout;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = oc + i;
        if (c < E.cout) {
          float f = fminf(fmaxf(v[i], 0.f), 1.f) * 255.f;
          if (E.round_u8) f = rintf(f);
          o[c] = static_cast<uint8_t>(f);
        }
      }
    } break;
    case kOutPSNCHWF32: {
      // conv channel ch = c*r*r + a*r + b  ->  out[n, c, oy*r + a, ox*r + b]  (+ base[n, oy, ox, c])
      float* o = reinterpret_cast<float*>(E.out);
      const int r = E.ps_r, rr = r * r;
      const int oc_total = E.cout / rr;
      const int OH = E.out_h * r, OW = E.out_w * r;
      const uint16_t* bpix = E.base != nullptr
                                 ? reinterpret_cast<const uint16_t*>(E.base) + pix * E.base_pitch
                                 : nullptr;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int ch = oc + i;
        if (ch < E.cout) {
          const int c = ch / rr, rem = ch - c * rr;
          const int a = rem / r, b = rem - a * r;
          float val = v[i];
          if (bpix != nullptr) {
            const uint16_t raw = __ldg(bpix + c);
            val += bf16 ? __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&raw))
                        : __half2float(*reinterpret_cast<const __half*>(&raw));
          }
          o[((static_cast<size_t>(n) * oc_total + c) * OH + (static_cast<size_t>(oy) * r + a)) * OW + static_cast<size_t>(ox) * r + b] = val;
        }
      }
    } break;
    default:
      break;
  }
}

}  // namespace

// ---------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_tcgen05_kernel(const __grid_constant__ ConvParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = smem_base;
  const uint32_t a_base = w_base + static_cast<uint32_t>(P.w_slots) * P.w_slot_bytes;
  const uint32_t bar_base = a_base + static_cast<uint32_t>(P.a_slots) * P.a_slot_bytes;
  // barrier map (8 bytes each)
  const uint32_t a_full = bar_base;                   // [kMaxASlots]
  const uint32_t a_empty = bar_base + 8 * kMaxASlots; // [kMaxASlots]
  const uint32_t w_full = bar_base + 16 * kMaxASlots; // [2]
  const uint32_t w_empty = w_full + 16;               // [2]
  const uint32_t t_full = w_empty + 16;               // [2]
  const uint32_t t_empty = t_full + 16;               // [2]
  const uint32_t tmem_slot = t_empty + 16;            // uint32
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&P.tmA[0]);
    prefetch_tmap(&P.tmW);
    for (int i = 0; i < kMaxASlots; ++i) {
      mbar_init(a_full + 8 * i, 1);
      mbar_init(a_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(w_full + 8 * i, 1);
      mbar_init(w_empty + 8 * i, 1);
      mbar_init(t_full + 8 * i, 1);
      mbar_init(t_empty + 8 * i, 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(static_cast<uint32_t>(kTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int in_rows = P.R + P.max_dr;
  const int tiles_per_img = P.tiles_y * P.tiles_x * P.n_chunks;

  if (warp == 0) {
    // ======================================================= TMA producer
    if (lane == 0) {
      uint32_t as = 0, aph = 0, ws = 0, wph = 0;
      bool w_loaded = false;
      for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img;
        int rem = tile - n * tiles_per_img;
        const int ty = rem / (P.tiles_x * P.n_chunks);
        rem -= ty * (P.tiles_x * P.n_chunks);
        const int tx = rem / P.n_chunks;
        const int nc = rem - tx * P.n_chunks;
        const int y0 = ty * P.R, x0 = tx * kTileW;
        for (int kbi = 0; kbi < P.nkb; ++kbi) {
          if (!(P.w_resident && w_loaded)) {
            mbar_wait(w_empty + 8 * ws, wph ^ 1, P.err, 1);
            if (P.dbg_flags & 2) {
              mbar_arrive(w_full + 8 * ws);
            } else {
              mbar_expect_tx(w_full + 8 * ws, P.w_tx);
              tma_load_3d(w_base + ws * P.w_slot_bytes, &P.tmW, w_full + 8 * ws, 0, nc * P.n_cta,
                          kbi * P.ntaps);
            }
            w_loaded = true;
            if (++ws == static_cast<uint32_t>(P.w_slots)) { ws = 0; wph ^= 1; }
          }
          const KBlock kb = P.kb[kbi];
          const CUtensorMap* tm = &P.tmA[kb.tmap];
          for (int r = 0; r < in_rows; ++r) {
            mbar_wait(a_empty + 8 * as, aph ^ 1, P.err, 2);
            if (P.dbg_flags & 2) {
              mbar_arrive(a_full + 8 * as);
              if (++as == static_cast<uint32_t>(P.a_slots)) { as = 0; aph ^= 1; }
              continue;
            }
            mbar_expect_tx(a_full + 8 * as, P.a_row_tx);
            const uint32_t dst = a_base + as * P.a_slot_bytes;
            const int y = y0 - 1 + r;
            if (P.desc_mode == 2) {
#pragma unroll
              for (int s = 0; s < 3; ++s)
                tma_load_5d(dst + s * P.a_sub_bytes, tm, a_full + 8 * as, kb.c0, x0 - 1 + s, kb.p, y, n);
            } else {
              tma_load_5d(dst, tm, a_full + 8 * as, kb.c0, x0 - 1, kb.p, y, n);
            }
            if (++as == static_cast<uint32_t>(P.a_slots)) { as = 0; aph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================= MMA issuer
    if (lane == 0) {
      uint32_t as = 0, aph = 0, ws = 0, wph = 0;
      bool w_ready = false;
      uint32_t wcur = w_base;
      int it = 0;
      for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
        const uint32_t ts = it & 1, tph = (it >> 1) & 1;
        mbar_wait(t_empty + 8 * ts, tph ^ 1, P.err, 3);
        tcgen05_after_sync();
        const uint32_t tacc = tmem_base + ts * kAccStageCols;
        uint32_t started = 0;
        for (int kbi = 0; kbi < P.nkb; ++kbi) {
          uint32_t this_ws = ws;
          if (!(P.w_resident && w_ready)) {
            mbar_wait(w_full + 8 * ws, wph, P.err, 4);
            tcgen05_after_sync();
            wcur = w_base + ws * P.w_slot_bytes;
            w_ready = true;
            this_ws = ws;
            if (++ws == static_cast<uint32_t>(P.w_slots)) { ws = 0; wph ^= 1; }
          }
          for (int r = 0; r < in_rows; ++r) {
            mbar_wait(a_full + 8 * as, aph, P.err, 5);
            tcgen05_after_sync();
            const uint32_t arow = a_base + as * P.a_slot_bytes;
            for (int t = 0; t < P.ntaps && !(P.dbg_flags & 1); ++t) {
              const Tap tap = P.taps[t];
              const int j = r - tap.dr;
              if (j < 0 || j >= P.R) continue;
              const uint32_t mask = P.ksmask[kbi][t];
              if (mask == 0) continue;
              const uint32_t acc = static_cast<uint32_t>(j * P.nsub + tap.sub);
              const uint32_t d_tmem = tacc + acc * P.acc_stride;
              uint32_t a_addr, boff;
              if (P.desc_mode == 2) {
                a_addr = arow + tap.shift * P.a_sub_bytes;
                boff = 0;
              } else {
                a_addr = arow + tap.shift * kRowBytes;
                boff = P.desc_mode == 1 ? static_cast<uint32_t>(tap.shift) : 0u;
              }
              const uint32_t b_addr = wcur + t * P.w_tile_bytes;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                if (mask & (1u << ks)) {
                  umma_f16(d_tmem, make_sdesc(a_addr + ks * 32, boff), make_sdesc(b_addr + ks * 32, 0),
                           P.idesc, (started >> acc) & 1u);
                  started |= 1u << acc;
                }
              }
            }
            umma_commit(a_empty + 8 * as);  // slot reusable once these MMAs have read it
            if (++as == static_cast<uint32_t>(P.a_slots)) { as = 0; aph ^= 1; }
          }
          if (!P.w_resident) umma_commit(w_empty + 8 * this_ws);
        }
        umma_commit(t_full + 8 * ts);  // accumulators of this tile complete
      }
    }
  } else {
    // ======================================================= epilogue (warps 2..5)
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;       // accumulator row == pixel inside the tile row
    const int nacc = P.R * P.nsub;
    int it = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
      const int n = tile / tiles_per_img;
      int rem = tile - n * tiles_per_img;
      const int ty = rem / (P.tiles_x * P.n_chunks);
      rem -= ty * (P.tiles_x * P.n_chunks);
      const int tx = rem / P.n_chunks;
      const int nc = rem - tx * P.n_chunks;
      const int y0 = ty * P.R, x0 = tx * kTileW;
      const uint32_t ts = it & 1, tph = (it >> 1) & 1;
      mbar_wait(t_full + 8 * ts, tph, P.err, 6);
      tcgen05_after_sync();
      const uint32_t tacc = tmem_base + ts * kAccStageCols + (static_cast<uint32_t>(q * 32) << 16);
      const int ax = x0 + m;
      for (int ai = 0; ai < nacc; ++ai) {
        const int j = ai / P.nsub;
        const int sub = ai - j * P.nsub;
        const int ay = y0 + j;
        if (ay >= P.H) break;  // warp-uniform
        const bool valid = ax < P.W;
        for (int cb = 0; cb < P.n_cta; cb += 16) {
          uint32_t raw[16];
          tmem_ld16(tacc + ai * P.acc_stride + cb, raw);
          tmem_ld_wait();
          if (valid && !(P.dbg_flags & 4)) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]);
            epilogue_chunk(P, n, ay, ax, sub, nc * P.n_cta + cb, v);
          }
        }
      }
      tcgen05_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty + 8 * ts);
    }
  }

  // ---------------------------------------------------------------- teardown
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(kTmemCols))
                 : "memory");
  }
}

// ---------------------------------------------------------------- host launcher
cudaError_t conv_tc_prepare() {
  return cudaFuncSetAttribute(conv3x3_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              kSmemBytes);
}

cudaError_t conv_tc_launch(const ConvParams& p, int grid, cudaStream_t stream) {
  conv3x3_tcgen05_kernel<<<grid, kConvThreads, kSmemBytes, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace ss4k
