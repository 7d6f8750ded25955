// Row-streaming tcgen05 / TMEM / TMA implicit-GEMM 3x3 convolution (stride 1, pad 1) for sm_100a.
//
// Replaces torch.nn.Conv2d(k=3) -> cuDNN/TensorRT on the reference's hot path
// (src/upscale/model/realesrgan/factory.py:44-67 SRVGG convs, basicsr RRDBNet convs reached from
//  factory.py:113-125, src/upscale/model/bsvd/model.py:22-53,231-323 BSVD convs).
//
// Formulation:
//   * GEMM M = 128 consecutive pixels of one image row (TMEM lanes); K = input channels in 64-channel
//     blocks x 3 horizontal taps; N = 3 vertical taps x NOUT output channels: ONE MMA with N = 3*NOUT
//     adds input row r into the accumulators of output rows r-1, r, r+1, which sit side by side in a
//     ring of TMEM accumulator slots.  Versus one MMA per tap this reads the activation operand from
//     shared memory 3x less often (the bound for small Cout) and issues 3x fewer instructions.
//   * the three horizontal taps are the same 130-pixel halo slab addressed with a 0/1/2-row shifted
//     start address (swizzle-128B K-major descriptors), so no im2col and no re-read of activations.
//   * a CTA owns a contiguous run of output rows of one 128-pixel column strip and STREAMS input rows
//     through a shared-memory ring: every input row is loaded once (2 halo rows per band), the MMA
//     stream never drains between rows, and the epilogue warps retire output row y as soon as input
//     row y+1 has been accumulated.  Rows are split evenly over the persistent grid (one CTA per SM).
//   * all weights of the conv (every K block and tap) stay resident in shared memory.
//   * warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (TMEM -> registers ->
//     bias / PReLU / ReLU6 / scaled residual adds -> NHWC / PixelShuffle / NCHW / uint8 stores).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_common.cuh"
#include "conv_params.h"

namespace ss4k {

namespace {

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(pred));
  return pred;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

constexpr uint64_t kSdescHi = (static_cast<uint64_t>(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
__device__ __forceinline__ uint64_t sdesc(uint32_t saddr) {
  return kSdescHi | static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
}

struct Band {
  int chunk, n, strip, yb, ye;
};
// units are output rows in (chunk, n, strip, y) order; a band is a run of rows inside one strip
__device__ __forceinline__ bool next_band(const StreamParams& P, int& u, int u1, Band& b) {
  if (u >= u1) return false;
  int t = u;
  const int y = t % P.H;
  t /= P.H;
  b.strip = t % P.strips;
  t /= P.strips;
  b.n = t % P.n_img;
  b.chunk = t / P.n_img;
  b.yb = y;
  const int rem = u1 - u;
  b.ye = rem < P.H - y ? y + rem : P.H;
  u += b.ye - b.yb;
  return true;
}

struct MmaOp {
  uint32_t col;    // TMEM column of the first accumulator slot written
  uint32_t boff;   // byte offset of the first weight row block inside a (kb, kx) tile
  uint32_t idesc;
  uint32_t acc;    // accumulate flag of the row's very first MMA
};

}  // namespace

template <int NOUT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_stream_kernel(const __grid_constant__ StreamParams P) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t kWTile = 3u * NOUT * 128u;  // one (K block, horizontal tap) weight tile
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = smem_base;
  const uint32_t a_base = w_base + static_cast<uint32_t>(P.nkb) * 3u * kWTile;
  const uint32_t bar_base = a_base + static_cast<uint32_t>(P.a_slots) * kASlotBytes;
  const uint32_t a_full = bar_base;
  const uint32_t a_empty = a_full + 8 * kMaxSASlots;
  const uint32_t acc_full = a_empty + 8 * kMaxSASlots;
  const uint32_t acc_empty = acc_full + 8 * kMaxAccSlots;
  const uint32_t w_full = acc_empty + 8 * kMaxAccSlots;
  const uint32_t w_empty = w_full + 8;
  const uint32_t tmem_slot = w_empty + 8;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&P.tmA[0]);
    prefetch_tmap(&P.tmW);
    for (int i = 0; i < kMaxSASlots; ++i) {
      mbar_init(a_full + 8 * i, 1);
      mbar_init(a_empty + 8 * i, 1);
    }
    for (int i = 0; i < kMaxAccSlots; ++i) {
      mbar_init(acc_full + 8 * i, 1);
      mbar_init(acc_empty + 8 * i, 4);  // one arrive per epilogue warp
    }
    mbar_init(w_full, 1);
    mbar_init(w_empty, 1);
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

  const int u0 = static_cast<int>(static_cast<int64_t>(blockIdx.x) * P.total_units / gridDim.x);
  const int u1 = static_cast<int>(static_cast<int64_t>(blockIdx.x + 1) * P.total_units / gridDim.x);
  const int S = P.acc_slots;

  if (warp == 0) {
    // ======================================================= TMA producer
    uint32_t as = 0, aph = 0, wph = 0;
    int loaded_chunk = -1;
    int u = u0;
    Band b;
    while (next_band(P, u, u1, b)) {
      if (b.chunk != loaded_chunk) {
        mbar_wait(w_empty, wph ^ 1, P.err, 1);
        if (elect_one()) {
          const int ntile = P.nkb * 3;
          mbar_expect_tx(w_full, static_cast<uint32_t>(ntile) * kWTile);
          for (int t = 0; t < ntile; ++t)
            tma_load_2d(w_base + t * kWTile, &P.tmW, w_full, 0, (b.chunk * ntile + t) * 3 * NOUT);
        }
        __syncwarp();
        wph ^= 1;
        loaded_chunk = b.chunk;
      }
      const int r0 = b.yb > 0 ? b.yb - 1 : 0;
      const int r1 = b.ye < P.H ? b.ye : P.H - 1;
      const int x0 = b.strip * kTileW - 1;
      for (int r = r0; r <= r1; ++r) {
        for (int kb = 0; kb < P.nkb; ++kb) {
          mbar_wait(a_empty + 8 * as, aph ^ 1, P.err, 2);
          if (elect_one()) {
            if (P.dbg_flags & 2) {
              mbar_arrive(a_full + 8 * as);
            } else {
              mbar_expect_tx(a_full + 8 * as, kBoxW * kRowBytes);
              tma_load_5d(a_base + as * kASlotBytes, &P.tmA[P.a_tm[kb]], a_full + 8 * as, 0, x0, P.a_kb[kb], r, b.n);
            }
          }
          __syncwarp();
          if (++as == static_cast<uint32_t>(P.a_slots)) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================= MMA issuer
    uint32_t as = 0, aph = 0, wph = 0;
    int cur_chunk = -1;
    int qs = 0, qk = 0;  // accumulator ring position of the band's first output row: slot, wrap count
    int u = u0;
    Band b, nb;
    bool has = next_band(P, u, u1, b);
    while (has) {
      const bool has_next = next_band(P, u, u1, nb);
      if (b.chunk != cur_chunk) {
        mbar_wait(w_full, wph, P.err, 4);
        tcgen05_after_sync();
        wph ^= 1;
        cur_chunk = b.chunk;
      }
      const int r0 = b.yb > 0 ? b.yb - 1 : 0;
      const int r1 = b.ye < P.H ? b.ye : P.H - 1;
      for (int r = r0; r <= r1; ++r) {
        const int y_lo = r - 1 > b.yb ? r - 1 : b.yb;
        const int y_hi = r + 1 < b.ye - 1 ? r + 1 : b.ye - 1;
        const int b_lo = y_lo - (r - 1), b_hi = y_hi - (r - 1);  // weight row blocks (block = 2 - ky)
        // ---- accumulator slots touched for the first time by this input row must be drained
        {
          const int f_lo = (r == r0) ? y_lo : r + 1;
          for (int y = f_lo; y <= y_hi; ++y) {
            int s = qs + (y - b.yb), k = qk;
            while (s >= S) { s -= S; ++k; }
            mbar_wait(acc_empty + 8 * s, (k & 1) ^ 1, P.err, 3);
          }
          tcgen05_after_sync();
        }
        // ---- MMA op lists: `first` for the row's first MMA (zero-initialises fresh slots), `rest` after
        MmaOp first[3], rest[2];
        int nfirst = 0, nrest = 0;
        auto emit = [&](MmaOp* list, int& cnt, int b0, int b1, uint32_t acc) {
          int s0 = qs + (r - 1 + b0 - b.yb);
          while (s0 >= S) s0 -= S;
          const int nblk = b1 - b0 + 1;
          if (s0 + nblk <= S) {
            list[cnt++] = MmaOp{static_cast<uint32_t>(s0 * NOUT), static_cast<uint32_t>(b0 * NOUT * 128), P.idesc[nblk - 1], acc};
          } else {
            const int n1 = S - s0;
            list[cnt++] = MmaOp{static_cast<uint32_t>(s0 * NOUT), static_cast<uint32_t>(b0 * NOUT * 128), P.idesc[n1 - 1], acc};
            list[cnt++] = MmaOp{0u, static_cast<uint32_t>((b0 + n1) * NOUT * 128), P.idesc[nblk - n1 - 1], acc};
          }
        };
        emit(rest, nrest, b_lo, b_hi, 1u);
        if (r == r0) {
          emit(first, nfirst, b_lo, b_hi, 0u);
        } else if (b_hi == 2) {
          if (b_lo <= 1) emit(first, nfirst, b_lo, 1, 1u);
          emit(first, nfirst, 2, 2, 0u);
        } else {
          emit(first, nfirst, b_lo, b_hi, 1u);
        }
        // ---- K loop: K blocks x horizontal taps x 16-channel steps
        for (int kb = 0; kb < P.nkb; ++kb) {
          mbar_wait(a_full + 8 * as, aph, P.err, 5);
          tcgen05_after_sync();
          const uint32_t arow = a_base + as * kASlotBytes;
          const int nks = P.nks[kb];
          if (!(P.dbg_flags & 1)) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint32_t wt = w_base + static_cast<uint32_t>(kb * 3 + kx) * kWTile;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                if (ks < nks) {
                  const uint64_t da = sdesc(arow + kx * kRowBytes + ks * 32);
                  const uint32_t wb = wt + ks * 32;
                  if (kb == 0 && kx == 0 && ks == 0) {
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                      if (i < nfirst) umma_f16_elect(tmem_base + first[i].col, da, sdesc(wb + first[i].boff), first[i].idesc, first[i].acc);
                  } else {
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                      if (i < nrest) umma_f16_elect(tmem_base + rest[i].col, da, sdesc(wb + rest[i].boff), rest[i].idesc, 1u);
                  }
                }
              }
            }
          }
          umma_commit_elect(a_empty + 8 * as);  // slab reusable once these MMAs have read it
          if (++as == static_cast<uint32_t>(P.a_slots)) { as = 0; aph ^= 1; }
        }
        // ---- output rows completed by this input row
        if (r - 1 >= b.yb) {
          int s = qs + (r - 1 - b.yb);
          while (s >= S) s -= S;
          umma_commit_elect(acc_full + 8 * s);
        }
        if (r == r1 && r <= b.ye - 1) {  // image bottom: row H-1 has no row below it
          int s = qs + (r - b.yb);
          while (s >= S) s -= S;
          umma_commit_elect(acc_full + 8 * s);
        }
      }
      qs += b.ye - b.yb;
      while (qs >= S) { qs -= S; ++qk; }
      if (has_next && nb.chunk != b.chunk) umma_commit_elect(w_empty);  // weights may be replaced
      b = nb;
      has = has_next;
    }
  } else {
    // ======================================================= epilogue (warps 2..5)
    const int qd = warp & 3;       // TMEM lane quarter this warp may access
    const int m = qd * 32 + lane;  // accumulator row == pixel inside the strip
    int s = 0, k = 0;
    int u = u0;
    Band b;
    while (next_band(P, u, u1, b)) {
      const int ax = b.strip * kTileW + m;
      const bool valid = ax < P.W;
      for (int y = b.yb; y < b.ye; ++y) {
        mbar_wait(acc_full + 8 * s, k & 1, P.err, 6);
        tcgen05_after_sync();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + static_cast<uint32_t>(s * NOUT);
        uint32_t raw[NOUT];
#pragma unroll
        for (int c = 0; c < NOUT; c += 16) tmem_ld16p(taddr + c, &raw[c]);
        tmem_ld_wait();
        tcgen05_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + 8 * s);  // slot free: the MMA stream may reuse it
        if (valid && !(P.dbg_flags & 4)) {
#pragma unroll
          for (int c = 0; c < NOUT; c += 16) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[c + i]);
            epilogue_chunk(P.ep, kModeConv3, b.n, y, ax, 0, b.chunk * NOUT + c, v);
          }
        }
        if (++s == S) { s = 0; ++k; }
      }
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
cudaError_t conv_stream_prepare() {
  cudaError_t e = cudaFuncSetAttribute(conv3x3_stream_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_stream_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_stream_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_stream_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  return e;
}

cudaError_t conv_stream_launch(const StreamParams& p, int nout, int grid, cudaStream_t stream) {
  switch (nout) {
    case 16: conv3x3_stream_kernel<16><<<grid, kConvThreads, kSmemBytes, stream>>>(p); break;
    case 32: conv3x3_stream_kernel<32><<<grid, kConvThreads, kSmemBytes, stream>>>(p); break;
    case 48: conv3x3_stream_kernel<48><<<grid, kConvThreads, kSmemBytes, stream>>>(p); break;
    case 64: conv3x3_stream_kernel<64><<<grid, kConvThreads, kSmemBytes, stream>>>(p); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace ss4k
