// HBM-bound layout / colour kernels around the convolution stack, plus the naive direct
// convolution used ONLY by the start-up self-probe (ss4k_create) to validate the tcgen05 path.
//
//   prep_*   : frame at the boundary (float NCHW / half NCHW / uint8 NHWC / NV12) -> 16-bit NHWC,
//              channel-padded, optional pixel_unshuffle(2) (RRDBNet x2 head) and the [0,1]
//              normalisation  (reference: fsrcnn_upscaler.py:170-176,237-241; basicsr pixel_unshuffle)
//   unprep_* : 16-bit NHWC -> float NCHW (operator-level tests)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <stdint.h>

#include "elementwise.h"

namespace ss4k {

namespace {

__device__ __forceinline__ uint16_t to16(float f, bool bf16) {
  if (bf16) {
    __nv_bfloat16 h = __float2bfloat16_rn(f);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half_rn(f);
  return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ float from16(uint16_t u, bool bf16) {
  if (bf16) return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&u));
  return __half2float(*reinterpret_cast<__half*>(&u));
}

// BT.709 limited-range YCbCr -> RGB in [0,1] (the reference moves rgb24 through ffmpeg pipes,
// twitchgrabber.py:97-98, so the matrix is this repo's own definition)
__device__ __forceinline__ float3 yuv709_to_rgb(float y, float u, float v) {
  const float yy = (y - 16.f) * (1.f / 219.f);
  const float cb = (u - 128.f) * (1.f / 224.f);
  const float cr = (v - 128.f) * (1.f / 224.f);
  float3 rgb;
  rgb.x = yy + 1.5748f * cr;
  rgb.y = yy - 0.187324f * cb - 0.468124f * cr;
  rgb.z = yy + 1.8556f * cb;
  rgb.x = fminf(fmaxf(rgb.x, 0.f), 1.f);
  rgb.y = fminf(fmaxf(rgb.y, 0.f), 1.f);
  rgb.z = fminf(fmaxf(rgb.z, 0.f), 1.f);
  return rgb;
}

// value of source channel c at source pixel (n, y, x); channel >= C reads `fill`
struct SrcF32NCHW {
  const float* p; int C, H, W;
  __device__ float operator()(int n, int c, int y, int x) const {
    return __ldg(p + ((static_cast<size_t>(n) * C + c) * H + y) * W + x);
  }
};
struct SrcF16NCHW {
  const __half* p; int C, H, W;
  __device__ float operator()(int n, int c, int y, int x) const {
    return __half2float(p[((static_cast<size_t>(n) * C + c) * H + y) * W + x]);
  }
};
struct SrcU8NHWC {
  const uint8_t* p; int C, H, W;
  __device__ float operator()(int n, int c, int y, int x) const {
    return static_cast<float>(__ldg(p + ((static_cast<size_t>(n) * H + y) * W + x) * C + c)) / 255.0f;
  }
};
struct SrcNV12 {
  const uint8_t* p; int C, H, W;  // C == 3
  __device__ float operator()(int n, int c, int y, int x) const {
    const uint8_t* frame = p + static_cast<size_t>(n) * (static_cast<size_t>(H) * W * 3 / 2);
    const float Y = static_cast<float>(__ldg(frame + static_cast<size_t>(y) * W + x));
    const uint8_t* uv = frame + static_cast<size_t>(H) * W + static_cast<size_t>(y >> 1) * W + (x & ~1);
    const float3 rgb = yuv709_to_rgb(Y, static_cast<float>(__ldg(uv)), static_cast<float>(__ldg(uv + 1)));
    return c == 0 ? rgb.x : (c == 1 ? rgb.y : rgb.z);
  }
};

// one thread per OUTPUT pixel; out is [N, H/us, W/us, pitch] 16-bit; channels [0, C*us*us) real,
// channel `fill_ch` (if >= 0) = fill_val (BSVD noise map), the rest zero.
template <class Src>
__global__ void prep_kernel(Src src, uint16_t* __restrict__ out, uint16_t* __restrict__ out_lo, int N,
                            int pitch, int us, int fill_ch, float fill_val, int bf16) {
  const int OH = src.H / us, OW = src.W / us;
  const size_t total = static_cast<size_t>(N) * OH * OW;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = static_cast<int>(idx % OW);
  const int oy = static_cast<int>((idx / OW) % OH);
  const int n = static_cast<int>(idx / (static_cast<size_t>(OW) * OH));
  const int creal = src.C * us * us;
  uint16_t* o = out + idx * pitch;
  uint16_t* ol = out_lo != nullptr ? out_lo + idx * pitch : nullptr;
  for (int c8 = 0; c8 < pitch; c8 += 8) {
    uint16_t v[8], vl[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ch = c8 + i;
      float f = 0.f;
      if (ch < creal) {
        // torch pixel_unshuffle: ch = c*us*us + i_row*us + j_col
        const int c = ch / (us * us);
        const int rem = ch - c * us * us;
        const int iy = rem / us, jx = rem - iy * us;
        f = src(n, c, oy * us + iy, ox * us + jx);
      } else if (ch == fill_ch) {
        f = fill_val;
      }
      v[i] = to16(f, bf16 != 0);
      vl[i] = to16(f - from16(v[i], bf16 != 0), bf16 != 0);
    }
    *reinterpret_cast<uint4*>(o + c8) = *reinterpret_cast<uint4*>(v);
    if (ol != nullptr) *reinterpret_cast<uint4*>(ol + c8) = *reinterpret_cast<uint4*>(vl);
  }
}

// Hot case of the service boundary (fsrcnn_upscaler.py:170-172 + RRDBNet x2's pixel_unshuffle(2)): uint8 NHWC RGB frame
// -> /255 -> pixel_unshuffle(2) -> 16-channel-pitch 16-bit NHWC.  One thread per TWO trunk pixels: 2 rows x 12 source
// bytes (three aligned 32-bit loads per row), 64 bytes out (four 16-byte stores).  W % 4 == 0.
__global__ void prep_u8_unshuffle2_kernel(const uint8_t* __restrict__ in, uint16_t* __restrict__ out, int N, int H, int W, int bf16) {
  const int OH = H >> 1, OW2 = W >> 2;
  const size_t total = static_cast<size_t>(N) * OH * OW2;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int px = static_cast<int>(idx % OW2);
  const int oy = static_cast<int>((idx / OW2) % OH);
  const int n = static_cast<int>(idx / (static_cast<size_t>(OW2) * OH));
  uint8_t b[2][12];
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(in + (static_cast<size_t>(n) * H + 2 * oy + dy) * W * 3 + static_cast<size_t>(px) * 12);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint32_t w = __ldg(p + i);
      b[dy][4 * i] = static_cast<uint8_t>(w); b[dy][4 * i + 1] = static_cast<uint8_t>(w >> 8);
      b[dy][4 * i + 2] = static_cast<uint8_t>(w >> 16); b[dy][4 * i + 3] = static_cast<uint8_t>(w >> 24);
    }
  }
  uint16_t* o = out + ((static_cast<size_t>(n) * OH + oy) * (W >> 1) + 2 * px) * 16;
#pragma unroll
  for (int q = 0; q < 2; ++q) {  // trunk pixel 2*px + q: source pixels (2*oy + i, 4*px + 2*q + j)
    uint16_t v[16];
#pragma unroll
    for (int ch = 0; ch < 12; ++ch) {  // torch pixel_unshuffle: ch = c*4 + i*2 + j
      const int c = ch >> 2, i = (ch >> 1) & 1, j = ch & 1;
      v[ch] = to16(static_cast<float>(b[i][(2 * q + j) * 3 + c]) / 255.0f, bf16 != 0);
    }
    v[12] = v[13] = v[14] = v[15] = 0;
    *reinterpret_cast<uint4*>(o + 16 * q) = *reinterpret_cast<uint4*>(v);
    *reinterpret_cast<uint4*>(o + 16 * q + 8) = *reinterpret_cast<uint4*>(v + 8);
  }
}

__global__ void unprep_kernel(const uint16_t* __restrict__ in, const uint16_t* __restrict__ in_lo,
                              float* __restrict__ out, int N, int C, int H, int W, int pitch, int coff,
                              int bf16) {
  const size_t total = static_cast<size_t>(N) * C * H * W;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % W);
  const int y = static_cast<int>((idx / W) % H);
  const int c = static_cast<int>((idx / (static_cast<size_t>(W) * H)) % C);
  const int n = static_cast<int>(idx / (static_cast<size_t>(W) * H * C));
  const size_t off = ((static_cast<size_t>(n) * H + y) * W + x) * pitch + coff + c;
  float f = from16(in[off], bf16 != 0);
  if (in_lo != nullptr) f += from16(in_lo[off], bf16 != 0);
  out[idx] = f;
}

// naive direct conv (self-probe only): NHWC 16-bit in, packed-order-agnostic OIHW float weights that
// were already rounded to the operand type by the host; fp32 accumulate; float NCHW out.
__global__ void ref_conv3x3_kernel(const uint16_t* __restrict__ in, const float* __restrict__ w,
                                   const float* __restrict__ bias, float* __restrict__ out, int N, int H,
                                   int W, int pitch, int cin, int cout, int bf16) {
  const size_t total = static_cast<size_t>(N) * cout * H * W;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % W);
  const int y = static_cast<int>((idx / W) % H);
  const int co = static_cast<int>((idx / (static_cast<size_t>(W) * H)) % cout);
  const int n = static_cast<int>(idx / (static_cast<size_t>(W) * H * cout));
  float acc = bias != nullptr ? bias[co] : 0.f;
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = y + ky - 1;
    if (iy < 0 || iy >= H) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = x + kx - 1;
      if (ix < 0 || ix >= W) continue;
      const uint16_t* px = in + ((static_cast<size_t>(n) * H + iy) * W + ix) * pitch;
      for (int ci = 0; ci < cin; ++ci)
        acc = fmaf(from16(px[ci], bf16 != 0), w[((static_cast<size_t>(co) * cin + ci) * 3 + ky) * 3 + kx], acc);
    }
  }
  out[idx] = acc;
}

// uint8 NHWC RGB -> NV12 (encoder side of the north star's colour stage; no reference implementation: the
// reference hands rgb24 to ffmpeg, output_stream.py:127).  BT.709 limited range in 15-bit fixed point, chroma =
// mean of the 2x2 block; integer arithmetic so that the oracle (oracle/colour.py) matches bit for bit.
// One thread per 2x2 block pair (2 rows x 4 pixels): 24 bytes in, 8 Y + 4 UV bytes out.
__global__ void rgb_to_nv12_kernel(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ nv12, int N, int H, int W) {
  const int W4 = W >> 2, H2 = H >> 1;
  const size_t total = static_cast<size_t>(N) * H2 * W4;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int bx = static_cast<int>(idx % W4);
  const int by = static_cast<int>((idx / W4) % H2);
  const int n = static_cast<int>(idx / (static_cast<size_t>(W4) * H2));
  const uint8_t* src = rgb + (static_cast<size_t>(n) * H + 2 * by) * W * 3 + static_cast<size_t>(bx) * 12;
  uint8_t* frame = nv12 + static_cast<size_t>(n) * (static_cast<size_t>(H) * W * 3 / 2);
  int sr[2] = {0, 0}, sg[2] = {0, 0}, sb[2] = {0, 0};
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(src + static_cast<size_t>(dy) * W * 3);  // 12-byte aligned: W % 4 == 0
    const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
    const uint8_t px[12] = {static_cast<uint8_t>(w0), static_cast<uint8_t>(w0 >> 8), static_cast<uint8_t>(w0 >> 16), static_cast<uint8_t>(w0 >> 24),
                            static_cast<uint8_t>(w1), static_cast<uint8_t>(w1 >> 8), static_cast<uint8_t>(w1 >> 16), static_cast<uint8_t>(w1 >> 24),
                            static_cast<uint8_t>(w2), static_cast<uint8_t>(w2 >> 8), static_cast<uint8_t>(w2 >> 16), static_cast<uint8_t>(w2 >> 24)};
    uint32_t ypack = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = px[3 * i], g = px[3 * i + 1], b = px[3 * i + 2];
      const int y = 16 + ((5983 * r + 20127 * g + 2032 * b + 16384) >> 15);
      ypack |= static_cast<uint32_t>(y) << (8 * i);
      sr[i >> 1] += r; sg[i >> 1] += g; sb[i >> 1] += b;
    }
    *reinterpret_cast<uint32_t*>(frame + (static_cast<size_t>(2 * by + dy)) * W + 4 * bx) = ypack;
  }
  uint32_t uvpack = 0;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int u = 128 + ((-3298 * sr[j] - 11094 * sg[j] + 14392 * sb[j] + 65536) >> 17);
    const int v = 128 + ((14392 * sr[j] - 13073 * sg[j] - 1319 * sb[j] + 65536) >> 17);
    uvpack |= (static_cast<uint32_t>(u) | (static_cast<uint32_t>(v) << 8)) << (16 * j);
  }
  *reinterpret_cast<uint32_t*>(frame + static_cast<size_t>(H) * W + static_cast<size_t>(by) * W + 4 * bx) = uvpack;
}

// Same conversion, 2 rows x 16 pixels per thread with 16-byte loads / stores (W % 16 == 0).
__global__ void rgb_to_nv12_wide_kernel(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ nv12, int N, int H, int W) {
  const int W16 = W >> 4, H2 = H >> 1;
  const size_t total = static_cast<size_t>(N) * H2 * W16;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int bx = static_cast<int>(idx % W16);
  const int by = static_cast<int>((idx / W16) % H2);
  const int n = static_cast<int>(idx / (static_cast<size_t>(W16) * H2));
  uint8_t* frame = nv12 + static_cast<size_t>(n) * (static_cast<size_t>(H) * W * 3 / 2);
  int sr[8], sg[8], sb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sr[j] = sg[j] = sb[j] = 0;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
    const uint4* p = reinterpret_cast<const uint4*>(rgb + (static_cast<size_t>(n) * H + 2 * by + dy) * W * 3 + static_cast<size_t>(bx) * 48);
    uint4 q[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
    const uint8_t* px = reinterpret_cast<const uint8_t*>(q);
    uint32_t yw[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = px[3 * i], g = px[3 * i + 1], b = px[3 * i + 2];
      const int y = 16 + ((5983 * r + 20127 * g + 2032 * b + 16384) >> 15);
      yw[i >> 2] |= static_cast<uint32_t>(y) << (8 * (i & 3));
      sr[i >> 1] += r; sg[i >> 1] += g; sb[i >> 1] += b;
    }
    *reinterpret_cast<uint4*>(frame + static_cast<size_t>(2 * by + dy) * W + 16 * bx) = make_uint4(yw[0], yw[1], yw[2], yw[3]);
  }
  uint32_t uvw[4] = {0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int u = 128 + ((-3298 * sr[j] - 11094 * sg[j] + 14392 * sb[j] + 65536) >> 17);
    const int v = 128 + ((14392 * sr[j] - 13073 * sg[j] - 1319 * sb[j] + 65536) >> 17);
    uvw[j >> 1] |= (static_cast<uint32_t>(u) | (static_cast<uint32_t>(v) << 8)) << (16 * (j & 1));
  }
  *reinterpret_cast<uint4*>(frame + static_cast<size_t>(H) * W + static_cast<size_t>(by) * W + 16 * bx) = make_uint4(uvw[0], uvw[1], uvw[2], uvw[3]);
}

template <class Src>
cudaError_t launch_prep(Src src, void* out, void* out_lo, int N, int pitch, int us, int fill_ch,
                        float fill_val, int bf16, cudaStream_t s) {
  const size_t total = static_cast<size_t>(N) * (src.H / us) * (src.W / us);
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
  prep_kernel<Src><<<blocks, threads, 0, s>>>(src, reinterpret_cast<uint16_t*>(out),
                                              reinterpret_cast<uint16_t*>(out_lo), N, pitch, us, fill_ch,
                                              fill_val, bf16);
  return cudaGetLastError();
}

}  // namespace

cudaError_t prep_launch(int in_fmt, const void* in, void* out, void* out_lo, int N, int C, int H, int W,
                        int pitch, int unshuffle, int fill_ch, float fill_val, int bf16, cudaStream_t s) {
  const int us = unshuffle > 1 ? unshuffle : 1;
  switch (in_fmt) {
    case 0:
      return launch_prep(SrcF32NCHW{reinterpret_cast<const float*>(in), C, H, W}, out, out_lo, N, pitch, us, fill_ch, fill_val, bf16, s);
    case 1:
      return launch_prep(SrcF16NCHW{reinterpret_cast<const __half*>(in), C, H, W}, out, out_lo, N, pitch, us, fill_ch, fill_val, bf16, s);
    case 2:
      if (us == 2 && C == 3 && pitch == 16 && out_lo == nullptr && fill_ch < 0 && W % 4 == 0 && H % 2 == 0) {
        const size_t total = static_cast<size_t>(N) * (H / 2) * (W / 4);
        prep_u8_unshuffle2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(
            reinterpret_cast<const uint8_t*>(in), reinterpret_cast<uint16_t*>(out), N, H, W, bf16);
        return cudaGetLastError();
      }
      return launch_prep(SrcU8NHWC{reinterpret_cast<const uint8_t*>(in), C, H, W}, out, out_lo, N, pitch, us, fill_ch, fill_val, bf16, s);
    case 3:
      return launch_prep(SrcNV12{reinterpret_cast<const uint8_t*>(in), 3, H, W}, out, out_lo, N, pitch, us, fill_ch, fill_val, bf16, s);
    default:
      return cudaErrorInvalidValue;
  }
}

cudaError_t rgb_to_nv12_launch(const void* rgb, void* nv12, int N, int H, int W, cudaStream_t s) {
  if (H % 2 || W % 4) return cudaErrorInvalidValue;
  if (W % 16 == 0 && (reinterpret_cast<uintptr_t>(rgb) & 15) == 0 && (reinterpret_cast<uintptr_t>(nv12) & 15) == 0) {
    const size_t tw = static_cast<size_t>(N) * (H / 2) * (W / 16);
    rgb_to_nv12_wide_kernel<<<static_cast<unsigned>((tw + 127) / 128), 128, 0, s>>>(reinterpret_cast<const uint8_t*>(rgb),
                                                                                   reinterpret_cast<uint8_t*>(nv12), N, H, W);
    return cudaGetLastError();
  }
  const size_t total = static_cast<size_t>(N) * (H / 2) * (W / 4);
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
  rgb_to_nv12_kernel<<<blocks, threads, 0, s>>>(reinterpret_cast<const uint8_t*>(rgb), reinterpret_cast<uint8_t*>(nv12), N, H, W);
  return cudaGetLastError();
}

cudaError_t unprep_launch(const void* in, const void* in_lo, float* out, int N, int C, int H, int W,
                          int pitch, int coff, int bf16, cudaStream_t s) {
  const size_t total = static_cast<size_t>(N) * C * H * W;
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
  unprep_kernel<<<blocks, threads, 0, s>>>(reinterpret_cast<const uint16_t*>(in),
                                           reinterpret_cast<const uint16_t*>(in_lo), out, N, C, H, W, pitch,
                                           coff, bf16);
  return cudaGetLastError();
}

cudaError_t ref_conv3x3_launch(const void* in, const float* w, const float* bias, float* out, int N, int H,
                               int W, int pitch, int cin, int cout, int bf16, cudaStream_t s) {
  const size_t total = static_cast<size_t>(N) * cout * H * W;
  const int threads = 128;
  const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
  ref_conv3x3_kernel<<<blocks, threads, 0, s>>>(reinterpret_cast<const uint16_t*>(in), w, bias, out, N, H, W,
                                                pitch, cin, cout, bf16);
  return cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------
// Tiled inference (RealESRGANer.tile_process / pre_process semantics, reached from realesrgan/factory.py:93-95,160-169):
// the padded crops of one shape class are gathered from the frame into a batch, run through the net as independent
// images and the un-padded centres pasted into the output frame.  Rows / columns beyond the frame are the reflect
// padding of RealESRGANer.pre_process (pre_pad, and the mod-2 pad of the x2 nets).
namespace {

__device__ __forceinline__ int reflect_hi(int i, int n) { return i < n ? i : 2 * (n - 1) - i; }
// RealESRGANer.pre_process pads twice (F.pad reflect by pre_pad, then the mod-2 pad of the result): n1 = n + pre_pad
__device__ __forceinline__ int reflect2(int i, int n, int n1) { return reflect_hi(reflect_hi(i, n1), n); }

// elements of ES bytes; NCHW planes (nhwc == 0) or NHWC pixels of C elements (nhwc == 1)
template <int ES>
__global__ void tile_gather_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const TileBox* __restrict__ boxes,
                                   int nbox, int nimg, int N, int C, int H, int W, int H1, int W1, int hc, int wc, int nhwc) {
  const size_t per = static_cast<size_t>(C) * hc * wc;
  const size_t total = static_cast<size_t>(nimg) * N * per;
  for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t img = idx / per;           // atlas image index a * N + n
    size_t r = idx - img * per;
    const int a = static_cast<int>(img / N), n = static_cast<int>(img - static_cast<size_t>(a) * N);
    int c, y, x;
    if (nhwc) { c = static_cast<int>(r % C); r /= C; x = static_cast<int>(r % wc); y = static_cast<int>(r / wc); }
    else { x = static_cast<int>(r % wc); r /= wc; y = static_cast<int>(r % hc); c = static_cast<int>(r / hc); }
    int k = -1;                             // the crop of this atlas image that covers (y, x), if any
    for (int j = 0; j < nbox; ++j)
      if (boxes[j].img == a && x >= boxes[j].atlas_x && x < boxes[j].atlas_x + boxes[j].crop_w && y < boxes[j].crop_h) k = j;
    if (k < 0) {                            // gap column / canvas below a shorter crop
#pragma unroll
      for (int b = 0; b < ES; ++b) out[idx * ES + b] = 0;
      continue;
    }
    const int sy = reflect2(boxes[k].src_y + y, H, H1), sx = reflect2(boxes[k].src_x + (x - boxes[k].atlas_x), W, W1);
    const size_t si = nhwc ? ((static_cast<size_t>(n) * H + sy) * W + sx) * C + c
                           : ((static_cast<size_t>(n) * C + c) * H + sy) * W + sx;
#pragma unroll
    for (int b = 0; b < ES; ++b) out[idx * ES + b] = in[si * ES + b];
  }
}

template <int ES>
__global__ void tile_paste_kernel(const uint8_t* __restrict__ crops, uint8_t* __restrict__ out, const TileBox* __restrict__ boxes,
                                  int N, int C, int hco, int wco, int scale, int OH, int OW, int nhwc) {
  const int k = blockIdx.y / N, n = blockIdx.y - k * N;
  const TileBox bx = boxes[k];
  // the pasted rectangle, clipped to the output frame (the pre_pad / mod-pad margin is cropped off, RealESRGANer.post_process)
  const int ph = min(bx.paste_h, OH - bx.dst_y), pw = min(bx.paste_w, OW - bx.dst_x);
  if (ph <= 0 || pw <= 0) return;
  const size_t total = static_cast<size_t>(C) * ph * pw;
  const size_t img = static_cast<size_t>(bx.img) * N + n;
  for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t r = idx;
    int c, y, x;
    if (nhwc) { c = static_cast<int>(r % C); r /= C; x = static_cast<int>(r % pw); y = static_cast<int>(r / pw); }
    else { x = static_cast<int>(r % pw); r /= pw; y = static_cast<int>(r % ph); c = static_cast<int>(r / ph); }
    const int cy = bx.off_y + y, cx = bx.atlas_x * scale + bx.off_x + x, oy = bx.dst_y + y, ox = bx.dst_x + x;
    const size_t si = nhwc ? ((img * hco + cy) * wco + cx) * C + c : ((img * C + c) * hco + cy) * wco + cx;
    const size_t di = nhwc ? ((static_cast<size_t>(n) * OH + oy) * OW + ox) * C + c
                           : ((static_cast<size_t>(n) * C + c) * OH + oy) * OW + ox;
#pragma unroll
    for (int b = 0; b < ES; ++b) out[di * ES + b] = crops[si * ES + b];
  }
}

int fmt_elem_size(int fmt) { return fmt == 0 ? 4 : (fmt == 1 ? 2 : (fmt == 2 ? 1 : 0)); }

}  // namespace

cudaError_t tile_gather_launch(int fmt, const void* in, void* out, const TileBox* boxes_dev, int nbox, int nimg, int N, int C, int H,
                               int W, int pre_pad, int hc, int wc, cudaStream_t s) {
  const int H1 = H + pre_pad, W1 = W + pre_pad;
  const int es = fmt_elem_size(fmt);
  if (es == 0) return cudaErrorInvalidValue;
  const size_t total = static_cast<size_t>(nimg) * N * C * hc * wc;
  const unsigned blocks = static_cast<unsigned>(std::min<size_t>((total + 255) / 256, 148 * 16));
  const uint8_t* i8 = reinterpret_cast<const uint8_t*>(in);
  uint8_t* o8 = reinterpret_cast<uint8_t*>(out);
  if (es == 4) tile_gather_kernel<4><<<blocks, 256, 0, s>>>(i8, o8, boxes_dev, nbox, nimg, N, C, H, W, H1, W1, hc, wc, 0);
  else if (es == 2) tile_gather_kernel<2><<<blocks, 256, 0, s>>>(i8, o8, boxes_dev, nbox, nimg, N, C, H, W, H1, W1, hc, wc, 0);
  else tile_gather_kernel<1><<<blocks, 256, 0, s>>>(i8, o8, boxes_dev, nbox, nimg, N, C, H, W, H1, W1, hc, wc, 1);
  return cudaGetLastError();
}

cudaError_t tile_paste_launch(int fmt, const void* crops, void* out, const TileBox* boxes_dev, int nbox, int N, int C, int hco,
                              int wco, int scale, int OH, int OW, cudaStream_t s) {
  const int es = fmt_elem_size(fmt);
  if (es == 0) return cudaErrorInvalidValue;
  const size_t per = static_cast<size_t>(C) * hco * wco;
  dim3 grid(static_cast<unsigned>(std::min<size_t>((per + 255) / 256, 256)), static_cast<unsigned>(nbox * N));
  const uint8_t* i8 = reinterpret_cast<const uint8_t*>(crops);
  uint8_t* o8 = reinterpret_cast<uint8_t*>(out);
  if (es == 4) tile_paste_kernel<4><<<grid, 256, 0, s>>>(i8, o8, boxes_dev, N, C, hco, wco, scale, OH, OW, 0);
  else if (es == 2) tile_paste_kernel<2><<<grid, 256, 0, s>>>(i8, o8, boxes_dev, N, C, hco, wco, scale, OH, OW, 0);
  else tile_paste_kernel<1><<<grid, 256, 0, s>>>(i8, o8, boxes_dev, N, C, hco, wco, scale, OH, OW, 1);
  return cudaGetLastError();
}

}  // namespace ss4k
