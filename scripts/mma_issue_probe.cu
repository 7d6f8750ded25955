// Probe: tcgen05.mma issue rate from one vs two issuing warps of the same CTA, and whether MMAs issued by two
// different warps may accumulate into the SAME TMEM columns (needed for a two-issuer conv kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mma_issue_probe.bin scripts/mma_issue_probe.cu
// Modes: 0 one warp issues 2*ITERS MMAs; 1 two warps issue ITERS each into different columns;
//        2 two warps issue ITERS each into the same columns.  A = B = 1.0 (fp16), so every MMA adds 16.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void umma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(0x40004040u) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

template <int N>
__global__ void __launch_bounds__(128, 1) probe(int mode, int iters, long long* out_clk, float* out_val) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const uint32_t a_s = base, b_s = base + 16384, bar = base + 16384 + 32768, slot = bar + 16;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(2) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (warp == 1 || warp == 2) {
    const bool active = mode != 0 || warp == 1;
    const int n = mode == 0 ? 2 * iters : iters;
    const uint32_t col = (mode == 1 && warp == 2) ? 256u : 0u;
    if (active) {
      // first MMA of each accumulator zero-initialises: in mode 2 warp 1 initialises, warp 2 waits a little
      if (mode == 2 && warp == 2) __nanosleep(2000);
      t0 = clock64();
      umma(col, a_s >> 4, b_s >> 4, idesc, (mode == 2 && warp == 2) ? 1u : 0u);
      for (int i = 1; i < n; i += 4) {  // K steps 0..3 of the 64-wide tile, like the conv kernel
        umma(col, (a_s >> 4) + 2, (b_s >> 4) + 2, idesc, 1u);
        if (i + 1 < n) umma(col, (a_s >> 4) + 4, (b_s >> 4) + 4, idesc, 1u);
        if (i + 2 < n) umma(col, (a_s >> 4) + 6, (b_s >> 4) + 6, idesc, 1u);
        if (i + 3 < n) umma(col, (a_s >> 4), (b_s >> 4), idesc, 1u);
      }
      t1 = clock64();
      commit(bar);
    } else {
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
    }
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (lane == 0 && blockIdx.x == 0) {
      out_clk[(warp - 1) * 2] = t1 - t0;
      out_clk[(warp - 1) * 2 + 1] = t2 - t0;
    }
    // read back column 0 (and 256) of this warp's lane quarter
    uint32_t v;
    const uint32_t taddr = (static_cast<uint32_t>((warp & 3) * 32) << 16) + col;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (blockIdx.x == 0) out_val[(warp - 1) * 32 + lane] = __uint_as_float(v);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(0u), "r"(512u) : "memory");
}

template <int N>
void run(int iters) {
  long long* clk; float* val;
  cudaMalloc(&clk, 4 * sizeof(long long)); cudaMalloc(&val, 64 * sizeof(float));
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(clk, 0, 4 * sizeof(long long)); cudaMemset(val, 0, 64 * sizeof(float));
    probe<N><<<148, 128, 60000>>>(mode, iters, clk, val);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[4]; float v[64];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(v, val, sizeof(v), cudaMemcpyDeviceToHost);
    const int total = 2 * iters;
    float vmin = 1e30f, vmax = -1e30f;
    for (int i = 0; i < 64; ++i) { if (mode == 0 && i >= 32) break; vmin = v[i] < vmin ? v[i] : vmin; vmax = v[i] > vmax ? v[i] : vmax; }
    const long long span = h[1] > h[3] ? h[1] : h[3];
    printf("N=%d mode=%d %s: issue clk w1=%lld w2=%lld, done clk w1=%lld w2=%lld -> %.1f clk per MMA (aggregate, %d MMAs); "
           "accumulator min=%.0f max=%.0f (expect %d)\n", N, mode, cudaGetErrorString(e), h[0], h[2], h[1], h[3],
           (double)span / total, total, vmin, vmax, mode == 1 ? 16 * iters : 16 * total);
  }
  cudaFree(clk); cudaFree(val);
}

// issue cost: K MMAs from an idle pipe (queue never fills for small K): clocks from first to after last issue
template <int N>
__global__ void __launch_bounds__(128, 1) probe_issue(long long* out_clk) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const uint32_t a_s = base, b_s = base + 16384, bar = base + 16384 + 32768, slot = bar + 16;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
  if (warp == 1) {
    uint32_t ph = 0;
    for (int k = 1; k <= 12; ++k) {
      const long long t0 = clock64();
      const uint32_t al = a_s >> 4, bl = b_s >> 4;
      umma(0, al, bl, idesc, 0u);
      if (k > 1) umma(0, al + 2, bl + 2, idesc, 1u);
      if (k > 2) umma(0, al + 4, bl + 4, idesc, 1u);
      if (k > 3) umma(0, al + 6, bl + 6, idesc, 1u);
      if (k > 4) umma(0, al + 8, bl + 8, idesc, 1u);
      if (k > 5) umma(0, al + 10, bl + 10, idesc, 1u);
      if (k > 6) umma(0, al + 12, bl + 12, idesc, 1u);
      if (k > 7) umma(0, al + 14, bl + 14, idesc, 1u);
      if (k > 8) umma(0, al + 16, bl + 16, idesc, 1u);
      if (k > 9) umma(0, al + 18, bl + 18, idesc, 1u);
      if (k > 10) umma(0, al + 20, bl + 20, idesc, 1u);
      if (k > 11) umma(0, al + 22, bl + 22, idesc, 1u);
      const long long t1 = clock64();
      commit(bar);
      const long long t2 = clock64();
      mbar_wait(bar, ph);
      ph ^= 1;
      const long long t3 = clock64();
      if (lane == 0 && blockIdx.x == 0) { out_clk[3 * k] = t1 - t0; out_clk[3 * k + 1] = t2 - t1; out_clk[3 * k + 2] = t3 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(0u), "r"(512u) : "memory");
}
template <int N>
void run_issue() {
  long long* clk; cudaMalloc(&clk, 64 * sizeof(long long)); cudaMemset(clk, 0, 64 * sizeof(long long));
  cudaFuncSetAttribute(probe_issue<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  probe_issue<N><<<148, 128, 60000>>>(clk);
  cudaDeviceSynchronize();
  probe_issue<N><<<148, 128, 60000>>>(clk);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[64]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  printf("N=%d issue probe (%s): K: issue clk / commit clk / total to completion\n", N, cudaGetErrorString(e));
  for (int k = 1; k <= 12; ++k) printf("  K=%2d  %5lld  %4lld  %5lld\n", k, h[3 * k], h[3 * k + 1], h[3 * k + 2]);
}

int main() {
  run_issue<32>();
  run_issue<96>();
  run<32>(2000);
  run<96>(2000);
  run<192>(2000);
  run<256>(2000);
  return 0;
}
