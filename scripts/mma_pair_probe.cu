// Probe for the CTA-pair (cta_group::2, M = 256) form of the streaming conv's MMA:
//   * which half of the N columns of D comes from which CTA's B tile (the weight-operand split the pair kernel relies on),
//   * clocks per MMA for N = 32 / 96 with both operands in shared memory, against cta_group::1 (mma_issue_probe.cu: 40 / 56).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mma_pair_probe.bin scripts/mma_pair_probe.cu
// A = 1.0 everywhere; CTA r's B row j holds the value 100 r + j + 1 in every K column, so after one K = 16 MMA
// D[m][n] = 16 * (value of the B row that produced column n).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(0x40004040u) : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar) {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
               ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
    if (clock64() - t0 > 2000000000ll) return false;
  }
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_probe(int iters, long long* out_clk, float* out_val, int* out_flag) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const uint32_t a_s = base, b_s = base + 16384, bar = base + 16384 + 32768, slot = bar + 16;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  // A: ones.  B: row j -> 100 rank + j + 1 (rows are 128 bytes, every K column equal, so the swizzle does not matter)
  for (uint32_t i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0x3C003C00u;
  for (uint32_t i = threadIdx.x; i < 32768 / 4; i += blockDim.x) {
    const uint32_t row = i / 32;
    const __half h = __float2half(static_cast<float>(100 * rank + row + 1));
    const uint16_t u = *reinterpret_cast<const uint16_t*>(&h);
    reinterpret_cast<uint32_t*>(gen + 16384)[i] = (static_cast<uint32_t>(u) << 16) | u;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
  // M = 256 (both CTAs' 128 rows), N, K = 16, fp16 x fp16 -> fp32
  const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (warp == 1 && rank == 0) {
    t0 = clock64();
    umma2(tbase, a_s >> 4, b_s >> 4, idesc, 0u);
    for (int i = 1; i < iters; ++i) umma2(tbase, (a_s >> 4) + 2 * (i & 3), (b_s >> 4) + 2 * (i & 3), idesc, 1u);
    t1 = clock64();
    commit2(bar);
  }
  bool ok = true;
  if (warp == 1) ok = mbar_wait_bounded(bar, 0);
  const long long t2 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 1 && blockIdx.x < 2) {
    if (lane == 0) {
      out_flag[blockIdx.x] = ok ? 1 : -1;
      out_flag[2 + blockIdx.x] = static_cast<int>(tbase);
      if (rank == 0) { out_clk[0] = t1 - t0; out_clk[1] = t2 - t0; }
    }
    if (ok) {
      // D row (lane 32 + lane of this CTA), all N columns
      for (int c = 0; c < N; ++c) {
        uint32_t v;
        const uint32_t taddr = tbase + (static_cast<uint32_t>(32) << 16) + c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (lane == 3) out_val[blockIdx.x * 256 + c] = __uint_as_float(v);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

template <int N>
void run(int iters) {
  long long* clk; float* val; int* flag;
  cudaMalloc(&clk, 4 * sizeof(long long)); cudaMalloc(&val, 512 * sizeof(float)); cudaMalloc(&flag, 4 * sizeof(int));
  cudaMemset(clk, 0, 4 * sizeof(long long)); cudaMemset(val, 0, 512 * sizeof(float)); cudaMemset(flag, 0, 4 * sizeof(int));
  cudaFuncSetAttribute(pair_probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  pair_probe<N><<<148, 128, 60000>>>(iters, clk, val, flag);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[4]; float v[512]; int f[4];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(v, val, sizeof(v), cudaMemcpyDeviceToHost);
  cudaMemcpy(f, flag, sizeof(f), cudaMemcpyDeviceToHost);
  printf("pair N=%d iters=%d %s: wait flags %d %d, tmem base %d %d, issue clk %lld, done clk %lld -> %.1f clk per MMA\n", N, iters,
         cudaGetErrorString(e), f[0], f[1], f[2], f[3], h[0], h[1], (double)h[1] / iters);
  if (iters == 1) {
    for (int cta = 0; cta < 2; ++cta) {
      printf("  CTA %d D row 35, columns 0..%d (value / 16 = B row id: 100 rank + row + 1):", cta, N - 1);
      for (int c = 0; c < N; ++c) printf(" %g", v[cta * 256 + c] / 16.f);
      printf("\n");
    }
  }
  cudaFree(clk); cudaFree(val); cudaFree(flag);
}

int main() {
  run<32>(1);
  run<96>(1);
  run<32>(2000);
  run<64>(2000);
  run<96>(2000);
  run<192>(2000);
  return 0;
}
