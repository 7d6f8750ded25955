// Fused residual dense block for sm_100a: the five 3x3 convs of one basicsr ResidualDenseBlock in ONE persistent
// launch of the row-streaming tcgen05 / TMEM / TMA kernel (conv_stream.cu).
//
// Replaces (reference call sites: src/upscale/model/realesrgan/factory.py:113-125, basicsr rrdbnet_arch.py
// ResidualDenseBlock.forward, restated in SURVEY.md Appendix A):
//     x1 = lrelu(conv1(x)); x2 = lrelu(conv2(cat(x, x1))); ... ; x5 = conv5(cat(x, x1..x4)); return x5 * 0.2 + x
//
// Why: at one 720p frame per step a trunk conv is 12 output rows per SM; the kernel boundary between two convs (grid
// drain, launch, barrier / TMEM set-up, weight load, first TMA round trip) cost 5.3 us of a 25-55 us conv (DESIGN.md
// section 4.1).  Here the frame is cut into column strips x row bands with the same band boundaries in every strip, one
// CTA per (image, strip, band), all co-resident (grid <= SM count, one CTA per SM).  A CTA keeps its band through six
// phases (conv1..4, conv5 as two 32-wide chunks): barriers, TMEM and the accumulator ring are set up once and the
// activation-slab / accumulator pipelines keep running across a phase change.  What a phase reads from an earlier one
// (its own band plus one halo row and one halo pixel column of the neighbouring bands) is guarded by per-warp progress
// counters in global memory, published by the epilogue warps of the CTA that wrote the rows (release / acquire at gpu
// scope, a proxy fence on both sides because the data moves through TMA).
//
// A phase boundary only stays free if nobody has to WAIT for those counters: store completion + release + acquire +
// TMA round trip is as long as a kernel boundary.  So every other phase shifts the bands up by half a band (rows
// falling off the top wrap to the bottom of the same strip): the first and last rows of a phase then depend on rows
// the previous phase produced in the MIDDLE of its bands, half a band (6 rows, > 10 us) earlier, and because the bands
// are aligned across strips the horizontal neighbours are at the same row at the same time.  Dependencies only point
// to earlier phases, so the waits cannot deadlock; they are bounded (trap after 4 s) like every wait of the engine.
//
// Weights stream through a ring of three K-block groups (3 x 36 KB, what the widest conv needs): the producer requests
// the next phase's groups as soon as the slots they go into have been released by the current phase's LAST row (the
// MMA warp commits a slot right after that row's MMAs on it), so phase 1 is fully resident before phase 0 ends and the
// later phases trickle in behind the last row instead of waiting for the whole previous phase to drain.
//
// Inside a phase everything is conv_stream.cu's scheme (see there): M = 128 pixels of a row, N = 3 vertical taps x 32
// output channels, K = 64-channel blocks x 3 horizontal taps; accumulator ring in TMEM initialised with the fp32 bias by
// the epilogue warps; row records planned by the producer warp; epilogue = LeakyReLU (conv1..4) or scaled residual
// adds (conv5), 16-bit pack, swizzled staging tile, TMA store into the slab's growth channels / the next slab's x.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_common.cuh"
#include "conv_params.h"

namespace ss4k {

namespace {

constexpr int NOUT = kRdbNout;
constexpr uint32_t kWTile = 3u * NOUT * 128u;      // one (K block, horizontal tap) weight tile
constexpr uint32_t kStageWarp = 32u * NOUT * 2u;   // one epilogue warp's 32-pixel output tile (2 KB)
constexpr int kInFlight = 2;                       // TMA stores per epilogue warp that may still be in flight when older rows are published
constexpr int kWGroups = kRdbMaxWTiles / 3;        // weight ring: K-block groups of three tiles resident at a time

__device__ __forceinline__ uint32_t elect_one_f() {
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
__device__ __forceinline__ void tma_load_2d_f(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d_f(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;"
               :
               : "l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void pdl_launch_dependents_f() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_f() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit_f() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0_f() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0_f() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void sts128_f(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// non-blocking phase test (mbarrier.try_wait may suspend the thread up to a system-dependent time limit: a polling loop
// that also has other things to look after must not use it)
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// row record flags (word 5): bits 0..1 accumulator blocks after the ring wrap, bits 8..11 K blocks of the row,
// bits 12..14 k-steps of the last K block
constexpr uint32_t kRecLast = 4u, kRecNewW = 8u, kRecFreeW = 16u;
constexpr uint32_t kRecWords = 8u;

struct Band {
  int p, yb, ye;
};
// This CTA's place in the frame: image n, column strip, rows [v0, v1) of the strip in band coordinates.  In a shifted
// phase band coordinate v is image row v - half (mod H): the band moves up, what falls off the top wraps to the bottom.
struct Place {
  int n, strip, v0, v1;
};
__device__ __forceinline__ Place cta_place(const RdbParams& P, unsigned cta) {
  Place pl;
  const int B = P.bands;
  const int sj = static_cast<int>(cta) / B, bi = static_cast<int>(cta) - sj * B;
  pl.n = sj / P.strips;
  pl.strip = sj - pl.n * P.strips;
  pl.v0 = bi * P.H / B;
  pl.v1 = (bi + 1) * P.H / B;
  return pl;
}
struct Cursor {
  int p, k;   // phase, rows of the phase already handed out
};
// A band that contains the wrap point of a shifted phase (band coordinate v < half, i.e. the strip's first band) starts
// `rot` rows into its range and takes the wrapped rows LAST: they depend on the rows the strip's last band finished the
// previous phase with.
__device__ __forceinline__ int band_rot(const RdbParams& P, int p, int v0) {
  return (P.ph[p].shift && v0 < P.half) ? P.half - v0 : 0;
}
// next run of consecutive image rows of this CTA, phase after phase
__device__ __forceinline__ bool next_band(const RdbParams& P, const Place& pl, Cursor& c, Band& b) {
  const int nrows = pl.v1 - pl.v0;
  if (c.k >= nrows) {
    if (++c.p >= kRdbPhases) return false;
    c.k = 0;
  }
  const int off = P.ph[c.p].shift ? P.half : 0;
  const int rot = band_rot(P, c.p, pl.v0);
  int v, cnt;
  if (c.k < nrows - rot) { v = pl.v0 + rot + c.k; cnt = nrows - rot - c.k; }
  else { v = pl.v0 + (c.k - (nrows - rot)); cnt = nrows - c.k; }
  int y = v - off;
  if (y < 0) y += P.H;   // the wrapped piece: rows at the bottom of the strip
  b.p = c.p;
  b.yb = y;
  b.ye = y + cnt;
  c.k += cnt;
  return true;
}

}  // namespace

__global__ void __launch_bounds__(kRdbThreads, 1)
rdb_fused_kernel(const __grid_constant__ RdbParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = smem_base;
  const uint32_t bias_base = w_base + kRdbMaxWTiles * kWTile;
  const uint32_t a_base = bias_base + kStreamBiasBytes;
  const uint32_t stage_base = a_base + static_cast<uint32_t>(P.a_slots) * kASlotBytes;
  const uint32_t bar_base = stage_base + kStreamEpiWarps * ((kStageWarp + 1023u) & ~1023u);
  const uint32_t a_full = bar_base;
  const uint32_t a_empty = a_full + 8 * kMaxSASlots;
  const uint32_t acc_full = a_empty + 8 * kMaxSASlots;
  const uint32_t acc_empty = acc_full + 8 * kMaxAccSlots;
  // weights live in a ring of kWGroups slots, one slot = the three horizontal-tap tiles of one K block (36 KB)
  const uint32_t wg_full = acc_empty + 8 * kMaxAccSlots;
  const uint32_t wg_empty = wg_full + 8 * kWGroups;
  const uint32_t tmem_slot = wg_empty + 8 * kWGroups;
  const uint32_t rec_base = tmem_slot + 32;  // row records, one per activation slab slot
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // warp-uniform for ptxas
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents_f();
  {
    const uint64_t* pw = reinterpret_cast<const uint64_t*>(&P);
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < static_cast<int>(sizeof(RdbParams) / 64); ++i) acc ^= pw[i * 8];
    asm volatile("" ::"l"(acc));
  }
  const int S = P.acc_slots;
  const Place pl = cta_place(P, blockIdx.x);
  const int nrows = pl.v1 - pl.v0;            // rows of this CTA per phase; position q of its output-row sequence is phase q / nrows
  const int qtot = nrows * kRdbPhases;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&P.tmA);
      prefetch_tmap(&P.tmW);
      prefetch_tmap(&P.tmO[0]);
      prefetch_tmap(&P.tmO[1]);
    }
    if (lane < 2 * kWGroups) mbar_init(wg_full + 8 * lane, 1);   // wg_full[0..2], wg_empty[0..2] are contiguous
    if (lane < kMaxSASlots) {
      mbar_init(a_full + 8 * lane, 1);
      mbar_init(a_empty + 8 * lane, 1);
    }
    if (lane >= 16 && lane < 16 + kMaxAccSlots) {
      mbar_init(acc_full + 8 * (lane - 16), 1);
      mbar_init(acc_empty + 8 * (lane - 16), 4);
    }
    fence_barrier_init();
    __syncwarp();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(static_cast<uint32_t>(kTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    float* sb = reinterpret_cast<float*>(smem_gen + (bias_base - smem_base));
    const int nb = P.ph[kRdbPhases - 1].bias0 + NOUT;
    if (warp != 0)
      for (int i = threadIdx.x - 32; i < nb; i += kRdbThreads - 32) sb[i] = __ldg(P.bias_f + i);
  }
  if (warp == 0) {
    asm volatile("bar.arrive 1, %0;" ::"n"(kRdbThreads) : "memory");
  } else {
    tcgen05_before_sync();
    asm volatile("bar.sync 1, %0;" ::"n"(kRdbThreads) : "memory");
    tcgen05_after_sync();
    if (*tmem_slot_ptr != 0u) __trap();  // one CTA per SM owns all 512 columns: the allocation starts at column 0
  }
  constexpr uint32_t tmem_base = 0u;

  if (warp == 0) {
    // ======================================================= TMA producer + row planner + dependency tracker
    uint32_t as = 0, aph = 0;
    int loaded_w = -1;  // phase whose first row has been planned (kRecNewW goes with it)
    // ---- weight stream: K-block groups of all phases, in order, through the ring of kWGroups slots
    int gl = 0, gl_p = 0, gl_kb = 0;   // next group to request: index in the launch, its phase, its K block
    int gtot = 0;
#pragma unroll
    for (int i = 0; i < kRdbPhases; ++i) gtot += P.ph[i].nkb;
    // requests every group whose slot is free; never blocks (every spin loop of this warp calls it: the MMA warp may be
    // waiting for weights whose slot was released while this warp is waiting for something else)
    auto try_weights = [&]() {
      while (gl < gtot) {
        const int slot = gl % kWGroups;
        if (gl >= kWGroups && !mbar_test_wait(wg_empty + 8 * slot, static_cast<uint32_t>((gl / kWGroups - 1) & 1))) return;
        if (elect_one_f()) {
          mbar_expect_tx(wg_full + 8 * slot, 3u * kWTile);
#pragma unroll
          for (int t = 0; t < 3; ++t)
            tma_load_2d_f(w_base + static_cast<uint32_t>(slot * 3 + t) * kWTile, &P.tmW, wg_full + 8 * slot, 0,
                          P.ph[gl_p].w_row0 + (gl_kb * 3 + t) * 3 * NOUT);
        }
        __syncwarp();
        ++gl;
        if (++gl_kb == P.ph[gl_p].nkb) { gl_kb = 0; ++gl_p; }
      }
    };
    // waits for an mbarrier phase while keeping the weight stream going; bounded like every wait of the engine
    auto wait_bar = [&](uint32_t bar, uint32_t parity) {
      if (mbar_test_wait(bar, parity)) return;
      const uint64_t t0 = globaltimer_ns();
      for (uint32_t it = 0; !mbar_test_wait(bar, parity); ++it) {
        try_weights();
        if ((it & 1023u) == 1023u && globaltimer_ns() - t0 > 4000000000ull) asm volatile("trap;");
      }
    };
    try_weights();   // weights are constants of the launch: phase 0 and phase 1 are requested before anything else
    bool dep_ready = false;
    int sL = 0, kL = 0;
    const bool use_ctr = !(P.dbg_flags & 1);
    const int B = P.bands, H = P.H;
    long long n_wait = 0, clk_wait = 0, n_poll = 0, clk_poll = 0, clk_aempty = 0, ph_start[kRdbPhases] = {0, 0, 0, 0, 0, 0};
    const long long clk_begin = clock64();
    const bool tracing = P.trace != nullptr;
    Cursor cur;
    cur.p = 0; cur.k = 0;
    Band b, nb;
    bool has = next_band(P, pl, cur, b);
    while (has) {
      const bool has_next = next_band(P, pl, cur, nb);
      const RdbPhase& ph = P.ph[b.p];
      const bool new_w = b.p != loaded_w;
      loaded_w = b.p;
      if (!dep_ready) {  // x was written by the previous launch
        pdl_wait_f();
        dep_ready = true;
      }
      const uint64_t pol_in = l2_policy(ph.l2_in);
      int gbase = 0;   // index of the phase's first weight group in the launch
#pragma unroll
      for (int i = 0; i < kRdbPhases; ++i)
        if (i < b.p) gbase += P.ph[i].nkb;
      const int r0 = b.yb > 0 ? b.yb - 1 : b.yb;
      const int r1 = b.ye < H ? b.ye : b.ye - 1;
      const int x0 = pl.strip * kTileW - 1;
      int y_lo = b.yb;
      // rows [ok_lo, ok_hi) written by phase ph.dep are known to be complete in this strip and its two neighbours
      int ok_lo = 0, ok_hi = 0;
      const int dep = ph.dep;
      const int doff = dep >= 0 && P.ph[dep >= 0 ? dep : 0].shift ? P.half : 0;
      for (int r = r0; r <= r1; ++r) {
        try_weights();
        // ---- the rows of phase `dep` this input row reads (the newest input channels; older channels follow by
        //      transitivity: a row is published after everything it was computed from had been acquired).  The bands are
        //      aligned across strips, so the three owners (this strip and its neighbours) sit at the same band / position:
        //      their 3 x 8 counters are read by 24 lanes at once, one L2 round trip per poll.
        if (dep >= 0 && use_ctr && !(r >= ok_lo && r < ok_hi)) {
          int v = r + doff;                                          // band coordinate of image row r in phase dep
          if (v >= H) v -= H;
          const int bi = ((v + 1) * B - 1) / H;                      // owner band: v0(bi) <= v < v0(bi + 1)
          const int vj0 = bi * H / B, vj1 = (bi + 1) * H / B;
          const int nj = vj1 - vj0;
          const int rot = band_rot(P, dep, vj0);
          int pos = v - vj0 - rot;                                   // the owners' processing position of this row inside the phase
          if (pos < 0) pos += nj;
          const int piece_end = v >= vj0 + rot ? vj1 : vj0 + rot;    // rows up to here follow v in the owners' order
          const uint32_t qbase = static_cast<uint32_t>(dep * nj);    // sequence position of the owners' first row of phase dep
          const uint32_t q = qbase + static_cast<uint32_t>(pos);
          const int d = lane >> 3;                                   // lanes 0..7: strip - 1, 8..15: strip, 16..23: strip + 1
          const int s2 = pl.strip + d - 1;
          const bool act = d < 3 && s2 >= 0 && s2 < P.strips;
          const uint32_t j = static_cast<uint32_t>((pl.n * P.strips + (act ? s2 : pl.strip)) * B + bi);
          const uint32_t* cp = P.ctr_use + j * kRdbCtrPerCta + (lane & 7);
          uint32_t V;
          uint64_t t0 = 0;
          long long c0 = 0;
          const long long cp0 = tracing ? clock64() : 0;
          ++n_poll;
          for (int it = 0;; ++it) {
            const uint32_t c = act ? ld_acquire_u32(cp) : 0x3FFFFFFFu;
            // counters 0..3 of a CTA: rows with even sequence position, 4..7: odd (one per TMEM lane quarter)
            uint32_t m = c;
            m = min(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = min(m, __shfl_xor_sync(0xffffffffu, m, 2));
            const uint32_t c_even = __shfl_sync(0xffffffffu, m, lane & 24), c_odd = __shfl_sync(0xffffffffu, m, (lane & 24) + 4);
            V = min(2u * c_even, 2u * c_odd + 1u);                  // sequence positions < V are complete (per owner)
            V = min(V, __shfl_xor_sync(0xffffffffu, V, 8));
            V = min(V, __shfl_xor_sync(0xffffffffu, V, 16));          // ... in all three strips
            if (q < V) break;
            try_weights();
            if (it == 0) { t0 = globaltimer_ns(); c0 = clock64(); ++n_wait; }
            if (globaltimer_ns() - t0 > 4000000000ull) {   // bounded like every wait of the engine (no out-of-line call here)
              if (P.err != nullptr && lane == 0) {
                P.err[0] = 900 + b.p; P.err[1] = static_cast<int32_t>(blockIdx.x); P.err[2] = static_cast<int32_t>(bi); P.err[3] = static_cast<int32_t>(q);
                __threadfence_system();
              }
              asm volatile("trap;");
            }
          }
          if (c0 != 0) clk_wait += clock64() - c0;
          if (tracing) clk_poll += clock64() - cp0;
          int hi_v = v + (static_cast<int>(V - qbase) - pos);          // rows of the owners verified so far, band coordinates
          hi_v = hi_v < piece_end ? hi_v : piece_end;
          const int hi_r = r + (hi_v - v);
          ok_lo = r;
          ok_hi = hi_r < H ? hi_r : H;
          fence_proxy_async_all();   // the acquired data was written through the async proxy and is read by TMA
          __syncwarp();
        }
        // ---- the row's record (see conv_stream.cu)
        {
          const int y = r - 1 > b.yb ? r - 1 : b.yb;
          if (y != y_lo) {
            y_lo = y;
            if (++sL == S) { sL = 0; ++kL; }
          }
        }
        const int y_hi = r + 1 < b.ye - 1 ? r + 1 : b.ye - 1;
        const int b_lo = y_lo - (r - 1);
        const int nblk = y_hi - y_lo + 1;
        const int nA = sL + nblk <= S ? nblk : S - sL;
        const int nB = nblk - nA;
        uint32_t fresh = 0;
        {
          const int f_lo = (r == r0) ? y_lo : r + 1;
          int sh = 0;
          for (int y = f_lo; y <= y_hi; ++y, sh += 8) {
            int s2 = sL + (y - y_lo), k2 = kL;
            if (s2 >= S) { s2 -= S; ++k2; }
            fresh |= (0x80u | ((k2 & 1) ? 0x40u : 0u) | static_cast<uint32_t>(s2)) << sh;
          }
        }
        uint32_t c0 = 0xFFu, c1 = 0xFFu;
        if (r - 1 >= b.yb) c0 = static_cast<uint32_t>(sL);
        if (r == r1 && r <= b.ye - 1) {
          int s2 = sL + (r - y_lo);
          if (s2 >= S) s2 -= S;
          c1 = static_cast<uint32_t>(s2);
        }
        const bool w_changes = !has_next || nb.p != b.p;   // the phase's last row releases its weight slots
        const uint32_t flags = static_cast<uint32_t>(nB) | ((r == r1 && !has_next) ? kRecLast : 0u) |
                               ((r == r0 && new_w) ? kRecNewW : 0u) | ((r == r1 && w_changes) ? kRecFreeW : 0u) |
                               (static_cast<uint32_t>(ph.nkb) << 8) | (static_cast<uint32_t>(ph.nks_last) << 12) |
                               (static_cast<uint32_t>(gbase) << 16);
        if (tracing && ph_start[b.p] == 0) ph_start[b.p] = clock64() - clk_begin;
        for (int kb = 0; kb < ph.nkb; ++kb) {
          try_weights();
          if (tracing) {
            const long long ca = clock64();
            wait_bar(a_empty + 8 * as, aph ^ 1);
            clk_aempty += clock64() - ca;
          } else {
            wait_bar(a_empty + 8 * as, aph ^ 1);
          }
          if (elect_one_f()) {
            if (kb == 0) {
              const uint32_t ra = rec_base + as * (kRecWords * 4u);
              sts128_f(ra, static_cast<uint32_t>(sL * NOUT), P.idesc[nA - 1], P.idesc[nB > 0 ? nB - 1 : 0],
                       static_cast<uint32_t>(b_lo * NOUT * 128) >> 4);
              sts128_f(ra + 16, static_cast<uint32_t>((b_lo + nA) * NOUT * 128) >> 4, flags, c0 | (c1 << 8), fresh);
            }
            mbar_expect_tx(a_full + 8 * as, kBoxW * kRowBytes);
            tma_load_5d_hint(a_base + as * kASlotBytes, &P.tmA, a_full + 8 * as, 0, x0, kb, r, pl.n, pol_in);
          }
          __syncwarp();
          if (++as == static_cast<uint32_t>(P.a_slots)) { as = 0; aph ^= 1; }
        }
      }
      sL += b.ye - y_lo;
      if (sL >= S) { sL -= S; ++kL; }
      b = nb;
      has = has_next;
    }
    if (tracing && lane == 0) {
      long long* t = P.trace + blockIdx.x * 16;
      t[0] = n_wait; t[1] = clk_wait; t[2] = n_poll; t[3] = clk_poll; t[4] = clk_aempty; t[5] = clock64() - clk_begin;
#pragma unroll
      for (int i = 0; i < kRdbPhases; ++i) t[6 + i] = ph_start[i];
    }
    if (P.next_w != nullptr && lane == 0) {
      const uint32_t per = ((P.next_w_bytes + gridDim.x - 1) / gridDim.x + 15u) & ~15u;
      const uint32_t off = blockIdx.x * per;
      if (off < P.next_w_bytes) {
        const uint32_t sz = P.next_w_bytes - off < per ? ((P.next_w_bytes - off) & ~15u) : per;
        if (sz > 0)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const uint8_t*>(P.next_w) + off), "r"(sz) : "memory");
      }
    }
  } else if (warp == 1) {
    // ======================================================= MMA issuer (conv_stream.cu's loop; K blocks per row from the record)
    uint32_t as = 0, aph = 0;
    const uint32_t n_aslots = static_cast<uint32_t>(P.a_slots);
    uint32_t rc[kRecWords];
    auto fetch = [&](uint32_t slot, uint32_t ph) {
      mbar_wait_u(a_full + 8 * slot, ph);
      uint32_t v0, v1, v2, v3, v4, v5, v6, v7;
      const uint32_t ra = rec_base + slot * (kRecWords * 4u);
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(ra) : "memory");
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v4), "=r"(v5), "=r"(v6), "=r"(v7) : "r"(ra + 16) : "memory");
      rc[0] = __shfl_sync(0xffffffffu, v0, 0); rc[1] = __shfl_sync(0xffffffffu, v1, 0);
      rc[2] = __shfl_sync(0xffffffffu, v2, 0); rc[3] = __shfl_sync(0xffffffffu, v3, 0);
      rc[4] = __shfl_sync(0xffffffffu, v4, 0); rc[5] = __shfl_sync(0xffffffffu, v5, 0);
      rc[6] = __shfl_sync(0xffffffffu, v6, 0); rc[7] = __shfl_sync(0xffffffffu, v7, 0);
    };
    auto prepare = [&]() {
      tcgen05_after_sync();
      uint32_t f = rc[7];
#pragma unroll
      for (int j = 0; j < 3; ++j, f >>= 8) {
        if (f & 0x80u) {
          mbar_wait_u(acc_empty + 8 * (f & 0x1Fu), (f >> 6) & 1u);
          tcgen05_after_sync();
        }
      }
    };
    {
      fetch(0, 0);
      prepare();
      bool last = false;
      while (!last) {
        const uint32_t colA = tmem_base + rc[0], idA = rc[1], idB = rc[2], woffA = rc[3], woffB = rc[4];
        const uint32_t flags = rc[5], cc = rc[6];
        const bool wrap = (flags & 3u) != 0;
        last = (flags & kRecLast) != 0;
        const bool overlap = !last;
        const int gbase = static_cast<int>((flags >> 16) & 15u);
        const int nkb = static_cast<int>((flags >> 8) & 15u);
        const int nks_last = static_cast<int>((flags >> 12) & 7u);
        for (int kb = 0; kb < nkb; ++kb) {
          if (kb > 0) {
            mbar_wait_u(a_full + 8 * as, aph);
            tcgen05_after_sync();
          }
          // this K block's weights: group gbase + kb of the launch, ring slot (gbase + kb) % kWGroups; the phase's first row
          // waits for them, its last row hands the slot back
          const int wg = gbase + kb;
          const uint32_t wslot = static_cast<uint32_t>(wg % kWGroups);
          if (flags & kRecNewW) {
            mbar_wait_u(wg_full + 8 * wslot, static_cast<uint32_t>((wg / kWGroups) & 1));
            tcgen05_after_sync();
          }
          const uint32_t a_lo = (a_base + as * kASlotBytes) >> 4;
          const uint32_t w_lo = (w_base + wslot * 3u * kWTile) >> 4;
          const uint32_t wA_lo = w_lo + woffA, wB_lo = w_lo + woffB;
          const int nks = kb == nkb - 1 ? nks_last : 4;
          const uint32_t as_cur = as;
          if (++as == n_aslots) { as = 0; aph ^= 1; }
#define SS4K_MMA(KX, KS)                                                                                                   \
          {                                                                                                                \
            umma_f16_lo(colA, a_lo + ((KX) * kRowBytes + (KS) * 32) / 16, wA_lo + ((KX) * kWTile + (KS) * 32) / 16, idA, 1u); \
            if (wrap) umma_f16_lo(tmem_base, a_lo + ((KX) * kRowBytes + (KS) * 32) / 16, wB_lo + ((KX) * kWTile + (KS) * 32) / 16, idB, 1u); \
          }
          if (nks == 4) {
            SS4K_MMA(0, 0) SS4K_MMA(0, 1) SS4K_MMA(0, 2) SS4K_MMA(0, 3)
            SS4K_MMA(1, 0) SS4K_MMA(1, 1) SS4K_MMA(1, 2) SS4K_MMA(1, 3)
          } else {
#pragma unroll
            for (int kx = 0; kx < 2; ++kx) {
#pragma unroll
              for (int ks = 0; ks < 3; ++ks)
                if (ks < nks) SS4K_MMA(kx, ks)
            }
          }
          if (kb == nkb - 1 && overlap) {  // next row: first slab, record, fresh accumulator slots
            fetch(as, aph);
            prepare();
          }
          if (nks == 4) {
            SS4K_MMA(2, 0) SS4K_MMA(2, 1) SS4K_MMA(2, 2) SS4K_MMA(2, 3)
          } else {
#pragma unroll
            for (int ks = 0; ks < 3; ++ks)
              if (ks < nks) SS4K_MMA(2, ks)
          }
#undef SS4K_MMA
          umma_commit_elect(a_empty + 8 * as_cur);
          if (flags & kRecFreeW) umma_commit_elect(wg_empty + 8 * wslot);
        }
        if ((cc & 0xFFu) != 0xFFu) umma_commit_elect(acc_full + 8 * (cc & 0xFFu));
        if (((cc >> 8) & 0xFFu) != 0xFFu) umma_commit_elect(acc_full + 8 * ((cc >> 8) & 0xFFu));
        if (!last && !overlap) {
          fetch(as, aph);
          prepare();
        }
      }
    }
  } else {
    // ======================================================= epilogue (warps 2..9)
    const int ew = warp - 2;
    const int qd = warp & 3;       // TMEM lane quarter this warp may access
    const int par = ew >> 2;       // this warp takes the CTA's output rows whose sequence position has this parity
    const int m = qd * 32 + lane;
    const uint32_t stage = stage_base + static_cast<uint32_t>(ew) * ((kStageWarp + 1023u) & ~1023u);
    uint32_t* const my_ctr = P.ctr_use + blockIdx.x * kRdbCtrPerCta + ew;
    // bias offset of the row at sequence position qq of this CTA
    auto bias_of = [&](int qq) -> int { return P.ph[qq / nrows].bias0; };
    auto init_slot = [&](int s_, int boff) {
      const uint32_t ta = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + static_cast<uint32_t>(s_ * NOUT);
      const uint32_t ba = bias_base + static_cast<uint32_t>(boff) * 4u;
#pragma unroll
      for (int c = 0; c < NOUT; c += 16) {
        uint32_t bv[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) lds128(ba + (c + 4 * i) * 4u, bv[4 * i], bv[4 * i + 1], bv[4 * i + 2], bv[4 * i + 3]);
        tmem_st16p(ta + c, bv);
      }
      tmem_st_wait();
      tcgen05_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + 8 * s_);
    };
    for (int q0 = par; q0 < S && q0 < qtot; q0 += 2) init_slot(q0, bias_of(q0));
    pdl_wait_f();  // residual loads and output stores touch tensors of the previous launch
    // clear the counters of the NEXT fused launch (its buffer is used neither by this launch nor by the previous one)
    if (lane == 0) P.ctr_zero[blockIdx.x * kRdbCtrPerCta + ew] = 0u;
    if (P.discard_ptr != nullptr) {
      const int64_t per = (P.discard_npx + gridDim.x - 1) / gridDim.x;
      const int64_t p0 = per * blockIdx.x, p1 = p0 + per < P.discard_npx ? p0 + per : P.discard_npx;
      uint8_t* const db = reinterpret_cast<uint8_t*>(P.discard_ptr);
      for (int64_t px = p0 + (ew * 32 + lane); px < p1; px += kStreamEpiWarps * 32) {
        uint8_t* const pp = db + px * P.discard_pitch_bytes;
#pragma unroll
        for (int l = 0; l < 3; ++l)
          if ((P.discard_mask >> l) & 1u) asm volatile("discard.global.L2 [%0], 128;" ::"l"(pp + l * 128) : "memory");
      }
    }
    auto h2 = [](float a, float c) -> uint32_t {
      const __half2 h = __floats2half2_rn(a, c);
      return *reinterpret_cast<const uint32_t*>(&h);
    };
    int s = 0, k = 0, q = 0;
    uint32_t issued = 0, published = 0;  // rows of this warp whose TMA store has been committed / announced in its counter
    Cursor cur;
    cur.p = 0; cur.k = 0;
    Band b;
    const float sl = P.slope, b1 = P.beta1, b2 = P.beta2;
    const bool has_r2 = P.res2 != nullptr;
    const int ax = pl.strip * kTileW + m;
    const bool valid = ax < P.W;
    while (next_band(P, pl, cur, b)) {
      const RdbPhase& ph = P.ph[b.p];
      const uint64_t pol_out = l2_policy(ph.l2_out);
      const bool last_phase = ph.residual != 0;
      for (int y = b.yb; y < b.ye; ++y, ++q) {
        if ((q & 1) == par) {
          uint4 r1v[NOUT / 8], r2v[NOUT / 8];
          if (last_phase && valid) {
            const size_t pix = (static_cast<size_t>(pl.n) * P.H + y) * P.W + ax;
            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(P.res1) + pix * P.res1_pitch + P.res1_coff + ph.res_c);
#pragma unroll
            for (int j = 0; j < NOUT / 8; ++j) r1v[j] = rp[j];
            if (has_r2) {
              const uint4* rq = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(P.res2) + pix * P.res2_pitch + P.res2_coff + ph.res_c);
#pragma unroll
              for (int j = 0; j < NOUT / 8; ++j) r2v[j] = rq[j];
            }
          }
          // Never go to sleep on unannounced rows: if the accumulator is not ready yet (this warp has run dry), wait
          // for the stores in flight and publish them first.  Costs nothing while rows keep coming, and it is what makes
          // the protocol deadlock-free for any geometry: a CTA that waits for data holds nothing back.
          if (issued > published) {
            uint32_t ready = lane == 0 ? mbar_test_wait(acc_full + 8 * s, static_cast<uint32_t>(k & 1)) : 0u;
            ready = __shfl_sync(0xffffffffu, ready, 0);
            if (!ready) {
              if (lane == 0) {
                bulk_wait0_f();
                fence_proxy_async_all();
                red_release_add(my_ctr, issued - published);
              }
              published = issued;
            }
          }
          mbar_wait_u(acc_full + 8 * s, k & 1);
          tcgen05_after_sync();
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + static_cast<uint32_t>(s * NOUT);
          uint32_t raw[NOUT];
#pragma unroll
          for (int c = 0; c < NOUT; c += 16) tmem_ld16p(taddr + c, &raw[c]);
          tmem_ld_wait();
          if (q + S < qtot) init_slot(s, bias_of(q + S));
          // the staging tile is free once this warp's previous store has READ it (publication waits for completion: below)
          if (lane == 0) bulk_wait_read0_f();
          __syncwarp();
          const uint32_t srow = stage + lane * (NOUT * 2u);
          const uint32_t sxor = (lane >> 1) & 3u;  // swizzle-64B position of the 16-byte chunk inside the [32 pixels][32 ch] tile
          if (!last_phase) {
#pragma unroll
            for (int j = 0; j < NOUT / 8; ++j) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float a = __uint_as_float(raw[8 * j + i]);
                v[i] = fmaxf(a, a * sl);
              }
              sts128_f(srow + ((static_cast<uint32_t>(j) ^ sxor) << 4), h2(v[0], v[1]), h2(v[2], v[3]), h2(v[4], v[5]), h2(v[6], v[7]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < NOUT / 8; ++j) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(raw[8 * j + i]);
              if (valid) {
                const uint32_t w4[4] = {r1v[j].x, r1v[j].y, r1v[j].z, r1v[j].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w4[i]));
                  v[2 * i] = fmaf(b1, f.x, v[2 * i]);
                  v[2 * i + 1] = fmaf(b1, f.y, v[2 * i + 1]);
                }
                if (has_r2) {
                  const uint32_t x4[4] = {r2v[j].x, r2v[j].y, r2v[j].z, r2v[j].w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&x4[i]));
                    v[2 * i] = fmaf(b2, f.x, v[2 * i]);
                    v[2 * i + 1] = fmaf(b2, f.y, v[2 * i + 1]);
                  }
                }
              }
              sts128_f(srow + ((static_cast<uint32_t>(j) ^ sxor) << 4), h2(v[0], v[1]), h2(v[2], v[3]), h2(v[4], v[5]), h2(v[6], v[7]));
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d_f(&P.tmO[ph.out_map], stage, ph.out_c0, pl.strip * kTileW + qd * 32, y, pl.n, pol_out);
            bulk_commit_f();
            // Publication: a row may be announced once its store is COMPLETE (global writes performed), which takes a few
            // thousand clocks: waiting for the newest store would make every row pay that latency.  Up to kInFlight
            // stores stay in flight; everything older is complete and gets published (the shifted schedule gives every
            // dependency half a band of slack, and this warp keeps storing rows until well after the last row anyone
            // waits for -- conv4's -- so nothing is left unpublished).
            asm volatile("cp.async.bulk.wait_group %0;" ::"n"(kInFlight) : "memory");
            if (issued + 1 > published + kInFlight) {
              fence_proxy_async_all();
              red_release_add(my_ctr, issued + 1 - kInFlight - published);
            }
          }
          ++issued;                                        // (warp-uniform bookkeeping)
          if (issued > published + kInFlight) published = issued - kInFlight;
        }
        if (++s == S) { s = 0; ++k; }
      }
    }
    // (the last kInFlight rows of conv5's second chunk are never announced: nobody polls for them, the next launch
    //  waits for the whole grid; the staging tiles must have been read before the CTA releases its shared memory)
    if (lane == 0) bulk_wait_read0_f();
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
cudaError_t rdb_fused_prepare() {
  return cudaFuncSetAttribute(rdb_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
}

// One CTA per SM, all co-resident (the progress-counter waits rely on it); 0 if the kernel cannot be resident.
int rdb_fused_max_ctas_per_sm() {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, rdb_fused_kernel, kRdbThreads, kSmemBytes) != cudaSuccess) return 0;
  return n;
}

cudaError_t rdb_fused_launch(const RdbParams& p, int grid, cudaStream_t stream, bool pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kRdbThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, rdb_fused_kernel, p);
}

}  // namespace ss4k
