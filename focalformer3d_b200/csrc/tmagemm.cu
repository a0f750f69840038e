// TMA-fed tcgen05 implicit GEMM on PRE-SPLIT activations (round 2).
//
//   y[row(m), n] = act( sum_{t,c} x[src(m,t), c] * w[t,c,n] + bias[n] + res[m,n] )            (contract of tcgemm.cu)
//
// tcgemm.cu converts every A element to its fp16 hi/lo pair inside the kernel, in registers, once per (element, tap):
// 9x per element in a 3x3 conv, up to 27x in a sparse conv.  ncu shows those producer warps are INSTRUCTION-ISSUE bound
// (~50 warp instructions per 512 gathered bytes: address maths, LDG, clamp / convert / back-convert / scale / convert,
// two swizzled stores), and the tensor pipe waits for them -- switching the MMAs from kind::tf32 to kind::f16 halved the
// tensor time and changed the step time by < 5 %.
//
// Here activations LIVE in split form: a row of C channels is stored as [hi(C) | lo(C)] fp16 (4C bytes, the same
// footprint as fp32), written once by the epilogue of the layer that produces it.  The A operand then needs no
// instruction at all: it is moved global -> 128B-swizzled shared memory by the TMA unit,
//   ROWS    one 2-D tile load per plane and K-step            (cp.async.bulk.tensor.2d        -> SASS UTMALDG)
//   CONV2D  one 4-D tile load per plane, tap and K-step: box = (64 ch, bw, bh, 1) of the NHWC map, zero padding =
//           the TMA's out-of-bounds fill                      (cp.async.bulk.tensor.4d)
//   SPARSE  rulebook gather: 32 lanes x one gather4 per plane: four arbitrary rows per instruction straight into
//           their swizzled tile rows; absent neighbours read a dedicated all-zero row
//                                                             (cp.async.bulk.tensor.2d.tile::gather4)
// and the weights by cp.async.bulk as before.  192 threads: warp 0 = producer (TMA issue only), warp 1 = TMEM alloc +
// single-thread tcgen05.mma issue, warps 2-5 = epilogue (tcgen05.ld -> bias / residual / activation -> fp32 rows and /
// or split rows), which now overlaps the next tile's main loop completely (double-buffered TMEM accumulators).
// Per-tile tap skipping (ff3d_sp_nbr_build masks) as in tcgemm.cu.  K-step = 64 halves = one 128-byte swizzled row.
#include "tc_common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace ff3d {

struct TmP {
  int mode, M;
  const int* m_dev;
  int cin, cout, taps;
  int xs_lo;                     // halves between the hi and the lo plane of an A row
  const __half* xs; int ldxs;    // the split A rows themselves (cp.async gather of the SPARSE mode)
  const void* wimg;
  const float* bias;
  const float* res; int ldres;
  const __half* res_s; int ldres_s, res_s_lo;
  float* y; int ldy;
  __half* ys; int ldys, ys_lo;
  int act, res_after_act;
  // CONV2D (stride 1): output patch bw x bh per tile
  int B, Ho, Wo, kw, pad, bw, bh, tiles_x, tiles_y;
  long long y_bstride, res_bstride;   // rows per batch element of y / ys / res
  int stride;                         // CONV2D input stride (TMA traversal stride of the W / H dimensions)
  int ux, uy, dx, dy;                 // output lattice (transposed conv, kernel == stride): row = (oy*uy + dy, ox*ux + dx)
  long long y_row0;
  // SPARSE
  const int* nbr; int nbr_stride;
  const int* y_off;              // element offsets into y (BEV scatter), fp32 output only
  const int* y_row;              // output row map (row index) for y / ys
  const uint32_t* tile_mask;
  int zero_row;
  int* overflow;
  int n_stages, cpt;             // pipeline K-steps in the weight images; 64-channel chunks per tap (cin >= 64)
  int tps, n_units;              // SPARSE, cin < 64: taps packed into one 64-wide K-step; skippable units (taps | stages)
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, int col, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}

// CPA (SPARSE only): gather the rows with 16-byte cp.async (LDGSTS) copies issued by the NPW producer warps instead of
// TMA gather4.  Measured: the TMA unit retires one gather4 (4 rows x 128 bytes) every ~55 cycles per SM whatever the number
// of issuing warps (~1.3 TB/s over the chip), 2-3x short of the MMA rate; LDGSTS moves the same pre-split bytes through the
// LSU path with ~4 instructions per 16 bytes and no registers; cp.async.mbarrier.arrive.noinc signals each stage exactly
// when its copies have landed.
// NPW = producer warps.  ROWS / CONV2D need one (a single thread issues two box loads per stage).  SPARSE issues
// 2 x 32 gather4 per stage, and a gather4 costs ~100 cycles of the issuing warp (measured: one producer warp delivers a
// stage every ~6900 cycles against 768 cycles of MMA time), so the gathers of a stage are spread over NPW warps.
// 16-byte global -> shared copy without registers (LDGSTS); src_bytes = 0 writes zeros (absent neighbour)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on `bar` when all cp.async copies issued so far by this thread have completed; .noinc: the arrival is one of the
// barrier's expected arrivals (it was counted at init)
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

template <int MODE, int BN, int NS, int NPW, bool CPA = false>
__global__ void __launch_bounds__(32 * (NPW + 5 + (CPA ? 1 : 0)), 1) tmagemm_kernel(const __grid_constant__ CUtensorMap tmA, const TmP p) {
  constexpr int MMA_WARP = NPW;
  constexpr int IDX_WARP = NPW + 5;                               // CPA only: stages the tiles' neighbour maps in shared memory
  // Chunked accumulation (dense modes, one producer warp): the tensor core's fp32 accumulate TRUNCATES, once per MMA
  // (16 products); over a long K that is a systematic bias (measured at K = 8192, all-positive products: -4.7e-5 relative,
  // 10x the fp32 SIMT kernel, gone when the same K is summed as 8 chunks).  So the K loop is cut into chunks of CHUNK
  // stages (32 MMA steps): every chunk accumulates into its own TMEM buffer (the two buffers alternate) and the epilogue
  // warps -- idle during the main loop anyway -- add each finished chunk into fp32 REGISTERS with round-to-nearest.
  // Short K loops (<= CHUNK_MIN stages: the 3x3 convs on <= 256 channels) stay one run -- their bias is inside the parity
  // budget and every extra drain is epilogue-warp time; the long ones (ROI MLP 294 stages, LSS bevencode 117, shared conv
  // 72) are cut.  Measured at K = 8192: relative error 4.9e-5 (mean -4.7e-5) un-chunked -> 2.6e-6 chunked by 8 stages.
  constexpr bool CHUNKED = (NPW == 1);
  constexpr int CHUNK = 16;
  constexpr int CHUNK_MIN = 40;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t A_BYTES = TC_BM * 128;
  constexpr uint32_t B_BYTES = BN * 128;
  constexpr uint32_t SLOT_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr int ACC_COLS = 2 * BN;                               // main | cross-term accumulator
  constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t full_bar = smem + (uint32_t)NS * SLOT_BYTES;
  const uint32_t empty_bar = full_bar + 8u * NS;
  const uint32_t tfull_bar = empty_bar + 8u * NS;                // [2]
  const uint32_t tempty_bar = tfull_bar + 16u;                   // [2]
  const uint32_t tmem_ptr = tempty_bar + 16u;
  // CPA: neighbour-map double buffer [2][27 taps][128 rows] int32, filled one to two tiles ahead by the index warp (the
  // producers' own index loads sat on the critical path: 36 % of their stall samples waited for an nbr value)
  // bias of the tile's BN columns, staged once per tile by the epilogue warps (double buffer): the per-chunk __ldg of the
  // bias sat on the epilogue's critical path (15 % of the samples of the K = 128 layers waited for it)
  const uint32_t bias_s = tmem_ptr + 16u;                         // [2][128] floats
  const uint32_t ifull_bar = bias_s + 2u * 128u * 4u;             // [2]
  const uint32_t iempty_bar = ifull_bar + 16u;                    // [2]
  const uint32_t idx_buf = iempty_bar + 16u;
  constexpr uint32_t IDX_TILE_BYTES = TC_MAX_TAPS * TC_BM * 4;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int Mv = p.M;
  if (p.m_dev) { int md = *p.m_dev; Mv = md < Mv ? md : Mv; }
  const int n_tiles_n = p.cout / BN;
  const int m_tiles = MODE == FF3D_GEMM_CONV2D ? p.B * p.tiles_y * p.tiles_x : (Mv + TC_BM - 1) / TC_BM;
  const int total_tiles = m_tiles * n_tiles_n;
  if ((int)blockIdx.x >= total_tiles) return;   // uniform for the whole CTA, before any barrier / TMEM use

  const bool masked = MODE == FF3D_GEMM_SPARSE && p.tile_mask != nullptr;
  // unit = one tap (cin >= 64: cpt K-steps) or one packed K-step of tps taps (cin < 64); bit u of the mask = unit u is used
  // by at least one row of the tile
  auto unit_mask = [&](int tile) -> uint32_t {
    if (!masked) return 0u;
    const uint32_t tm = __ldg(p.tile_mask + tile / n_tiles_n);
    uint32_t um = tm;
    if (p.tps > 1) {
      um = 0;
      const uint32_t grp_bits = (1u << p.tps) - 1u;
      for (int u = 0; u < p.n_units; ++u)
        if ((tm >> (u * p.tps)) & grp_bits) um |= 1u << u;
    }
    return um ? um : 1u;
  };
  auto stage_count = [&](uint32_t um) -> int { return masked ? __popc(um) * p.cpt : p.n_stages; };

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(full_bar + 8u * s, CPA ? NPW * 32 + 1 : NPW); mbar_init(empty_bar + 8u * s, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar + 8u * i, 1); mbar_init(tempty_bar + 8u * i, 128); }
    if (CPA)
      for (int i = 0; i < 2; ++i) { mbar_init(ifull_bar + 8u * i, 32); mbar_init(iempty_bar + 8u * i, NPW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = (uint32_t)lds32(tmem_ptr);
  const uint32_t a_tx = MODE == FF3D_GEMM_CONV2D ? 2u * (uint32_t)(p.bw * p.bh) * 128u : 2u * A_BYTES;

  if (warp < NPW) {
    // =========================== producers: TMA issue only ===========================
    constexpr int RPW = TC_BM / NPW;                              // SPARSE: tile rows gathered by each producer warp
    Ring ring{0, 0u};
    const uint8_t* wbase = static_cast<const uint8_t*>(p.wimg);
    const size_t stage_bytes = 2 * (size_t)B_BYTES;
    if constexpr (CPA) {
      // ---- SPARSE gather by cp.async: lane (q, j) copies 16-byte chunk j of rows q, q+4, ... of this warp's RPW-row slice
      static_assert(MODE == FF3D_GEMM_SPARSE && RPW % 4 == 0, "cp.async gather is the SPARSE producer");
      constexpr int RI = RPW / 4;
      const int j = lane & 7, q = lane >> 3;
      const char* xbase = reinterpret_cast<const char*>(p.xs);
      const size_t row_bytes = (size_t)p.ldxs * 2;
      int it = 0;                                                  // tiles processed by this CTA (index buffer = it & 1)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int mtile = tile / n_tiles_n;
        const int ntile = tile - mtile * n_tiles_n;
        const uint8_t* wsrc = wbase + (size_t)ntile * p.n_stages * stage_bytes;
        uint32_t mm = masked ? unit_mask(tile) : (p.n_units >= 32 ? 0xFFFFFFFFu : ((1u << p.n_units) - 1u));
        // this lane's tap inside unit u and its channel offset: cin >= 64 -> tap u, the lane's 8 channels of chunk c;
        // cin < 64 -> the K-step packs tps taps: 16-byte chunk j belongs to tap u*tps + (8j / cin)
        const int lane_tap = p.tps > 1 ? (j * 8) / p.cin : 0;
        const int lane_col = p.tps > 1 ? (j * 8) % p.cin : j * 8;
        // the tile's neighbour map, staged by the index warp: [tap][row] (rows past the count and absent taps hold -1)
        const uint32_t ib = idx_buf + (uint32_t)(it & 1) * IDX_TILE_BYTES + (uint32_t)(warp * RPW + q) * 4u;
        const uint32_t tmk = masked ? __ldg(p.tile_mask + mtile) : 0xFFFFFFFFu;
        mbar_wait(ifull_bar + 8u * (it & 1), (uint32_t)((it >> 1) & 1));
        while (mm) {
          const int u = __ffs(mm) - 1;
          mm &= mm - 1u;
          const int t = u * p.tps + lane_tap;
          const bool tap_ok = t < p.taps && ((tmk >> t) & 1u);      // taps the index warp skipped hold stale rows
          int idx[RI];
#pragma unroll
          for (int i = 0; i < RI; ++i) idx[i] = tap_ok ? lds32(ib + (uint32_t)(t * TC_BM + 4 * i) * 4u) : -1;
          for (int c = 0; c < p.cpt; ++c) {
            const uint32_t slot_a = smem + (uint32_t)ring.slot * SLOT_BYTES;
            const uint32_t bar = full_bar + 8u * ring.slot;
            // every lane polls the slot's barrier itself (same word, same phase: the warp stays converged -- the
            // lane-0 wait + __syncwarp of the first version cost 8 % of the kernel's samples in branch resolution)
            mbar_wait(empty_bar + 8u * ring.slot, ring.phase ^ 1u);
            if (warp == 0 && lane == 0) {
              mbar_arrive_expect_tx(bar, 2 * B_BYTES);
              bulk_g2s(slot_a + 2 * A_BYTES, wsrc + (size_t)(u * p.cpt + c) * stage_bytes, 2 * B_BYTES, bar);
            }
#pragma unroll
            for (int plane = 0; plane < 2; ++plane) {
              const size_t col_bytes = (size_t)(plane * p.xs_lo + c * 64 + lane_col) * 2;
#pragma unroll
              for (int i = 0; i < RI; ++i) {
                const int row = warp * RPW + q + 4 * i;
                const int r = idx[i];
                const uint32_t dst = slot_a + (uint32_t)plane * A_BYTES + (uint32_t)(row * 128 + ((j ^ (row & 7)) << 4));
                cp_async16(dst, xbase + (r < 0 ? 0 : (size_t)r * row_bytes) + col_bytes, r < 0 ? 0u : 16u);
              }
            }
            // the hardware arrives on the stage's barrier once THIS thread's copies have landed (the CUTLASS sm100
            // cp.async + UMMA pipeline): exact signalling, nothing to wait for here
            cp_async_arrive_noinc(bar);
            if (++ring.slot == NS) { ring.slot = 0; ring.phase ^= 1u; }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(iempty_bar + 8u * (it & 1));     // this warp is done with the index buffer
      }
    } else
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mtile = tile / n_tiles_n;
      const int ntile = tile - mtile * n_tiles_n;
      const uint8_t* wsrc = wbase + (size_t)ntile * p.n_stages * stage_bytes;
      const int m0 = mtile * TC_BM;
      int cb = 0, y0 = 0, x0 = 0;
      if (MODE == FF3D_GEMM_CONV2D) {
        const int per_img = p.tiles_y * p.tiles_x;
        cb = mtile / per_img;
        const int r = mtile - cb * per_img;
        y0 = (r / p.tiles_x) * p.bh;
        x0 = (r - (r / p.tiles_x) * p.tiles_x) * p.bw;
      }
      // one pipeline stage: tap / unit `t` (weight image t*cpt + c), 64-channel chunk `c`
      auto issue = [&](int t, int c) {
        const uint32_t slot_a = smem + (uint32_t)ring.slot * SLOT_BYTES;
        const uint32_t bar = full_bar + 8u * ring.slot;
        if (lane == 0) {
          mbar_wait(empty_bar + 8u * ring.slot, ring.phase ^ 1u);
          // every producer warp announces its own bytes; warp 0 also owns the weight stage
          const uint32_t a_bytes = MODE == FF3D_GEMM_SPARSE ? 2u * RPW * 128u : a_tx;
          mbar_arrive_expect_tx(bar, a_bytes + (warp == 0 ? 2 * B_BYTES : 0u));
          if (warp == 0) bulk_g2s(slot_a + 2 * A_BYTES, wsrc + (size_t)(t * p.cpt + c) * stage_bytes, 2 * B_BYTES, bar);
          if (MODE == FF3D_GEMM_ROWS) {
            tma_load_2d(slot_a, &tmA, c * 64, m0, bar);
            tma_load_2d(slot_a + A_BYTES, &tmA, p.xs_lo + c * 64, m0, bar);
          } else if (MODE == FF3D_GEMM_CONV2D) {
            const int ky = t / p.kw, kx = t - ky * p.kw;
            const int ix0 = x0 * p.stride + kx - p.pad, iy0 = y0 * p.stride + ky - p.pad;    // first input pixel of the box
            tma_load_4d(slot_a, &tmA, c * 64, ix0, iy0, cb, bar);
            tma_load_4d(slot_a + A_BYTES, &tmA, p.xs_lo + c * 64, ix0, iy0, cb, bar);
          }
        }
        __syncwarp();
        if (MODE == FF3D_GEMM_SPARSE) {
          // this warp gathers tile rows [warp*RPW, warp*RPW + RPW): work item = (plane, group of four rows); absent
          // neighbours and rows past the count read the all-zero row
          constexpr int GROUPS = RPW / 4;
          for (int item = lane; item < 2 * GROUPS; item += 32) {
            const int plane = item / GROUPS, g = item - plane * GROUPS;
            const int row0 = warp * RPW + 4 * g;
            int r[4];
            const int* nb = p.nbr + (size_t)t * p.nbr_stride + m0 + row0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              int v = (m0 + row0 + i < Mv) ? __ldg(nb + i) : -1;
              r[i] = v < 0 ? p.zero_row : v;
            }
            tma_gather4(slot_a + (uint32_t)plane * A_BYTES + (uint32_t)row0 * 128u, &tmA, plane * p.xs_lo + c * 64, r[0], r[1],
                        r[2], r[3], bar);
          }
        }
        ring.advance(1, NS);
      };
      if (masked) {
        for (uint32_t mm = unit_mask(tile); mm; mm &= mm - 1u) {
          const int t = __ffs(mm) - 1;
          for (int c = 0; c < p.cpt; ++c) issue(t, c);
        }
      } else {
        for (int t = 0; t < p.taps; ++t)
          for (int c = 0; c < p.cpt; ++c) issue(t, c);
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc_wide = make_idesc<true>(2 * BN), idesc_cross = make_idesc<true>(BN);
      Ring ring{0, 0u};
      int gc = 0;                                        // accumulator hand-overs so far (tiles, or chunks when CHUNKED)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nst = stage_count(unit_mask(tile));
        uint32_t d_main = 0, d_cross = 0;
        int ab = 0;
        for (int s = 0; s < nst; ++s) {
          const bool cut = CHUNKED && nst > CHUNK_MIN;
          const int sc = cut ? s % CHUNK : s;            // stage inside the current accumulation run
          if (sc == 0) {
            // the epilogue that last read this accumulator buffer (two hand-overs ago) has drained it
            ab = gc & 1;
            mbar_wait(tempty_bar + 8u * ab, (uint32_t)(((gc >> 1) & 1) ^ 1));
            tc_fence_after();
            d_main = tmem_base + (uint32_t)(ab * ACC_COLS);
            d_cross = d_main + BN;
          }
          mbar_wait(full_bar + 8u * ring.slot, ring.phase);
          tc_fence_after();
          const uint32_t a_hi = smem + (uint32_t)ring.slot * SLOT_BYTES, a_lo = a_hi + A_BYTES;
          const uint32_t b_hi = a_hi + 2 * A_BYTES;
          static_assert(B_BYTES % 1024 == 0, "B_lo must continue B_hi's 8-row swizzle atoms");
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 4 x (K = 16 halves = 32 bytes) per 128-byte swizzled row
            const uint64_t dbh = make_desc(b_hi + k * 32);   // as an N = 2*BN operand it runs on into B_lo
            umma<true>(d_main, make_desc(a_hi + k * 32), dbh, idesc_wide, (sc | k) ? 1u : 0u);   // main += A_hi*B_hi ; cross += A_hi*B_lo
            umma<true>(d_cross, make_desc(a_lo + k * 32), dbh, idesc_cross, 1u);                 // cross += A_lo*B_hi
          }
          umma_commit(empty_bar + 8u * ring.slot);
          ring.advance(1, NS);
          if (s == nst - 1 || (cut && sc == CHUNK - 1)) {
            umma_commit(tfull_bar + 8u * ab);            // this run's accumulator is complete
            ++gc;
          }
        }
      }
    }
    __syncwarp();
  } else if (CPA && warp == IDX_WARP) {
    // =========================== index warp (CPA): neighbour maps of the coming tiles -> shared memory ===========================
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int mtile = tile / n_tiles_n;
      const int m0 = mtile * TC_BM;
      const uint32_t tm = masked ? __ldg(p.tile_mask + mtile) : 0xFFFFFFFFu;      // tap-level mask of the tile
      if (it >= 2) mbar_wait(iempty_bar + 8u * (it & 1), (uint32_t)(((it >> 1) - 1) & 1));
      const uint32_t ib = idx_buf + (uint32_t)(it & 1) * IDX_TILE_BYTES;
      // lane l: rows l, l+32, l+64, l+96 of every present tap, as 4-byte cp.async copies (128-byte coalesced per warp
      // instruction, no registers, all ~100 copies of the tile in flight at once); rows past the count are zero-filled:
      // they gather row 0 and their outputs are never stored
      for (int t = 0; t < p.taps; ++t) {
        if (!((tm >> t) & 1u)) continue;
        const int* nb = p.nbr + (size_t)t * p.nbr_stride + m0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool ok = m0 + lane + 32 * i < Mv;
          cp_async4(ib + (uint32_t)(t * TC_BM + lane + 32 * i) * 4u, ok ? nb + lane + 32 * i : p.nbr, ok ? 4u : 0u);
        }
      }
      cp_async_arrive_noinc(ifull_bar + 8u * (it & 1));              // 32 arrivals, each when that lane's copies have landed
    }
  } else {
    // =========================== epilogue (four warps: TMEM lane quarter = warp % 4) ===========================
    const int r = (warp & 3) * 32 + lane;                     // tile row <-> TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    bool ovf = false;
    int gc = 0;                                          // accumulator hand-overs consumed so far
    // branch-free epilogue arithmetic: activation as a clamp, the residual added before and / or after it through 0 / 1
    // multipliers (an absent bias or residual is a zero addend)
    const float act_lo = p.act == FF3D_ACT_NONE ? -INFINITY : 0.f, act_hi = p.act == FF3D_ACT_RELU6 ? 6.f : INFINITY;
    const float m_pre = p.res_after_act ? 0.f : 1.f, m_post = 1.f - m_pre;
    const int et = (warp & 3) * 32 + lane;               // epilogue thread index 0..127
    int eit = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++eit) {
      const int mtile = tile / n_tiles_n;
      const int n0 = (tile - mtile * n_tiles_n) * BN;
      // stage the bias of this tile's columns (one value per thread); one named barrier per tile among the four epilogue
      // warps (two buffers: a warp is never more than one tile ahead of another)
      const uint32_t bsm = bias_s + (uint32_t)(eit & 1) * 512u;
      if (et < BN) sts32(bsm + (uint32_t)et * 4u, p.bias ? __float_as_int(__ldg(p.bias + n0 + et)) : 0);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int nst = stage_count(unit_mask(tile));
      const int n_runs = (CHUNKED && nst > CHUNK_MIN) ? (nst + CHUNK - 1) / CHUNK : 1;
      // output row of this thread
      bool rvalid;
      long long row = 0, rrow = 0;
      if (MODE == FF3D_GEMM_CONV2D) {
        const int per_img = p.tiles_y * p.tiles_x;
        const int cb = mtile / per_img;
        const int q = mtile - cb * per_img;
        const int dy = r / p.bw, dx = r - dy * p.bw;
        const int oy = (q / p.tiles_x) * p.bh + dy, ox = (q - (q / p.tiles_x) * p.tiles_x) * p.bw + dx;
        rvalid = r < p.bw * p.bh && oy < p.Ho && ox < p.Wo;
        row = cb * p.y_bstride + p.y_row0 + (long long)(oy * p.uy + p.dy) * (p.Wo * p.ux) + ox * p.ux + p.dx;
        rrow = cb * p.res_bstride + (long long)oy * p.Wo + ox;
      } else {
        const int m = mtile * TC_BM + r;
        rvalid = m < Mv;
        row = rrow = m;
        if (MODE == FF3D_GEMM_SPARSE && rvalid && p.y_row) row = __ldg(p.y_row + m);
      }
      float* yp = nullptr;
      if (rvalid && p.y) {
        if (MODE == FF3D_GEMM_SPARSE && p.y_off) yp = p.y + __ldg(p.y_off + (int)rrow);
        else yp = p.y + row * p.ldy;
      }
      __half* ysp = (rvalid && p.ys) ? p.ys + row * p.ldys : nullptr;
      // running sums of the chunks (CHUNKED): one fp32 register per output column of this thread's row
      float racc[CHUNKED ? BN : 1];
      if constexpr (CHUNKED) {
        for (int run = 0; run < n_runs - 1; ++run) {
          const int abr = gc & 1;
          mbar_wait(tfull_bar + 8u * abr, (uint32_t)((gc >> 1) & 1));
          tc_fence_after();
          const uint32_t accr = lane_base + (uint32_t)(abr * ACC_COLS);
#pragma unroll
          for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t v[16], v2[16];
            tmem_ld16_nowait(accr + (uint32_t)c0, v);            // main and cross in flight together, one wait
            tmem_ld16_nowait(accr + (uint32_t)(BN + c0), v2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float t = fmaf(__uint_as_float(v2[i]), 1.f / 2048.f, __uint_as_float(v[i]));
              racc[c0 + i] = run == 0 ? t : racc[c0 + i] + t;
            }
          }
          tc_fence_before();
          mbar_arrive(tempty_bar + 8u * abr);
          ++gc;
        }
      }
      const int ab = gc & 1;
      mbar_wait(tfull_bar + 8u * ab, (uint32_t)((gc >> 1) & 1));
      tc_fence_after();
      ++gc;
      const uint32_t acc = lane_base + (uint32_t)(ab * ACC_COLS);
#pragma unroll (CHUNKED ? BN / 16 : 1)
      for (int c0 = 0; c0 < BN; c0 += 16) {
        const int n = n0 + c0;
        float rs[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) rs[i] = 0.f;
        const bool has_res = (p.res != nullptr || p.res_s != nullptr);
        if (has_res && rvalid) {
          if (p.res_s) {
            // residual stored in split form: value = hi + lo / 2048
            const __half* hp = p.res_s + rrow * p.ldres_s + n;
            const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(hp)), h1 = __ldg(reinterpret_cast<const uint4*>(hp + 8));
            const uint4 l0 = __ldg(reinterpret_cast<const uint4*>(hp + p.res_s_lo)), l1 = __ldg(reinterpret_cast<const uint4*>(hp + p.res_s_lo + 8));
            const __half2* hh0 = reinterpret_cast<const __half2*>(&h0);
            const __half2* hh1 = reinterpret_cast<const __half2*>(&h1);
            const __half2* ll0 = reinterpret_cast<const __half2*>(&l0);
            const __half2* ll1 = reinterpret_cast<const __half2*>(&l1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 a = __half22float2(hh0[i]), b = __half22float2(ll0[i]);
              rs[2 * i] = fmaf(b.x, 1.f / 2048.f, a.x); rs[2 * i + 1] = fmaf(b.y, 1.f / 2048.f, a.y);
              const float2 c = __half22float2(hh1[i]), d = __half22float2(ll1[i]);
              rs[8 + 2 * i] = fmaf(d.x, 1.f / 2048.f, c.x); rs[8 + 2 * i + 1] = fmaf(d.y, 1.f / 2048.f, c.y);
            }
          } else {
            const float* rp = p.res + rrow * p.ldres + n;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(rp + i));
              rs[i] = t.x; rs[i + 1] = t.y; rs[i + 2] = t.z; rs[i + 3] = t.w;
            }
          }
        }
        float v[16];
        {
          uint32_t ua[16], ub[16];
          tmem_ld16_nowait(acc + (uint32_t)c0, ua);          // warp-collective: all lanes execute; one wait for both
          tmem_ld16_nowait(acc + (uint32_t)(BN + c0), ub);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaf(__uint_as_float(ub[i]), 1.f / 2048.f, __uint_as_float(ua[i]));   // cross terms: 2^11 scale
        }
        if (rvalid) {
#pragma unroll
          for (int i4 = 0; i4 < 16; i4 += 4) {
            const int4 b4 = lds128i(bsm + (uint32_t)(c0 + i4) * 4u);
            const float bb[4] = {__int_as_float(b4.x), __int_as_float(b4.y), __int_as_float(b4.z), __int_as_float(b4.w)};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = i4 + j;
              float a = v[i];
              if constexpr (CHUNKED) { if (n_runs > 1) a += racc[c0 + i]; }
              a += bb[j];
              a = fmaf(m_pre, rs[i], a);
              a = fminf(fmaxf(a, act_lo), act_hi);
              v[i] = fmaf(m_post, rs[i], a);
            }
          }
          if (yp) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(yp + n + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
          if (ysp) {
            uint32_t hw[8], lw[8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              split_f16x4(make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), hw[2 * i], hw[2 * i + 1], lw[2 * i],
                          lw[2 * i + 1], ovf);
            uint4* dh = reinterpret_cast<uint4*>(ysp + n);
            uint4* dl = reinterpret_cast<uint4*>(ysp + p.ys_lo + n);
            dh[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]); dh[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
            dl[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]); dl[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar + 8u * ab);                     // 128 arrivals free the accumulator buffer
    }
    if (ovf && p.overflow) atomicOr(p.overflow, 1);
  }
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// fp32 rows -> split rows [hi(C) | lo(C)] (for activations produced by non-GEMM kernels); 8 channels per thread
__global__ void split_rows_kernel(const float* __restrict__ x, int ldx, const int* __restrict__ n_dev, long long rows, int C,
                                  __half* __restrict__ ys, int ldys, int ys_lo, int* overflow) {
  long long n = rows;
  if (n_dev) { const long long nd = *n_dev; n = nd < n ? nd : n; }
  const int c8 = C / 8;
  const long long total = n * c8;
  bool ovf = false;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / c8;
    const int c = (int)(e - row * c8) * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + row * ldx + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(x + row * ldx + c + 4));
    uint32_t hw[4], lw[4];
    split_f16x4(a, hw[0], hw[1], lw[0], lw[1], ovf);
    split_f16x4(b, hw[2], hw[3], lw[2], lw[3], ovf);
    *reinterpret_cast<uint4*>(ys + row * ldys + c) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(ys + row * ldys + ys_lo + c) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
  if (ovf && overflow) atomicOr(overflow, 1);
}

// split rows -> fp32 rows (value = hi + lo / 2048): for consumers that are not TMA GEMMs, and for tests
__global__ void unsplit_rows_kernel(const __half* __restrict__ xs, int ldxs, int xs_lo, const int* __restrict__ n_dev,
                                    long long rows, int C, float* __restrict__ y, int ldy) {
  long long n = rows;
  if (n_dev) { const long long nd = *n_dev; n = nd < n ? nd : n; }
  const int c8 = C / 8;
  const long long total = n * c8;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / c8;
    const int c = (int)(e - row * c8) * 8;
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(xs + row * ldxs + c));
    const uint4 l = __ldg(reinterpret_cast<const uint4*>(xs + row * ldxs + xs_lo + c));
    const __half2* hh = reinterpret_cast<const __half2*>(&h);
    const __half2* ll = reinterpret_cast<const __half2*>(&l);
    float o[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 a = __half22float2(hh[i]), b = __half22float2(ll[i]);
      o[2 * i] = fmaf(b.x, 1.f / 2048.f, a.x);
      o[2 * i + 1] = fmaf(b.y, 1.f / 2048.f, a.y);
    }
    *reinterpret_cast<float4*>(y + row * ldy + c) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(y + row * ldy + c + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
}

// ---- host side -------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled encode_tiled() {
  // resolved once through the runtime (no link-time dependency on libcuda); thread-safe static initialiser
  static const PFN_encodeTiled fn = []() -> PFN_encodeTiled {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<PFN_encodeTiled>(ptr);
  }();
  return fn;
}

// tensor map over split activation rows: rank 2 = (halves of a row, rows), rank 4 = (halves, W, H, B)
static int make_map(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box, int pixel_stride = 1) {
  PFN_encodeTiled enc = encode_tiled();
  if (!enc) { set_error("ff3d_tmagemm: cuTensorMapEncodeTiled is not available from this driver"); return FF3D_ECUDA; }
  // strided conv: the W / H dimensions are traversed with the conv stride (every stride-th pixel of the box is copied)
  cuuint32_t estr[4] = {1, (cuuint32_t)pixel_stride, (cuuint32_t)pixel_stride, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("ff3d_tmagemm: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return FF3D_ECUDA; }
  return FF3D_OK;
}

template <int MODE, int BN, int NPW, bool CPA>
static int launch_tm_cfg(const CUtensorMap& tm, const TmP& p, long long m_tiles, cudaStream_t st) {
  // 64 / 48 / 40 / 36 KB per stage; the cp.async gather also holds two neighbour-map tiles (27 KB)
  constexpr int NS = BN == 128 ? 3 : (BN == 64 ? 4 : ((CPA && BN == 32) ? 4 : 5));
  constexpr size_t SLOT_BYTES = 2 * (size_t)TC_BM * 128 + 2 * (size_t)BN * 128;
  constexpr size_t IDX_BYTES = 2 * 128 * 4 + (CPA ? 2 * (size_t)TC_MAX_TAPS * TC_BM * 4 + 32 : 0);   // bias staging (+ index maps)
  const size_t smem = NS * SLOT_BYTES + (2 * NS + 4) * sizeof(uint64_t) + 32 + IDX_BYTES + 1024;
  static_assert(NS * SLOT_BYTES + (2 * NS + 4) * sizeof(uint64_t) + 32 + IDX_BYTES + 1024 <= 227 * 1024, "shared memory budget");
  static const cudaError_t attr =
      cudaFuncSetAttribute(tmagemm_kernel<MODE, BN, NS, NPW, CPA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (attr != cudaSuccess) { set_error("ff3d_tmagemm: cudaFuncSetAttribute: %s", cudaGetErrorString(attr)); return FF3D_ECUDA; }
  const long long tiles = m_tiles * (p.cout / BN);
  const long long resident = num_sms();
  dim3 grid((unsigned)(tiles < resident ? tiles : resident));
  tmagemm_kernel<MODE, BN, NS, NPW, CPA><<<grid, 32 * (NPW + 5 + (CPA ? 1 : 0)), smem, st>>>(tm, p);
  return check_launch("ff3d_tmagemm");
}

// SPARSE gather engine: FF3D_SPARSE_GATHER = "cpasync" (default: LDGSTS copies by 8 producer warps) or "tma" (tile::gather4
// issued by FF3D_TMA_NPW = 1 / 4 / 8 producer warps); read once
static int sparse_gather_cfg() {
  static const int v = []() {
    const char* g = getenv("FF3D_SPARSE_GATHER");
    if (!g || strcmp(g, "tma") != 0) return 0;                     // cp.async
    const char* e = getenv("FF3D_TMA_NPW");
    const int n = e ? atoi(e) : 8;
    return (n == 1 || n == 4 || n == 8) ? n : 8;
  }();
  return v;
}

template <int MODE, int BN>
static int launch_tm(const CUtensorMap& tm, const TmP& p, long long m_tiles, cudaStream_t st) {
  if constexpr (MODE == FF3D_GEMM_SPARSE) {
    switch (p.tps > 1 ? 0 : sparse_gather_cfg()) {
      case 0: return launch_tm_cfg<MODE, BN, 8, true>(tm, p, m_tiles, st);      // 8 producer warps (4 measured 20 % slower)
      case 1: return launch_tm_cfg<MODE, BN, 1, false>(tm, p, m_tiles, st);
      case 4: return launch_tm_cfg<MODE, BN, 4, false>(tm, p, m_tiles, st);
      default: return launch_tm_cfg<MODE, BN, 8, false>(tm, p, m_tiles, st);
    }
  } else {
    return launch_tm_cfg<MODE, BN, 1, false>(tm, p, m_tiles, st);
  }
}

template <int MODE>
static int launch_tm_bn(const CUtensorMap& tm, const TmP& p, long long m_tiles, int bn, cudaStream_t st) {
  if (bn == 128) return launch_tm<MODE, 128>(tm, p, m_tiles, st);
  if (bn == 64) return launch_tm<MODE, 64>(tm, p, m_tiles, st);
  if constexpr (MODE == FF3D_GEMM_SPARSE) {
    // narrow levels of the sparse encoder (C = 16 / 32): cp.async gather only
    if (bn == 32) return launch_tm_cfg<MODE, 32, 8, true>(tm, p, m_tiles, st);
    if (bn == 16) return launch_tm_cfg<MODE, 16, 8, true>(tm, p, m_tiles, st);
  } else {
    // narrow dense outputs (e.g. the 128 -> 10(16) heat-map conv)
    if (bn == 32) return launch_tm_cfg<MODE, 32, 1, false>(tm, p, m_tiles, st);
    if (bn == 16) return launch_tm_cfg<MODE, 16, 1, false>(tm, p, m_tiles, st);
  }
  set_error("ff3d_tmagemm: unsupported N tile %d", bn);
  return FF3D_EINVAL;
}

}  // namespace ff3d

// output patch (bw x bh <= 128 pixels) of the CONV2D tiles: least padded pixels, then the widest patch
extern "C" void ff3d_tmagemm_conv_patch(int Ho, int Wo, int* bw_out, int* bh_out) {
  long long best = -1;
  int bbw = 16, bbh = 8;
  for (int bw = 4; bw <= 128; ++bw) {
    const int bh_max = 128 / bw;
    for (int bh = 1; bh <= bh_max; ++bh) {
      if (bw * bh < 96) continue;                                      // keep the 128-row MMA at least 3/4 full
      const long long tiles = (long long)((Wo + bw - 1) / bw) * ((Ho + bh - 1) / bh);
      const long long cost = tiles * 128;                              // every tile costs a full 128-row MMA
      if (best < 0 || cost < best || (cost == best && bw > bbw)) { best = cost; bbw = bw; bbh = bh; }
    }
  }
  *bw_out = bbw;
  *bh_out = bbh;
}

extern "C" int ff3d_tmagemm_supported(const ff3d_gemm_desc* d) {
  if (!d || !d->xs) return 0;
  if (d->x2) return 0;
  const bool wide_in = d->cin >= 64 && d->cin % 64 == 0;
  const bool wide_out = d->cout % 128 == 0 || d->cout == 64;
  const bool narrow_out = d->cout == 16 || d->cout == 32;
  if (d->mode == FF3D_GEMM_SPARSE) {
    // cp.async gather: also the narrow levels (cin 8 / 16 / 32 pack 8 / 4 / 2 taps into one K-step; cout 16 / 32)
    if (d->taps > ff3d::TC_MAX_TAPS) return 0;
    if (!(wide_in || d->cin == 8 || d->cin == 16 || d->cin == 32)) return 0;
    return (wide_out || narrow_out) ? 1 : 0;
  }
  if (!wide_in || !(wide_out || narrow_out)) return 0;
  if (d->mode == FF3D_GEMM_CONV2D && (d->stride < 1 || d->stride > 2 || (d->stride != 1 && (d->ux > 1 || d->uy > 1)))) return 0;
  return 1;
}

extern "C" int ff3d_tmagemm(const ff3d_gemm_desc* d, const void* wimg16, int bn, int* overflow_dev, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(d != nullptr && wimg16 != nullptr, "ff3d_tmagemm: null argument");
  FF3D_REQUIRE(ff3d_tmagemm_supported(d), "ff3d_tmagemm: unsupported layer (needs split A rows, cin %% 64 == 0, cout 64 or a "
                                          "multiple of 128, conv stride 1; got cin=%d cout=%d mode=%d)", d->cin, d->cout, d->mode);
  if (bn == 0) bn = d->cout % 128 == 0 ? 128 : d->cout;
  FF3D_REQUIRE((bn == 16 || bn == 32 || bn == 64 || bn == 128) && d->cout % bn == 0,
               "ff3d_tmagemm: N tile %d does not divide cout=%d", bn, d->cout);
  FF3D_REQUIRE(d->y != nullptr || d->ys != nullptr, "ff3d_tmagemm: no output");
  FF3D_REQUIRE((reinterpret_cast<uintptr_t>(d->xs) & 15) == 0 && d->ldxs % 8 == 0 && d->xs_lo % 8 == 0,
               "ff3d_tmagemm: split A rows must be 16-byte aligned (base, row stride, plane offset)");
  FF3D_REQUIRE((reinterpret_cast<uintptr_t>(wimg16) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0,
               "ff3d_tmagemm: wimg and bias must be 16-byte aligned");
  FF3D_REQUIRE(!d->ys || ((reinterpret_cast<uintptr_t>(d->ys) & 15) == 0 && d->ldys % 8 == 0 && d->ys_lo % 8 == 0),
               "ff3d_tmagemm: split output rows must be 16-byte aligned");
  FF3D_REQUIRE(!d->res_s || ((reinterpret_cast<uintptr_t>(d->res_s) & 15) == 0 && d->ldres_s % 8 == 0 && d->res_s_lo % 8 == 0),
               "ff3d_tmagemm: split residual rows must be 16-byte aligned");
  FF3D_REQUIRE(!d->y || ((reinterpret_cast<uintptr_t>(d->y) & 15) == 0 && (d->ldy % 4 == 0 || d->y_off)),
               "ff3d_tmagemm: fp32 output rows must be 16-byte aligned");
  FF3D_REQUIRE(!d->res || ((reinterpret_cast<uintptr_t>(d->res) & 15) == 0 && d->ldres % 4 == 0),
               "ff3d_tmagemm: fp32 residual rows must be 16-byte aligned");
  if (d->M <= 0) return FF3D_OK;
  TmP p = {};
  p.mode = d->mode; p.M = d->M; p.m_dev = d->m_dev;
  p.cin = d->cin; p.cout = d->cout; p.taps = d->taps; p.xs_lo = d->xs_lo;
  p.xs = static_cast<const __half*>(d->xs); p.ldxs = d->ldxs;
  p.wimg = wimg16; p.bias = d->bias;
  p.res = d->res; p.ldres = d->ldres;
  p.res_s = static_cast<const __half*>(d->res_s); p.ldres_s = d->ldres_s; p.res_s_lo = d->res_s_lo;
  p.y = d->y; p.ldy = d->ldy;
  p.ys = static_cast<__half*>(d->ys); p.ldys = d->ldys; p.ys_lo = d->ys_lo;
  p.act = d->act; p.res_after_act = d->res_after_act;
  p.nbr = d->nbr; p.nbr_stride = d->nbr_stride; p.y_off = d->y_off; p.y_row = d->y_row;
  p.tile_mask = d->mode == FF3D_GEMM_SPARSE ? d->tile_mask : nullptr;
  p.zero_row = d->zero_row;
  p.overflow = overflow_dev;
  p.tps = d->cin >= 64 ? 1 : 64 / d->cin;
  p.cpt = d->cin >= 64 ? d->cin / 64 : 1;
  p.n_units = d->cin >= 64 ? d->taps : (d->taps + p.tps - 1) / p.tps;
  p.n_stages = p.n_units * p.cpt;
  cudaStream_t st = as_stream(stream);
  CUtensorMap tm;
  const cuuint64_t row_bytes = (cuuint64_t)d->ldxs * 2;
  if (d->mode == FF3D_GEMM_ROWS) {
    FF3D_REQUIRE(d->taps == 1, "ff3d_tmagemm: ROWS mode has one tap");
    cuuint64_t dims[2] = {(cuuint64_t)(d->xs_lo + d->cin), (cuuint64_t)d->M};
    cuuint64_t str[1] = {row_bytes};
    cuuint32_t box[2] = {64, 128};
    int rc = make_map(&tm, d->xs, 2, dims, str, box);
    if (rc) return rc;
    return launch_tm_bn<FF3D_GEMM_ROWS>(tm, p, cdiv(d->M, TC_BM), bn, st);
  }
  if (d->mode == FF3D_GEMM_SPARSE) {
    FF3D_REQUIRE(d->nbr != nullptr && d->nbr_stride >= d->M, "ff3d_tmagemm: sparse mode needs nbr [taps, >= M]");
    FF3D_REQUIRE(d->xs_rows > 0 && (sparse_gather_cfg() == 0 || (d->zero_row >= 0 && d->xs_rows > d->zero_row)),
                 "ff3d_tmagemm: the gather4 engine needs an all-zero row inside xs");
    FF3D_REQUIRE(!d->y_off || d->y, "ff3d_tmagemm: y_off addresses the fp32 output");
    memset(&tm, 0, sizeof(tm));
    if (p.tps == 1 && sparse_gather_cfg() != 0) {                      // the tensor map is only read by the gather4 engine
      cuuint64_t dims[2] = {(cuuint64_t)(d->xs_lo + d->cin), (cuuint64_t)d->xs_rows};
      cuuint64_t str[1] = {row_bytes};
      cuuint32_t box[2] = {64, 1};                                     // gather4: one row per box, four boxes per instruction
      int rc = make_map(&tm, d->xs, 2, dims, str, box);
      if (rc) return rc;
    }
    return launch_tm_bn<FF3D_GEMM_SPARSE>(tm, p, cdiv(d->M, TC_BM), bn, st);
  }
  FF3D_REQUIRE(d->mode == FF3D_GEMM_CONV2D && d->taps == d->kh * d->kw && (long long)d->B * d->Ho * d->Wo == d->M,
               "ff3d_tmagemm: bad conv geometry");
  FF3D_REQUIRE(d->Ho == (d->H + 2 * d->pad - d->kh) / d->stride + 1 && d->Wo == (d->W + 2 * d->pad - d->kw) / d->stride + 1,
               "ff3d_tmagemm: conv output geometry");
  p.stride = d->stride;
  p.B = d->B; p.Ho = d->Ho; p.Wo = d->Wo; p.kw = d->kw; p.pad = d->pad;
  ff3d_tmagemm_conv_patch(d->Ho, d->Wo, &p.bw, &p.bh);
  p.tiles_x = cdiv(d->Wo, p.bw); p.tiles_y = cdiv(d->Ho, p.bh);
  const long long xbs = d->x_bstride ? d->x_bstride : (long long)d->H * d->W;
  p.y_bstride = d->y_bstride ? d->y_bstride : (long long)d->Ho * d->Wo;
  p.res_bstride = (long long)d->Ho * d->Wo;
  p.ux = d->ux > 0 ? d->ux : 1; p.uy = d->uy > 0 ? d->uy : 1; p.dx = d->dx; p.dy = d->dy; p.y_row0 = d->y_row0;
  FF3D_REQUIRE((p.ux == 1 && p.uy == 1) || (!d->res && !d->res_s), "ff3d_tmagemm: no residual on an up-sampling lattice");
  if (!d->y_bstride) p.y_bstride = (long long)d->Ho * p.uy * d->Wo * p.ux;
  {
    cuuint64_t dims[4] = {(cuuint64_t)(d->xs_lo + d->cin), (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t str[3] = {row_bytes, row_bytes * (cuuint64_t)d->W, row_bytes * (cuuint64_t)xbs};
    cuuint32_t box[4] = {64, (cuuint32_t)((p.bw - 1) * d->stride + 1), (cuuint32_t)((p.bh - 1) * d->stride + 1), 1};
    FF3D_REQUIRE(box[1] <= 256 && box[2] <= 256, "ff3d_tmagemm: strided box exceeds the TMA limit");
    int rc = make_map(&tm, d->xs, 4, dims, str, box, d->stride);
    if (rc) return rc;
  }
  return launch_tm_bn<FF3D_GEMM_CONV2D>(tm, p, (long long)d->B * p.tiles_y * p.tiles_x, bn, st);
}

extern "C" int ff3d_split_rows(const float* x, int ldx, const int* n_dev, long long rows, int C, void* ys, int ldys, int ys_lo,
                               int* overflow_dev, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(C % 8 == 0 && ldx % 4 == 0 && ldys % 8 == 0 && ys_lo % 8 == 0, "ff3d_split_rows: C %% 8, 16-byte aligned rows");
  if (rows <= 0) return FF3D_OK;
  long long nb = (rows * (C / 8) + 255) / 256, cap = (long long)num_sms() * 16;
  split_rows_kernel<<<(int)(nb > cap ? cap : nb), 256, 0, as_stream(stream)>>>(x, ldx, n_dev, rows, C, static_cast<__half*>(ys), ldys,
                                                                            ys_lo, overflow_dev);
  return check_launch("ff3d_split_rows");
}

extern "C" int ff3d_unsplit_rows(const void* xs, int ldxs, int xs_lo, const int* n_dev, long long rows, int C, float* y, int ldy,
                                 ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(C % 8 == 0 && ldy % 4 == 0 && ldxs % 8 == 0 && xs_lo % 8 == 0, "ff3d_unsplit_rows: C %% 8, 16-byte aligned rows");
  if (rows <= 0) return FF3D_OK;
  long long nb = (rows * (C / 8) + 255) / 256, cap = (long long)num_sms() * 16;
  unsplit_rows_kernel<<<(int)(nb > cap ? cap : nb), 256, 0, as_stream(stream)>>>(static_cast<const __half*>(xs), ldxs, xs_lo, n_dev,
                                                                              rows, C, y, ldy);
  return check_launch("ff3d_unsplit_rows");
}
