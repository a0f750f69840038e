// tcgen05 implicit-GEMM with 3xTF32 split accumulation (fp32-grade accuracy on the 5th-gen tensor cores).
//
//   y[row(m), n] = act( sum_{t,c} x[src(m,t), c] * w[t,c,n] + bias[n] + res[m,n] )        (same contract as igemm.cu)
//
// B200 has no fp32 tensor-core MMA; kind::tf32 keeps 10 mantissa bits.  The parity bar of this path (heatmaps and box
// regressions within 1e-3 of an fp32 oracle, bit-exact top-k) does not survive ~40 stacked TF32 layers, so each operand
// is split a = a_hi + a_lo (a_hi = rn_tf32(a), a_lo = rn_tf32(a - a_hi): a_hi + a_lo reproduces a to ~2^-24 relative) and D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi accumulates in fp32 in TMEM.
//
// The two small cross terms go to their own TMEM accumulator: the tensor core's fp32 accumulate truncates, so every
// accumulation step costs up to 1 ulp of the running sum; keeping the 2/3 of the steps that carry 2^-11-sized terms
// out of the main accumulator cuts that bias 3x for free (measured: K=1024 error 3.1e-5 -> see tests).
//
// Persistent CTAs (one per SM for BN = 128, two for BN <= 64) loop over (128*MT) x BN output tiles; 128*G + 64 threads:
//   warps 0 .. 4G-1  G A-producer groups of 128 threads (G = number of smem slots; group g owns slot g and the pipeline
//              stages g, g+G, ...).  One gather UNIT = the 32 K-values (128 B: one tap x 32 channels, or 32/cin taps when
//              cin < 32) of a 128-row sub-tile: a quarter-warp reads one row (full 128-byte line per request), 8 LDG.128
//              per thread.  Units are double-buffered in registers (ping-pong, BN = 128): the next unit's loads are in
//              flight while this one is split into hi / lo TF32 and stored (STS.128) into the 128B-swizzled K-major smem
//              tiles -> fence.proxy.async -> mbarrier arrive.  The same warps run the epilogue of the PREVIOUS tile after
//              issuing this tile's gathers (double-buffered TMEM accumulators): tcgen05.ld 32x32b (thread <-> TMEM lane
//              <-> tile row), bias / residual / activation, 64-byte row segments to global memory.
//   warp 4G    B producer: one cp.async.bulk (UBLKCP) per stage of the host-pre-swizzled [hi | lo] weight image.
//   warp 4G+1  TMEM alloc/dealloc + single-thread tcgen05.mma issue (per K=8 step: one 128 x 2BN and one 128 x BN MMA),
//              tcgen05.commit releases smem stages / signals the epilogue.
// All shared-memory accesses use 32-bit shared-space addresses (LDS / STS / mbarrier on shared::cta), never generic
// pointers.  Sparse mode stages the tile's whole neighbour map (taps x rows int32), conv mode one int4 of row geometry per
// row, in shared memory up front, so the gather's dependent index load is off the per-stage critical path.
//
// Round 2:
//  * F16 = true: fp16 hi/lo operand split instead of TF32 hi/lo:  a = hi + 2^-11 lo, hi = rn_f16(a), lo = rn_f16((a-hi) 2^11)
//    (saturating; a device flag is raised beyond +-65504).  Same 22-bit products, but kind::f16 MMAs run at twice the TF32
//    rate and a 128-byte shared-memory row covers 64 K-values instead of 32: half the tensor time AND half the
//    shared-memory bytes per product (the TF32 kernel's limiter).  y = D_main + 2^-11 D_cross.
//  * Sparse mode takes a per-128-row-tile tap mask (ff3d_sp_nbr_build): rows are mask-sorted by the rulebook builder, and
//    all three warp roles walk only the taps (cin >= K-step) / multi-tap stages (cin < K-step) whose bit is set, so the
//    (row, tap) pairs that exist in no row of the tile are neither gathered, split, copied nor multiplied.
//  * Row offsets are 64-bit (row index * ld), the one-time kernel attribute setup is a thread-safe static initialiser.
#include "tc_common.cuh"

namespace ff3d {

struct TcP {
  int mode, M;
  const int* m_dev;
  int cin, cout, taps;
  const float* x; int ldx;
  const float* x2;
  const void* wimg;         // [n_tiles][n_stages][2][BN*128 bytes] pre-swizzled hi/lo images (tf32 words or fp16 halves)
  const float* bias;
  const float* res; int ldres;
  float* y; int ldy;
  int act, res_after_act;
  int B, H, W, Ho, Wo, kh, kw, stride, pad;
  long long x_bstride, y_bstride, y_row0;
  int ux, uy, dx, dy;
  const int* nbr; int nbr_stride;
  const int* y_off;
  const uint32_t* tile_mask;   // SPARSE, optional: per 128-row tile the OR of its rows' tap masks
  const int* y_row;            // SPARSE, optional: output row of tile position m
  __half* ys; int ldys, ys_lo; // optional split (fp16 hi | lo) copy of the output rows for a TMA-fed consumer (tmagemm.cu)
  int* overflow;            // F16: device flag raised when an activation saturates the fp16 range
  int n_stages, tps, cpt;   // pipeline K-steps; taps per stage (cin < KS); KS-channel chunks per tap (cin >= KS)
  int n_units;              // skippable units of a tile: taps (cin >= KS) or multi-tap stages (cin < KS)
};

template <int N>
__device__ __forceinline__ void producer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }

// G = number of 128-thread A-producer groups = number of smem slots: 2 (and two CTAs per SM) for BN <= 64, 3 for
// BN = 128 (one CTA per SM).  n_slots == G makes every group the sole owner of one slot, so a producer is never more
// than one mbarrier phase ahead of the MMA issuer (the parity wait cannot tell phases two apart).
// MT = 128-row sub-tiles per CTA tile.  MT = 2 (BN = 128, large M) halves the weight bytes streamed from L2 per
// flop -- with the hi/lo split the B images are the larger half of the L2 -> SM traffic and these layers are
// L2-bandwidth bound -- at the price of single-buffered accumulators (TMEM: 2 sub-tiles x (main|cross) x 128 = 512).
// F16: fp16 hi/lo operand split (K-step 64) instead of TF32 hi/lo (K-step 32).
template <int MODE, int BN, int G, int MT, bool F16>
__global__ void __launch_bounds__(128 * G + 64, (BN <= 64 ? 2 : 1)) tcgemm_kernel(const TcP p) {
  constexpr int NPROD = 128 * G;
  constexpr int TM = TC_BM * MT;                                 // rows per CTA tile
  constexpr int KS = F16 ? 64 : 32;                              // K values per pipeline stage (one 128-byte smem row)
  constexpr int UPS = F16 ? 2 * MT : MT;                         // gather units (8 x 16 bytes per thread) per stage
  constexpr bool DEFER = MT == 1;                                // double-buffered accumulators -> deferred epilogue
  // one CTA per SM has registers to spare: keep the NEXT unit's gather in flight while this one is split and
  // stored, so a producer group always has 16 KB outstanding (the gather is latency-, not bandwidth-bound)
  constexpr bool PREFETCH = (BN == 128);
  constexpr int n_slots = G;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [slots][A_hi 16K | A_lo 16K | B_hi BN*128 | B_lo BN*128], barriers, TMEM pointer, per-tile aux
  constexpr uint32_t A_BYTES = TC_BM * 128;
  constexpr uint32_t B_BYTES = BN * 128;
  constexpr uint32_t SLOT_BYTES = MT * 2 * A_BYTES + 2 * B_BYTES;
  constexpr int ACC_COLS = MT * 2 * BN;                          // per sub-tile: main | cross-term accumulator
  constexpr int NBUF = DEFER ? 2 : 1;
  constexpr int TMEM_COLS = NBUF * ACC_COLS < 32 ? 32 : NBUF * ACC_COLS;
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;   // shared-space address of slot 0
  const uint32_t full_bar = smem + (uint32_t)n_slots * SLOT_BYTES;   // [n_slots] x 8 bytes
  const uint32_t empty_bar = full_bar + 8u * n_slots;
  const uint32_t tfull_bar = empty_bar + 8u * n_slots;           // [2]
  const uint32_t tempty_bar = tfull_bar + 16u;                   // [2]
  const uint32_t tmem_ptr = tempty_bar + 16u;
  // SPARSE: source ROW of (unit position | tap, tile row) [<= 27][TM]; CONV2D: int4 row info [TM] (16-byte aligned)
  const uint32_t aux_s = (tmem_ptr + 4u + 15u) & ~15u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int Mv = p.M;
  if (p.m_dev) { int md = *p.m_dev; Mv = md < Mv ? md : Mv; }
  const int n_tiles_n = p.cout / BN;
  const int total_tiles = ((Mv + TM - 1) / TM) * n_tiles_n;
  if ((int)blockIdx.x >= total_tiles) return;   // uniform for the whole CTA, before any barrier / TMEM use

  // unit mask of a CTA tile (identical in all three warp roles): bit u = unit u (tap, or multi-tap stage when
  // cin < KS) is used by at least one row of the tile.  No mask / other modes: every unit.
  const bool masked = MODE == FF3D_GEMM_SPARSE && p.tile_mask != nullptr;
  auto unit_mask = [&](int tile) -> uint32_t {
    if (!masked) return 0u;                                      // callers use p.n_stages instead
    const int mtile = (tile / n_tiles_n) * MT;
    uint32_t tm = 0;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
      if ((mtile + mt) * TC_BM < Mv) tm |= __ldg(p.tile_mask + mtile + mt);
    uint32_t um = tm;
    if (p.tps > 1) {
      um = 0;
      const uint32_t grp_bits = (1u << p.tps) - 1u;
      for (int u = 0; u < p.n_units; ++u)
        if ((tm >> (u * p.tps)) & grp_bits) um |= 1u << u;
    }
    return um ? um : 1u;                                         // never an empty tile: unit 0 then gathers zeros
  };
  auto stage_count = [&](uint32_t um) -> int { return masked ? __popc(um) * p.cpt : p.n_stages; };

  if (tid == 0) {
    for (int s = 0; s < n_slots; ++s) { mbar_init(full_bar + 8u * s, TC_BM + 1); mbar_init(empty_bar + 8u * s, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar + 8u * i, 1); mbar_init(tempty_bar + 8u * i, NPROD); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4 * G + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = (uint32_t)lds32(tmem_ptr);

  if (warp < 4 * G) {
    // =========================== A producers (+ deferred epilogue) ===========================
    // 8 consecutive lanes read the 8 16-byte chunks of ONE 128-byte line (a full line per quarter-warp request)
    const int grp = warp >> 2;
    const int pw = warp & 3;
    const int j = lane & 7;
    const int q = lane >> 3;
    const int ptid = tid;                                     // producers are threads [0, NPROD)
    // TF32 small-cin lane geometry: cin = 16 -> 2 taps per 32-wide stage, cin = 8 -> 4 taps
    const int qshift = p.cin == 16 ? 2 : 1;
    const int lane_tap = p.cin >= 32 ? 0 : (j >> qshift);
    const int lane_coff = p.cin >= 32 ? j * 4 : (j & ((1 << qshift) - 1)) * 4;
    // F16: rows of one warp instruction are base + {0, 4, 1, 5}[q]: the two rows of a half-warp differ in bit 2 of
    // (row & 7), so their 8-byte stores land in different swizzle halves of the 128-byte line (no bank conflict)
    const int rq = (q & 1) * 4 + (q >> 1);
    bool ovf = false;
    const int r = tid & 127;                                  // epilogue: thread <-> tile row (TMEM lane)
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int slot = grp;                                     // n_slots == G: group g owns slot g
    uint32_t phase = 0u;
    int gbase = 0;                                            // (stages issued by this CTA so far) mod G

    auto epilogue = [&](int tile, int it) {
      const int m0 = (tile / n_tiles_n) * TM;
      const int n0 = (tile - (tile / n_tiles_n) * n_tiles_n) * BN;
      const int ab = DEFER ? (it & 1) : 0;
      mbar_wait(tfull_bar + 8u * ab, (uint32_t)(DEFER ? ((it >> 1) & 1) : (it & 1)));
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
      const int m = m0 + mt * TC_BM + r;
      const bool rvalid = m < Mv;
      float* yp = nullptr;
      __half* ysp = nullptr;
      if (rvalid) {
        long long orow = m;
        if (MODE == FF3D_GEMM_CONV2D) {
          int hw = p.Ho * p.Wo;
          int cb = m / hw;
          int rr = m - cb * hw;
          int coy = rr / p.Wo, cox = rr - (rr / p.Wo) * p.Wo;
          orow = cb * p.y_bstride + p.y_row0 + (long long)(coy * p.uy + p.dy) * (p.Wo * p.ux) + cox * p.ux + p.dx;
        } else if (MODE == FF3D_GEMM_SPARSE && p.y_row) {
          orow = __ldg(p.y_row + m);
        }
        if (MODE == FF3D_GEMM_SPARSE && p.y_off) yp = p.y + __ldg(p.y_off + m);
        else if (p.y) yp = p.y + orow * p.ldy;
        if (p.ys) ysp = p.ys + orow * p.ldys;
      }
      const uint32_t acc = lane_base + (uint32_t)(ab * ACC_COLS + mt * 2 * BN);
#pragma unroll 1
      for (int c0 = grp * 16; c0 < BN; c0 += 16 * G) {       // 16-column chunks dealt round-robin to the groups
        const int n = n0 + c0;
        // bias / residual loads first: their latency overlaps the TMEM read.  Residual row segment = 64 contiguous
        // bytes per thread -> four 16-byte loads (this thread-per-row epilogue is LSU-wavefront bound)
        float bs[16], rs[16];
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + n + i));   // bias is padded + 16B aligned
            bs[i] = t.x; bs[i + 1] = t.y; bs[i + 2] = t.z; bs[i + 3] = t.w;
          }
        }
        if (p.res && rvalid) {
          const float* rp = p.res + (long long)m * p.ldres + n;
          if ((reinterpret_cast<uintptr_t>(rp) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(rp + i));
              rs[i] = t.x; rs[i + 1] = t.y; rs[i + 2] = t.z; rs[i + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) rs[i] = __ldg(rp + i);
          }
        }
        float v[16], v2[16];
        tmem_ld16(acc + (uint32_t)c0, v);                   // warp-collective: all lanes execute
        tmem_ld16(acc + (uint32_t)(BN + c0), v2);
        if (rvalid) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            // F16: the cross terms carry the 2^11 scale of the lo parts
            float a = F16 ? fmaf(v2[i], 1.f / 2048.f, v[i]) : v[i] + v2[i];
            if (p.bias) a += bs[i];
            if (p.res_after_act) a = apply_act(a, p.act);
            if (p.res) a += rs[i];
            if (!p.res_after_act) a = apply_act(a, p.act);
            v[i] = a;
          }
          if (yp) {
            if ((reinterpret_cast<uintptr_t>(yp + n) & 15) == 0) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(yp + n + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) yp[n + i] = v[i];
            }
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
      }   // mt
      tc_fence_before();
      mbar_arrive(tempty_bar + 8u * ab);                     // NPROD arrivals free the accumulator buffer
    };

    int it = 0, prev_tile = -1;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int m0 = (tile / n_tiles_n) * TM;
      const uint32_t um = unit_mask(tile);
      const int nst = stage_count(um);
      // ---- per-tile gather metadata (both groups are past the previous tile's stages after the first barrier)
      producer_bar<NPROD>();
      if (MODE == FF3D_GEMM_SPARSE) {
        if (masked && p.tps == 1) {
          // one staged row list per PRESENT tap, indexed by its position in the tile's unit list
          int li = 0;
          for (uint32_t mm = um; mm; mm &= mm - 1u, ++li) {
            const int t = __ffs(mm) - 1;
            for (int rr = ptid; rr < TM; rr += NPROD)
              sts32(aux_s + 4u * (uint32_t)(li * TM + rr),
                    (m0 + rr < Mv) ? __ldg(p.nbr + (size_t)t * p.nbr_stride + m0 + rr) : -1);
          }
        } else {
          for (int i = ptid; i < p.taps * TM; i += NPROD) {
            int t = i / TM, rr = i - t * TM;
            sts32(aux_s + 4u * i, (m0 + rr < Mv) ? __ldg(p.nbr + (size_t)t * p.nbr_stride + m0 + rr) : -1);
          }
        }
      } else if (MODE == FF3D_GEMM_CONV2D) {
        for (int rr0 = ptid; rr0 < TM; rr0 += NPROD) {
          int mm = m0 + rr0;
          int4 info = make_int4(0, -32768, -32768, 0);       // rows past M: every tap is out of bounds
          if (mm < Mv) {
            int hw = p.Ho * p.Wo;
            int b = mm / hw;
            int rr = mm - b * hw;
            int oy = rr / p.Wo, ox = rr - (rr / p.Wo) * p.Wo;
            info = make_int4((int)(b * p.x_bstride), oy * p.stride - p.pad, ox * p.stride - p.pad, 1);
          }
          sts128i(aux_s + 16u * rr0, info);
        }
      }
      producer_bar<NPROD>();
      // my stages of this tile: local stage s has global index gbase + s (mod G), mine are those == grp
      int s = (grp - gbase) % G;
      if (s < 0) s += G;
      gbase = (gbase + nst) % G;
      int t, cidx;                                           // unit position and KS-channel chunk of stage s (cin >= KS)
      t = s / p.cpt; cidx = s - t * p.cpt;
      int uu = 0;                                            // gather unit inside stage s
      // source ROW feeding (tile row, tap) -- or -1
      auto src_row = [&](int row, int tap_or_pos, int tap, bool tap_ok) -> long long {
        if (MODE == FF3D_GEMM_ROWS) return (tap_ok && m0 + row < Mv) ? (long long)(m0 + row) : -1;
        if (MODE == FF3D_GEMM_CONV2D) {
          // staged per-tile row info (batch pixel base, top-left input y, x); rows past M hold y = x = -32768
          const int ky = tap / p.kw, kx = tap - ky * p.kw;
          const int4 info = lds128i(aux_s + 16u * (uint32_t)row);
          const int iy = info.y + ky, ix = info.z + kx;
          if (tap_ok && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W)
            return (long long)info.x + (long long)(iy * p.W + ix);
          return -1;
        }
        return tap_ok ? (long long)lds32(aux_s + 4u * (uint32_t)(tap_or_pos * TM + row)) : -1;
      };
      auto load4 = [&](long long so, int coff) -> float4 {
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (so >= 0) {
          const long long e = so * p.ldx + coff;
          val = __ldg(reinterpret_cast<const float4*>(p.x + e));
          if (MODE == FF3D_GEMM_ROWS && p.x2) {
            float4 u = __ldg(reinterpret_cast<const float4*>(p.x2 + e));
            val.x += u.x; val.y += u.y; val.z += u.z; val.w += u.w;
          }
        }
        return val;
      };
      // one gather unit = 8 x 16 bytes per thread.  TF32: the 32 K-values of 128 rows (8 rows per thread);
      // F16: half of the 64 K-values x 128 rows (4 rows x the two 128-byte lines of the row's K-step)
      auto gather = [&](float4* v) {
        if constexpr (!F16) {
          int tap, pos, coff;
          if (p.cin >= KS) { pos = t; tap = t; coff = cidx * KS + lane_coff; }
          else {
            const int uid = masked ? nth_bit(um, s) : s;
            tap = uid * p.tps + lane_tap; pos = tap; coff = lane_coff;
          }
          const bool tap_ok = tap < p.taps;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = uu * TC_BM + pw * 32 + i * 4 + q;
            v[i] = load4(src_row(row, pos, tap, tap_ok), coff);
          }
        } else {
          const int mt = uu >> 1, hf = uu & 1;
          int tapL[2], posL[2], coffL[2];
          const int uid = (p.cin >= KS) ? 0 : (masked ? nth_bit(um, s) : s);
#pragma unroll
          for (int L = 0; L < 2; ++L) {
            const int k0 = L * 32 + 4 * j;                     // K index of my 4 values inside the 64-wide stage
            if (p.cin >= KS) { tapL[L] = t; posL[L] = t; coffL[L] = cidx * KS + k0; }
            else { tapL[L] = uid * p.tps + k0 / p.cin; posL[L] = tapL[L]; coffL[L] = k0 % p.cin; }
          }
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const int i = hf * 4 + ii;
            const int row = mt * TC_BM + pw * 32 + (i >> 1) * 8 + (i & 1) * 2 + rq;
#pragma unroll
            for (int L = 0; L < 2; ++L)
              v[ii * 2 + L] = load4(src_row(row, posL[L], tapL[L], tapL[L] < p.taps), coffL[L]);
          }
        }
      };
      auto advance_cursor = [&]() {
        if (++uu < UPS) return;
        uu = 0;
        s += G;
        if (p.cin >= KS) { cidx += G; while (cidx >= p.cpt) { cidx -= p.cpt; ++t; } }
      };
      // split one gathered unit into the hi / lo images of this group's smem slot; the last unit of a stage hands the
      // slot to the MMA issuer
      auto commit_unit = [&](const float4* v, int u) {
        if (u == 0) mbar_wait(empty_bar + 8u * slot, phase ^ 1u);
        const int mt = F16 ? (u >> 1) : u;
        const uint32_t a_hi = smem + (uint32_t)slot * SLOT_BYTES + (uint32_t)mt * (2 * A_BYTES);
        const uint32_t a_lo = a_hi + A_BYTES;
        if constexpr (!F16) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 h, l;
            split_tf32(v[i].x, h.x, l.x);
            split_tf32(v[i].y, h.y, l.y);
            split_tf32(v[i].z, h.z, l.z);
            split_tf32(v[i].w, h.w, l.w);
            const uint32_t o = swz(pw * 32 + i * 4 + q, j);
            sts128(a_hi + o, h);
            sts128(a_lo + o, l);
          }
        } else {
          const int hf = u & 1;
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const int i = hf * 4 + ii;
            const int row = pw * 32 + (i >> 1) * 8 + (i & 1) * 2 + rq;
#pragma unroll
            for (int L = 0; L < 2; ++L) {
              uint32_t h01, h23, l01, l23;
              split_f16x4(v[ii * 2 + L], h01, h23, l01, l23, ovf);
              // K 32L + 4j .. +3 -> halves [32L + 4j, +4) of the row = 16-byte chunk 4L + j/2, 8-byte half (j & 1)
              const uint32_t o = (uint32_t)(row * 128 + ((((L * 4 + (j >> 1)) ^ (row & 7)) << 4) | ((j & 1) << 3)));
              sts64(a_hi + o, h01, h23);
              sts64(a_lo + o, l01, l23);
            }
          }
        }
        if (u == UPS - 1) {
          fence_proxy_async();
          mbar_arrive(full_bar + 8u * slot);
          phase ^= 1u;
        }
      };
      if (PREFETCH) {
        // two register sets in ping-pong: the loads of the next unit are in flight while this one is split and stored
        // (no register moves between the sets -- a move would wait for the very loads it is supposed to overlap)
        float4 va[8], vb[8];
        int ua = 0, ub = 0;
        bool have = s < nst;
        if (have) { gather(va); ua = uu; advance_cursor(); }
        while (have) {
          bool more = s < nst;
          if (more) { gather(vb); ub = uu; advance_cursor(); }
          commit_unit(va, ua);
          if (!more) break;
          have = s < nst;
          if (have) { gather(va); ua = uu; advance_cursor(); }
          commit_unit(vb, ub);
        }
      } else {
        float4 v[8];
        while (s < nst) {
          gather(v);
          const int u = uu;
          advance_cursor();
          commit_unit(v, u);
        }
      }
      // epilogue of the PREVIOUS tile: its MMAs have had a whole tile's worth of gathers to finish, and the
      // tensor core keeps working on this tile out of the other TMEM accumulator buffer meanwhile
      if (DEFER) {
        if (prev_tile >= 0) epilogue(prev_tile, it - 1);
        prev_tile = tile;
      } else {
        epilogue(tile, it);                                  // single-buffered accumulators (MT = 2)
      }
    }
    if (DEFER && prev_tile >= 0) epilogue(prev_tile, it - 1);
    if (ovf && p.overflow) atomicOr(p.overflow, 1);
  } else if (warp == 4 * G) {
    // =========================== B producer ===========================
    if (lane == 0) {
      Ring ring{0, 0u};
      const uint8_t* wbase = static_cast<const uint8_t*>(p.wimg);
      const size_t stage_bytes = 2 * (size_t)B_BYTES;
      auto copy_stage = [&](const uint8_t* wsrc, int img) {
        mbar_wait(empty_bar + 8u * ring.slot, ring.phase ^ 1u);
        const uint32_t b_hi = smem + (uint32_t)ring.slot * SLOT_BYTES + MT * 2 * A_BYTES;
        mbar_arrive_expect_tx(full_bar + 8u * ring.slot, 2 * B_BYTES);
        bulk_g2s(b_hi, wsrc + (size_t)img * stage_bytes, 2 * B_BYTES, full_bar + 8u * ring.slot);
        ring.advance(1, n_slots);
      };
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int ntile = tile - (tile / n_tiles_n) * n_tiles_n;
        const uint8_t* wsrc = wbase + (size_t)ntile * p.n_stages * stage_bytes;
        if (masked) {
          for (uint32_t mm = unit_mask(tile); mm; mm &= mm - 1u) {
            const int uid = __ffs(mm) - 1;
            for (int c = 0; c < p.cpt; ++c) copy_stage(wsrc, uid * p.cpt + c);
          }
        } else {
          for (int s = 0; s < p.n_stages; ++s) copy_stage(wsrc, s);
        }
      }
    }
  } else {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      // [main | cross] accumulators are adjacent TMEM columns and [B_hi | B_lo] adjacent smem tiles, so
      // A_hi x [B_hi | B_lo] is ONE N = 2*BN MMA; A_lo x B_hi (N = BN) completes the cross term: 2 MMAs and
      // 2 reads of the A tiles per K step instead of 3 (the kernel is shared-memory-bandwidth bound)
      const uint32_t idesc_wide = make_idesc<F16>(2 * BN), idesc_cross = make_idesc<F16>(BN);
      Ring ring{0, 0u};
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int ab = DEFER ? (it & 1) : 0;
        const int nst = stage_count(unit_mask(tile));
        // the epilogue that last read this accumulator buffer (tile it-2, or it-1 when single-buffered) has drained it
        mbar_wait(tempty_bar + 8u * ab, (uint32_t)((DEFER ? ((it >> 1) & 1) : (it & 1)) ^ 1));
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)(ab * ACC_COLS);
        for (int s = 0; s < nst; ++s) {
          mbar_wait(full_bar + 8u * ring.slot, ring.phase);
          tc_fence_after();
          const uint32_t slot_a = smem + (uint32_t)ring.slot * SLOT_BYTES;
          const uint32_t b_hi = slot_a + MT * 2 * A_BYTES;
          static_assert(B_BYTES % 1024 == 0, "B_lo must continue B_hi's 8-row swizzle atoms");
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 4 x (K = 8 tf32 | 16 halves = 32 bytes) per 128-byte swizzled row
            const uint64_t dbh = make_desc(b_hi + k * 32);   // as an N = 2*BN operand it runs on into B_lo
            const uint32_t acc = (s | k) ? 1u : 0u;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint32_t a_hi = slot_a + mt * 2 * A_BYTES, a_lo = a_hi + A_BYTES;
              const uint64_t dah = make_desc(a_hi + k * 32), dal = make_desc(a_lo + k * 32);
              const uint32_t d_main = d_base + (uint32_t)(mt * 2 * BN), d_cross = d_main + BN;
              umma<F16>(d_main, dah, dbh, idesc_wide, acc);    // main += A_hi*B_hi ; cross += A_hi*B_lo
              umma<F16>(d_cross, dal, dbh, idesc_cross, 1u);   // cross += A_lo*B_hi
            }
          }
          umma_commit(empty_bar + 8u * ring.slot);   // frees the smem slot once these MMAs have read it
          ring.advance(1, n_slots);
        }
        umma_commit(tfull_bar + 8u * ab);            // accumulator of this tile complete
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4 * G + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

template <int MODE, int BN, int G, int MT, bool F16>
static int launch_tc_cfg(const TcP& p, int n_tiles_n, cudaStream_t st) {
  constexpr size_t SLOT_BYTES = (size_t)MT * 2 * TC_BM * 128 + 2 * (size_t)BN * 128;
  constexpr int TM = TC_BM * MT;
  constexpr int n_slots = G;                            // BN <= 64: <= 112 KB per CTA so two CTAs share an SM
  size_t smem = (size_t)n_slots * SLOT_BYTES + (2 * n_slots + 4) * sizeof(uint64_t) + 32 +
                (MODE == FF3D_GEMM_SPARSE ? (size_t)TC_MAX_TAPS * TM * sizeof(int)
                                          : (MODE == FF3D_GEMM_CONV2D ? (size_t)TM * 16 : 0)) + 1024;
  // one-time opt-in to > 48 KB of dynamic shared memory: a function-local static initialiser is thread-safe (C++11)
  static const cudaError_t attr = cudaFuncSetAttribute(tcgemm_kernel<MODE, BN, G, MT, F16>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (attr != cudaSuccess) {
    set_error("ff3d_tcgemm: cudaFuncSetAttribute: %s", cudaGetErrorString(attr));
    return FF3D_ECUDA;
  }
  // persistent CTAs: one (BN = 128) or two (BN <= 64) per SM, each looping over output tiles
  long long tiles = (long long)cdiv(p.M, TM) * n_tiles_n;
  long long resident = (long long)num_sms() * (BN <= 64 ? 2 : 1);
  dim3 grid((unsigned)(tiles < resident ? tiles : resident));
  tcgemm_kernel<MODE, BN, G, MT, F16><<<grid, 128 * G + 64, smem, st>>>(p);
  return check_launch("ff3d_tcgemm");
}

template <int MODE, int BN, bool F16>
static int launch_tc(const TcP& p, int n_tiles_n, cudaStream_t st) {
  if constexpr (BN == 128) {
    // enough 256-row tiles to fill the machine -> share each weight stage between two row sub-tiles
    // (only the sparse gather profits: dense layers lose more from the single-buffered accumulators, measured)
    if constexpr (MODE == FF3D_GEMM_SPARSE) {
      if ((long long)cdiv(p.M, 2 * TC_BM) * n_tiles_n >= num_sms()) return launch_tc_cfg<MODE, BN, 2, 2, F16>(p, n_tiles_n, st);
    }
    return launch_tc_cfg<MODE, BN, 3, 1, F16>(p, n_tiles_n, st);
  } else {
    return launch_tc_cfg<MODE, BN, 2, 1, F16>(p, n_tiles_n, st);
  }
}

template <int MODE, bool F16>
static int launch_tc_mode(const TcP& p, int bn, int n_tiles_n, cudaStream_t st) {
  switch (bn) {
    case 16: return launch_tc<MODE, 16, F16>(p, n_tiles_n, st);
    case 32: return launch_tc<MODE, 32, F16>(p, n_tiles_n, st);
    case 64: return launch_tc<MODE, 64, F16>(p, n_tiles_n, st);
    case 128: return launch_tc<MODE, 128, F16>(p, n_tiles_n, st);
  }
  set_error("ff3d_tcgemm: unsupported N tile %d", bn);
  return FF3D_EINVAL;
}

static int ntile_for(int cin, int cout, int ks) {
  const bool cin_ok = (cin >= ks) ? (cin % ks == 0) : (cin >= 8 && (ks % cin) == 0 && (cin & (cin - 1)) == 0);
  if (!cin_ok) return 0;
  if (cout % 128 == 0) return 128;
  if (cout == 64 || cout == 32 || cout == 16) return cout;
  return 0;
}
static int stages_for(int cin, int taps, int ks) {
  if (cin >= ks) return taps * (cin / ks);
  int tps = ks / cin;
  return (taps + tps - 1) / tps;
}

template <bool F16>
static int tcgemm_run(const ff3d_gemm_desc* d, const void* wimg, int bn, int* overflow_dev, ff3d_stream_t stream) {
  constexpr int KS = F16 ? 64 : 32;
  FF3D_REQUIRE(d != nullptr && wimg != nullptr, "ff3d_tcgemm: null argument");
  if (bn == 0) bn = ntile_for(d->cin, d->cout, KS);
  FF3D_REQUIRE(bn > 0 && ntile_for(d->cin, d->cout, KS) > 0 && d->cout % bn == 0 &&
                   (bn == 16 || bn == 32 || bn == 64 || bn == 128),
               "ff3d_tcgemm: shape cin=%d cout=%d (N tile %d) is not tensor-core tileable", d->cin, d->cout, bn);
  FF3D_REQUIRE(d->ldx % 4 == 0 && d->x && (d->y || d->ys) && d->taps > 0, "ff3d_tcgemm: bad operands");
  FF3D_REQUIRE(!d->ys || ((reinterpret_cast<uintptr_t>(d->ys) & 15) == 0 && d->ldys % 8 == 0 && d->ys_lo % 8 == 0),
               "ff3d_tcgemm: split output rows must be 16-byte aligned");
  FF3D_REQUIRE(!d->y_off || d->y, "ff3d_tcgemm: y_off addresses the fp32 output");
  FF3D_REQUIRE((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(wimg) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0,
               "ff3d_tcgemm: x, wimg and bias must be 16-byte aligned");
  if (d->M <= 0) return FF3D_OK;
  TcP p;
  p.mode = d->mode; p.M = d->M; p.m_dev = d->m_dev;
  p.cin = d->cin; p.cout = d->cout; p.taps = d->taps;
  p.x = d->x; p.ldx = d->ldx; p.x2 = d->x2; p.wimg = wimg; p.bias = d->bias;
  p.res = d->res; p.ldres = d->ldres; p.y = d->y; p.ldy = d->ldy; p.act = d->act; p.res_after_act = d->res_after_act;
  p.B = d->B; p.H = d->H; p.W = d->W; p.Ho = d->Ho; p.Wo = d->Wo; p.kh = d->kh; p.kw = d->kw;
  p.stride = d->stride; p.pad = d->pad;
  p.ux = d->ux > 0 ? d->ux : 1; p.uy = d->uy > 0 ? d->uy : 1; p.dx = d->dx; p.dy = d->dy;
  p.x_bstride = d->x_bstride; p.y_bstride = d->y_bstride; p.y_row0 = d->y_row0;
  p.nbr = d->nbr; p.nbr_stride = d->nbr_stride; p.y_off = d->y_off;
  p.tile_mask = d->mode == FF3D_GEMM_SPARSE ? d->tile_mask : nullptr;
  p.y_row = d->mode == FF3D_GEMM_SPARSE ? d->y_row : nullptr;
  p.ys = static_cast<__half*>(d->ys); p.ldys = d->ldys; p.ys_lo = d->ys_lo;
  p.overflow = overflow_dev;
  p.n_stages = stages_for(d->cin, d->taps, KS);
  p.tps = d->cin >= KS ? 1 : KS / d->cin;
  p.cpt = d->cin >= KS ? d->cin / KS : 1;
  p.n_units = d->cin >= KS ? d->taps : p.n_stages;
  if (d->mode == FF3D_GEMM_CONV2D) {
    FF3D_REQUIRE(d->taps == d->kh * d->kw && (long long)d->B * d->Ho * d->Wo == d->M, "ff3d_tcgemm: bad conv geometry");
    if (p.x_bstride == 0) p.x_bstride = (long long)d->H * d->W;
    if (p.y_bstride == 0) p.y_bstride = (long long)d->Ho * p.uy * d->Wo * p.ux;
    FF3D_REQUIRE((long long)d->B * p.x_bstride < 0x7FFFFFFFLL, "ff3d_tcgemm: conv input has more than 2^31 pixels");
  } else if (d->mode == FF3D_GEMM_SPARSE) {
    FF3D_REQUIRE(d->nbr != nullptr && d->nbr_stride >= d->M && d->taps <= TC_MAX_TAPS,
                 "ff3d_tcgemm: sparse mode needs nbr [taps <= 27, >=M]");
  } else {
    FF3D_REQUIRE(d->mode == FF3D_GEMM_ROWS && d->taps == 1, "ff3d_tcgemm: bad mode");
  }
  int n_tiles_n = d->cout / bn;
  cudaStream_t st = as_stream(stream);
  if (d->mode == FF3D_GEMM_ROWS) return launch_tc_mode<FF3D_GEMM_ROWS, F16>(p, bn, n_tiles_n, st);
  if (d->mode == FF3D_GEMM_CONV2D) return launch_tc_mode<FF3D_GEMM_CONV2D, F16>(p, bn, n_tiles_n, st);
  return launch_tc_mode<FF3D_GEMM_SPARSE, F16>(p, bn, n_tiles_n, st);
}

}  // namespace ff3d

// N tile used for a given cout (0 = shape not supported by the tensor-core path)
extern "C" int ff3d_tcgemm_ntile(int cin, int cout) { return ff3d::ntile_for(cin, cout, 32); }
extern "C" int ff3d_tcgemm_stages(int cin, int taps) { return ff3d::stages_for(cin, taps, 32); }
extern "C" int ff3d_tcgemm_f16_ntile(int cin, int cout) { return ff3d::ntile_for(cin, cout, 64); }
extern "C" int ff3d_tcgemm_f16_stages(int cin, int taps) { return ff3d::stages_for(cin, taps, 64); }

extern "C" int ff3d_tcgemm(const ff3d_gemm_desc* d, const float* wimg, ff3d_stream_t stream) {
  return ff3d::tcgemm_run<false>(d, wimg, 0, nullptr, stream);
}

// bn = N tile the weight images were packed for (0 = ff3d_tcgemm_ntile's default).  A smaller tile than the default
// (e.g. 64 for cout = 512) gives small-M / huge-K layers enough CTAs to fill the machine.
extern "C" int ff3d_tcgemm_bn(const ff3d_gemm_desc* d, const float* wimg, int bn, ff3d_stream_t stream) {
  return ff3d::tcgemm_run<false>(d, wimg, bn, nullptr, stream);
}

// fp16 hi/lo split variant (kind::f16): wimg16 = [cout/ntile][ff3d_tcgemm_f16_stages][2][ntile*64] halves (hi | lo*2^11);
// overflow_dev (may be NULL) is OR-ed with 1 when an activation left the fp16 range (the result is then invalid)
extern "C" int ff3d_tcgemm_f16(const ff3d_gemm_desc* d, const void* wimg16, int bn, int* overflow_dev, ff3d_stream_t stream) {
  return ff3d::tcgemm_run<true>(d, wimg16, bn, overflow_dev, stream);
}
