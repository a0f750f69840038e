// EXPERIMENTAL (round-2 work item 1) -- NOT part of libff3d.so, not yet run on hardware.
// Compile check only:  make -C focalformer3d_b200/csrc experimental
//
// tcgen05 implicit GEMM with an fp16 hi/lo operand split instead of 3xTF32:
//     a = hi + 2^-11 * lo,   hi = rn_f16(a),   lo = rn_f16((a - hi) * 2^11)            (saturating conversions)
//     D_main += A_hi*B_hi ;  D_cross += A_hi*B_lo + A_lo*B_hi ;  y = D_main + 2^-11 * D_cross
// Same 22-bit products as the TF32 split (tools/split_precision_study.py) for |a| < 65504, but kind::f16 MMAs run at twice
// the TF32 rate and every operand byte in shared memory covers twice the K: one pipeline stage = 64 K-values per row
// (128 bytes of halves), so the shared-memory traffic per product -- the limiter of the TF32 kernel -- halves.
// Values beyond the fp16 range saturate and raise the device-side overflow flag (the caller must fail loudly).
//
// Differences from tcgemm.cu (everything else -- barriers, ring, TMEM double buffering, epilogue -- is identical):
//   * gather unit = 4 row-instructions x 2 loads: lane j of a quarter-warp loads K 4j..4j+3 of the row's first and of its
//     second 128-byte line (both requests are full lines), converts, and writes two 8-byte pieces per tile (STS.64);
//     the four rows of one instruction are base + {0, 4, 1, 5} so that the two rows of a half-warp land in different
//     swizzle halves (bank-conflict free 8-byte stores).
//   * cin in {8, 16, 32} packs 8 / 4 / 2 taps into one 64-wide K step; cin >= 64 walks 64-channel chunks per tap.
//   * weight images: [n_tiles][n_stages][2][BN*64] halves (hi | lo * 2^11), 128B-swizzled rows.
#include "../common.cuh"
#include <cuda_fp16.h>

namespace ff3d_x {
using namespace ff3d;

struct TcP {
  int mode, M;
  const int* m_dev;
  int cin, cout, taps;
  const float* x; int ldx;
  const float* x2;
  const __half* wimg;       // [n_tiles][n_stages][2][BN*64] pre-swizzled hi / (lo * 2^11) fp16 images
  int* overflow;            // device flag: set when an activation saturates the fp16 range
  const float* bias;
  const float* res; int ldres;
  float* y; int ldy;
  int act, res_after_act;
  int B, H, W, Ho, Wo, kh, kw, stride, pad;
  long long x_bstride, y_bstride, y_row0;
  int ux, uy, dx, dy;
  const int* nbr; int nbr_stride;
  const int* y_off;
  int n_stages, tps, cpt;   // pipeline K-steps; taps per stage (cin < 64); 64-channel chunks per tap (cin >= 64)
};

constexpr int TC_BM = 128;
constexpr int TC_MAX_TAPS = 27;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// All shared-memory traffic of this kernel goes through 32-bit shared-space addresses: the 1024-byte alignment of the
// dynamic window is established on the ADDRESS, not by rounding a generic pointer (which loses the address space and
// makes every access a generic LD/ST on the long scoreboard).
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts128i(uint32_t addr, int4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, int v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int lds32(uint32_t addr) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));   // volatile: ordered against the barrier asms
  return v;
}
__device__ __forceinline__ int4 lds128i(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(1)<<16 |
// SBO(1024B>>4)<<32 | version(1)<<46 | layout SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), K-major A and B,
// N>>3 at [17,23), M>>4 at [24,29)
// (A = B = F16: format code 0 at [7,10) and [10,13))
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// round-to-nearest fp32 -> tf32 (result in fp32 layout, low 13 mantissa bits zero)
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// byte offset of (row r, 16-byte chunk j) inside a 128B-swizzled K-major tile whose base is 1024B aligned
__device__ __forceinline__ uint32_t swz(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

// fp16 hi/lo split of four fp32 values -> two packed half2 words each.  Saturating: beyond +-65504 the hi part clamps
// (and the caller's overflow flag is raised); lo = (v - hi) * 2^11 is at most 2^-11 * 2^11 * |v| / 2 and cannot overflow
// unless hi already did.
constexpr float F16_MAX = 65504.f;
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void split_f16x4(const float4& v, uint32_t& h01, uint32_t& h23, uint32_t& l01, uint32_t& l23,
                                            bool& ovf) {
  float c0 = fminf(fmaxf(v.x, -F16_MAX), F16_MAX), c1 = fminf(fmaxf(v.y, -F16_MAX), F16_MAX);
  float c2 = fminf(fmaxf(v.z, -F16_MAX), F16_MAX), c3 = fminf(fmaxf(v.w, -F16_MAX), F16_MAX);
  ovf = ovf || c0 != v.x || c1 != v.y || c2 != v.z || c3 != v.w;      // also true for NaN
  __half2 ha = __floats2half2_rn(c0, c1), hb = __floats2half2_rn(c2, c3);
  float2 fa = __half22float2(ha), fb = __half22float2(hb);
  h01 = *reinterpret_cast<uint32_t*>(&ha);
  h23 = *reinterpret_cast<uint32_t*>(&hb);
  auto lo = [](float x, float h) { return fminf(fmaxf((x - h) * 2048.f, -F16_MAX), F16_MAX); };
  l01 = pack_h2(lo(c0, fa.x), lo(c1, fa.y));
  l23 = pack_h2(lo(c2, fb.x), lo(c3, fb.y));
}

struct Ring {
  int slot;
  uint32_t phase;
  __device__ __forceinline__ void advance(int k, int n) {
    slot += k;
    while (slot >= n) { slot -= n; phase ^= 1u; }
  }
};

template <int N>
__device__ __forceinline__ void producer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }

// G = number of 128-thread A-producer groups = number of smem slots: 2 (and two CTAs per SM) for BN <= 64, 3 for
// BN = 128 (one CTA per SM).  n_slots == G makes every group the sole owner of one slot, so a producer is never more
// than one mbarrier phase ahead of the MMA issuer (the parity wait cannot tell phases two apart).
// MT = 128-row sub-tiles per CTA tile.  MT = 2 (BN = 128, large M) halves the weight bytes streamed from L2 per
// flop -- with 3xTF32 the B images (hi + lo) are the larger half of the L2 -> SM traffic and these layers are
// L2-bandwidth bound -- at the price of single-buffered accumulators (TMEM: 2 sub-tiles x (main|cross) x 128 = 512).
template <int MODE, int BN, int G, int MT>
__global__ void __launch_bounds__(128 * G + 64, (BN <= 64 ? 2 : 1)) tcgemm_kernel(const TcP p, int n_slots) {
  constexpr int NPROD = 128 * G;
  constexpr int TM = TC_BM * MT;                                 // rows per CTA tile
  constexpr bool DEFER = MT == 1;                                // double-buffered accumulators -> deferred epilogue
  // one CTA per SM has registers to spare: keep the NEXT stage's gather in flight while this one is split and
  // stored, so a producer group always has 16 KB outstanding (the gather is latency-, not bandwidth-bound)
  constexpr bool PREFETCH = (BN == 128);
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
  // SPARSE: nbr element offsets [taps][TM]; CONV2D: int4 row info [128] (16-byte aligned)
  const uint32_t aux_s = (tmem_ptr + 4u + 15u) & ~15u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int Mv = p.M;
  if (p.m_dev) { int md = *p.m_dev; Mv = md < Mv ? md : Mv; }
  const int n_tiles_n = p.cout / BN;
  const int total_tiles = ((Mv + TM - 1) / TM) * n_tiles_n;
  if ((int)blockIdx.x >= total_tiles) return;   // uniform for the whole CTA, before any barrier / TMEM use

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
  const int n_stages = p.n_stages;

  if (warp < 4 * G) {
    // =========================== A producers (+ deferred epilogue) ===========================
    // 8 consecutive lanes read the 8 16-byte chunks of ONE row (a full 128-byte line per quarter-warp request),
    // 4 rows per warp instruction, 8 instructions per stage.
    const int grp = warp >> 2;
    const int pw = warp & 3;
    const int j = lane & 7;
    const int q = lane >> 3;
    const int ptid = tid;                                     // producers are threads [0, NPROD)
    // rows of one warp instruction: base(i) + {0, 4, 1, 5}[q]: the two rows of a half-warp differ in bit 2 of (row & 7),
    // so their 8-byte stores land in different swizzle halves of the 128-byte line (no bank conflict)
    const int rq = (q & 1) * 4 + (q >> 1);
    bool ovf = false;
    const int r = tid & 127;                                  // epilogue: thread <-> tile row (TMEM lane)
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    Ring ring{grp % n_slots, (uint32_t)((grp / n_slots) & 1)};

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
      if (rvalid) {
        if (MODE == FF3D_GEMM_CONV2D) {
          int hw = p.Ho * p.Wo;
          int cb = m / hw;
          int rr = m - cb * hw;
          int coy = rr / p.Wo, cox = rr - (rr / p.Wo) * p.Wo;
          long long row = cb * p.y_bstride + p.y_row0 + (long long)(coy * p.uy + p.dy) * (p.Wo * p.ux) + cox * p.ux + p.dx;
          yp = p.y + row * p.ldy;
        } else if (MODE == FF3D_GEMM_SPARSE && p.y_off) {
          yp = p.y + __ldg(p.y_off + m);
        } else {
          yp = p.y + (long long)m * p.ldy;
        }
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
            float a = fmaf(v2[i], 1.f / 2048.f, v[i]);      // cross terms carry the 2^11 scale of the lo parts
            if (p.bias) a += bs[i];
            if (p.res_after_act) a = apply_act(a, p.act);
            if (p.res) a += rs[i];
            if (!p.res_after_act) a = apply_act(a, p.act);
            v[i] = a;
          }
          if ((reinterpret_cast<uintptr_t>(yp + n) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(yp + n + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) yp[n + i] = v[i];
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
      // ---- per-tile gather metadata (both groups are past the previous tile's stages after the first barrier)
      producer_bar<NPROD>();
      if (MODE == FF3D_GEMM_SPARSE) {
        for (int i = ptid; i < p.taps * TM; i += NPROD) {
          int t = i / TM, rr = i - t * TM;
          int v = (m0 + rr < Mv) ? __ldg(p.nbr + (size_t)t * p.nbr_stride + m0 + rr) : -1;
          sts32(aux_s + 4u * i, v < 0 ? -1 : v * p.ldx);    // element offset of the source row
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
      // my stages of this tile: global stage index (it * n_stages + s) has my parity
      int s = (grp - it * n_stages) % G;                      // first stage of this tile with (it*n_stages + s) % G == grp
      if (s < 0) s += G;
      int t, cidx;                                           // tap and 64-channel chunk of stage s (cin >= 64)
      t = s / p.cpt; cidx = s - t * p.cpt;
      int uu = 0;                                            // unit of the gather cursor inside stage s: (sub-tile, half)
      constexpr int UPS = 2 * MT;                            // units per stage
      // one gather unit = 4 row-instructions x 2 loads (first / second 128-byte line of the row's 64 K-values)
      auto gather = [&](float4* v) {
        const int mt = uu >> 1, hf = uu & 1;
        int tapL[2], coffL[2], kyL[2] = {0, 0}, kxL[2] = {0, 0};
#pragma unroll
        for (int L = 0; L < 2; ++L) {
          const int k0 = L * 32 + 4 * j;                     // K index of my 4 values inside the 64-wide stage
          if (p.cin >= 64) { tapL[L] = t; coffL[L] = cidx * 64 + k0; }
          else { tapL[L] = s * p.tps + k0 / p.cin; coffL[L] = k0 % p.cin; }
          if (MODE == FF3D_GEMM_CONV2D) { kyL[L] = tapL[L] / p.kw; kxL[L] = tapL[L] - kyL[L] * p.kw; }
        }
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int i = hf * 4 + ii;
          const int row = mt * TC_BM + pw * 32 + (i >> 1) * 8 + (i & 1) * 2 + rq;
#pragma unroll
          for (int L = 0; L < 2; ++L) {
            const bool tap_ok = tapL[L] < p.taps;
            long long so = -1;                               // element offset of the source row feeding (row, tap)
            if (MODE == FF3D_GEMM_ROWS) {
              if (tap_ok && m0 + row < Mv) so = (long long)(m0 + row) * p.ldx;
            } else if (MODE == FF3D_GEMM_CONV2D) {
              const int4 info = lds128i(aux_s + 16u * (uint32_t)row);
              const int iy = info.y + kyL[L], ix = info.z + kxL[L];
              if (tap_ok && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W)
                so = ((long long)info.x + (long long)(iy * p.W + ix)) * p.ldx;
            } else {
              if (tap_ok) so = (long long)lds32(aux_s + 4u * (uint32_t)(tapL[L] * TM + row));
            }
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (so >= 0) {
              val = __ldg(reinterpret_cast<const float4*>(p.x + so + coffL[L]));
              if (MODE == FF3D_GEMM_ROWS && p.x2) {
                float4 u = __ldg(reinterpret_cast<const float4*>(p.x2 + so + coffL[L]));
                val.x += u.x; val.y += u.y; val.z += u.z; val.w += u.w;
              }
            }
            v[ii * 2 + L] = val;
          }
        }
      };
      auto advance_cursor = [&]() {
        if (++uu < UPS) return;
        uu = 0;
        s += G;
        if (p.cin >= 64) { cidx += G; while (cidx >= p.cpt) { cidx -= p.cpt; ++t; } }
      };
      // convert one gathered unit to the hi / lo fp16 images of this group's smem slot; the last unit of a stage hands
      // the slot to the MMA issuer
      auto commit_unit = [&](const float4* v, int u) {
        const int mt = u >> 1, hf = u & 1;
        if (u == 0) mbar_wait(empty_bar + 8u * ring.slot, ring.phase ^ 1u);
        const uint32_t a_hi = smem + (uint32_t)ring.slot * SLOT_BYTES + (uint32_t)mt * (2 * A_BYTES);
        const uint32_t a_lo = a_hi + A_BYTES;
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
        if (u == UPS - 1) {
          fence_proxy_async();
          mbar_arrive(full_bar + 8u * ring.slot);
          ring.advance(G, n_slots);
        }
      };
      if (PREFETCH) {
        // two register sets in ping-pong: the loads of the next unit are in flight while this one is converted and stored
        float4 va[8], vb[8];
        int ua = 0, ub = 0;
        bool have = s < n_stages;
        if (have) { gather(va); ua = uu; advance_cursor(); }
        while (have) {
          bool more = s < n_stages;
          if (more) { gather(vb); ub = uu; advance_cursor(); }
          commit_unit(va, ua);
          if (!more) break;
          have = s < n_stages;
          if (have) { gather(va); ua = uu; advance_cursor(); }
          commit_unit(vb, ub);
        }
      } else {
        float4 v[8];
        while (s < n_stages) {
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
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int ntile = tile - (tile / n_tiles_n) * n_tiles_n;
        const __half* wsrc = p.wimg + (size_t)ntile * n_stages * (2 * BN * 64);
        for (int s = 0; s < n_stages; ++s) {
          mbar_wait(empty_bar + 8u * ring.slot, ring.phase ^ 1u);
          const uint32_t b_hi = smem + (uint32_t)ring.slot * SLOT_BYTES + MT * 2 * A_BYTES;
          mbar_arrive_expect_tx(full_bar + 8u * ring.slot, 2 * B_BYTES);
          bulk_g2s(b_hi, wsrc + (size_t)s * (2 * BN * 64), 2 * B_BYTES, full_bar + 8u * ring.slot);
          ring.advance(1, n_slots);
        }
      }
    }
  } else {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      // [main | cross] accumulators are adjacent TMEM columns and [B_hi | B_lo] adjacent smem tiles, so
      // A_hi x [B_hi | B_lo] is ONE N = 2*BN MMA; A_lo x B_hi (N = BN) completes the cross term: 2 MMAs and
      // 2 reads of the A tiles per K step instead of 3 (the kernel is shared-memory-bandwidth bound)
      const uint32_t idesc_wide = make_idesc(2 * BN), idesc_cross = make_idesc(BN);
      Ring ring{0, 0u};
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int ab = DEFER ? (it & 1) : 0;
        // the epilogue that last read this accumulator buffer (tile it-2, or it-1 when single-buffered) has drained it
        mbar_wait(tempty_bar + 8u * ab, (uint32_t)((DEFER ? ((it >> 1) & 1) : (it & 1)) ^ 1));
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)(ab * ACC_COLS);
        for (int s = 0; s < n_stages; ++s) {
          mbar_wait(full_bar + 8u * ring.slot, ring.phase);
          tc_fence_after();
          const uint32_t slot_a = smem + (uint32_t)ring.slot * SLOT_BYTES;
          const uint32_t b_hi = slot_a + MT * 2 * A_BYTES;
          static_assert(B_BYTES % 1024 == 0, "B_lo must continue B_hi's 8-row swizzle atoms");
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 4 x (K = 16 halves = 32 bytes) per 128-byte swizzled row
            const uint64_t dbh = make_desc(b_hi + k * 32);   // as an N = 2*BN operand it runs on into B_lo
            const uint32_t acc = (s | k) ? 1u : 0u;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint32_t a_hi = slot_a + mt * 2 * A_BYTES, a_lo = a_hi + A_BYTES;
              const uint64_t dah = make_desc(a_hi + k * 32), dal = make_desc(a_lo + k * 32);
              const uint32_t d_main = d_base + (uint32_t)(mt * 2 * BN), d_cross = d_main + BN;
              umma_f16(d_main, dah, dbh, idesc_wide, acc);    // main += A_hi*B_hi ; cross += A_hi*B_lo
              umma_f16(d_cross, dal, dbh, idesc_cross, 1u);   // cross += A_lo*B_hi
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

template <int MODE, int BN, int G, int MT>
static int launch_tc_cfg(const TcP& p, int n_tiles_n, cudaStream_t st) {
  constexpr size_t SLOT_BYTES = (size_t)MT * 2 * TC_BM * 128 + 2 * (size_t)BN * 128;
  constexpr int TM = TC_BM * MT;
  const int n_slots = G;                                // BN <= 64: <= 112 KB per CTA so two CTAs share an SM
  size_t smem = (size_t)n_slots * SLOT_BYTES + (2 * n_slots + 4) * sizeof(uint64_t) + 32 +
                (MODE == FF3D_GEMM_SPARSE ? (size_t)TC_MAX_TAPS * TM * sizeof(int)
                                          : (MODE == FF3D_GEMM_CONV2D ? (size_t)TM * 16 : 0)) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tcgemm_kernel<MODE, BN, G, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  // persistent CTAs: one (BN = 128) or two (BN <= 64) per SM, each looping over output tiles
  long long tiles = (long long)cdiv(p.M, TM) * n_tiles_n;
  long long resident = (long long)num_sms() * (BN <= 64 ? 2 : 1);
  dim3 grid((unsigned)(tiles < resident ? tiles : resident));
  tcgemm_kernel<MODE, BN, G, MT><<<grid, 128 * G + 64, smem, st>>>(p, n_slots);
  return check_launch("ff3d_x_tcgemm_f16");
}

template <int MODE, int BN>
static int launch_tc(const TcP& p, int n_tiles_n, cudaStream_t st) {
  if constexpr (BN == 128) {
    // enough 256-row tiles to fill the machine -> share each weight stage between two row sub-tiles
    // (only the sparse gather profits: dense layers lose more from the single-buffered accumulators, measured)
    if constexpr (MODE == FF3D_GEMM_SPARSE) {
      if ((long long)cdiv(p.M, 2 * TC_BM) * n_tiles_n >= num_sms()) return launch_tc_cfg<MODE, BN, 2, 2>(p, n_tiles_n, st);
    }
    return launch_tc_cfg<MODE, BN, 3, 1>(p, n_tiles_n, st);
  } else {
    return launch_tc_cfg<MODE, BN, 2, 1>(p, n_tiles_n, st);
  }
}

template <int MODE>
static int launch_tc_mode(const TcP& p, int bn, int n_tiles_n, cudaStream_t st) {
  switch (bn) {
    case 16: return launch_tc<MODE, 16>(p, n_tiles_n, st);
    case 32: return launch_tc<MODE, 32>(p, n_tiles_n, st);
    case 64: return launch_tc<MODE, 64>(p, n_tiles_n, st);
    case 128: return launch_tc<MODE, 128>(p, n_tiles_n, st);
  }
  set_error("ff3d_x_tcgemm_f16: unsupported N tile %d", bn);
  return FF3D_EINVAL;
}

}  // namespace ff3d_x

// N tile used for a given cout (0 = shape not supported by the tensor-core path)
extern "C" int ff3d_x_tcgemm_f16_ntile(int cin, int cout) {
  if (!(cin == 8 || cin == 16 || cin == 32 || (cin >= 64 && cin % 64 == 0))) return 0;
  if (cout % 128 == 0) return 128;
  if (cout == 64 || cout == 32 || cout == 16) return cout;
  return 0;
}
extern "C" int ff3d_x_tcgemm_f16_stages(int cin, int taps) {
  if (cin >= 64) return taps * (cin / 64);
  int tps = 64 / cin;
  return (taps + tps - 1) / tps;
}

// Experimental entry point (same descriptor as ff3d_tcgemm).  wimg16 = fp16 images packed by
// focalformer3d_b200/experimental_f16.py; overflow_dev (may be NULL) is OR-ed with 1 when an activation left the fp16 range.
extern "C" int ff3d_x_tcgemm_f16(const ff3d_gemm_desc* d, const void* wimg16, int bn, int* overflow_dev,
                                 ff3d_stream_t stream) {
  using namespace ff3d;
  using namespace ff3d_x;
  FF3D_REQUIRE(d != nullptr && wimg16 != nullptr, "ff3d_x_tcgemm_f16: null argument");
  if (bn == 0) bn = ff3d_x_tcgemm_f16_ntile(d->cin, d->cout);
  FF3D_REQUIRE(bn > 0 && ff3d_x_tcgemm_f16_ntile(d->cin, d->cout) > 0 && d->cout % bn == 0 &&
                   (bn == 16 || bn == 32 || bn == 64 || bn == 128),
               "ff3d_x_tcgemm_f16: shape cin=%d cout=%d (N tile %d) is not tileable", d->cin, d->cout, bn);
  FF3D_REQUIRE(d->ldx % 4 == 0 && d->x && d->y && d->taps > 0, "ff3d_x_tcgemm_f16: bad operands");
  FF3D_REQUIRE((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(wimg16) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0,
               "ff3d_x_tcgemm_f16: x, wimg and bias must be 16-byte aligned");
  if (d->M <= 0) return FF3D_OK;
  TcP p;
  p.mode = d->mode; p.M = d->M; p.m_dev = d->m_dev;
  p.cin = d->cin; p.cout = d->cout; p.taps = d->taps;
  p.x = d->x; p.ldx = d->ldx; p.x2 = d->x2; p.wimg = static_cast<const __half*>(wimg16); p.overflow = overflow_dev;
  p.bias = d->bias;
  p.res = d->res; p.ldres = d->ldres; p.y = d->y; p.ldy = d->ldy; p.act = d->act; p.res_after_act = d->res_after_act;
  p.B = d->B; p.H = d->H; p.W = d->W; p.Ho = d->Ho; p.Wo = d->Wo; p.kh = d->kh; p.kw = d->kw;
  p.stride = d->stride; p.pad = d->pad;
  p.ux = d->ux > 0 ? d->ux : 1; p.uy = d->uy > 0 ? d->uy : 1; p.dx = d->dx; p.dy = d->dy;
  p.x_bstride = d->x_bstride; p.y_bstride = d->y_bstride; p.y_row0 = d->y_row0;
  p.nbr = d->nbr; p.nbr_stride = d->nbr_stride; p.y_off = d->y_off;
  p.n_stages = ff3d_x_tcgemm_f16_stages(d->cin, d->taps);
  p.tps = d->cin >= 64 ? 1 : 64 / d->cin;
  p.cpt = d->cin >= 64 ? d->cin / 64 : 1;
  if (d->mode == FF3D_GEMM_CONV2D) {
    FF3D_REQUIRE(d->taps == d->kh * d->kw && (long long)d->B * d->Ho * d->Wo == d->M, "ff3d_x_tcgemm_f16: bad conv geometry");
    if (p.x_bstride == 0) p.x_bstride = (long long)d->H * d->W;
    if (p.y_bstride == 0) p.y_bstride = (long long)d->Ho * p.uy * d->Wo * p.ux;
  } else if (d->mode == FF3D_GEMM_SPARSE) {
    FF3D_REQUIRE(d->nbr != nullptr && d->nbr_stride >= d->M && d->taps <= TC_MAX_TAPS,
                 "ff3d_x_tcgemm_f16: sparse mode needs nbr [taps <= 27, >=M]");
  } else {
    FF3D_REQUIRE(d->mode == FF3D_GEMM_ROWS && d->taps == 1, "ff3d_x_tcgemm_f16: bad mode");
  }
  int n_tiles_n = d->cout / bn;
  cudaStream_t st = as_stream(stream);
  if (d->mode == FF3D_GEMM_ROWS) return launch_tc_mode<FF3D_GEMM_ROWS>(p, bn, n_tiles_n, st);
  if (d->mode == FF3D_GEMM_CONV2D) return launch_tc_mode<FF3D_GEMM_CONV2D>(p, bn, n_tiles_n, st);
  return launch_tc_mode<FF3D_GEMM_SPARSE>(p, bn, n_tiles_n, st);
}
