// One kernel per decoder stage: the `num_layers` DeformableTransformer decoder layers + the prediction heads of one
// FocalDecoder stage (focal_decoder.py:826-958, cfg FocalFormer3D_L.py:285-313; [upstream] mmcv BaseTransformerLayer with
// operation_order = (self_attn, norm, cross_attn, norm, ffn, norm)), replacing ~45 launches (13 per layer: qk / v / out
// projections, mha_core, LayerNorm x3, offsets+weights projection, msda, output projection, two FFN GEMMs; + 2 head GEMMs).
//
// Decomposition.  A CTA owns a block of 32 queries of ONE scene for the whole stage; the 32 x 128 query block, its
// positional embedding and every intermediate (q, attention output, sampling offsets / weights, sampled values, the FFN
// hidden layer in 256-wide chunks, the head hidden layer) live in shared memory -- nothing but K / V crosses CTAs.
// Every projection is a [32 x K] x [K x N] product on the tensor cores: mma.sync m16n8k16 (fp16 in, fp32 accumulate)
// with the same hi/lo operand split as the tcgen05 GEMMs (value = hi + lo / 2048, three MMAs per product, the 2^-11-sized
// cross terms in their own accumulator), i.e. fp32-grade results.  A 32-row block is two m16 tiles, far below the 128-row
// tile of tcgen05.mma -- which is why this stage runs on the warp-level MMA and not on the TMEM path: the whole stage is
// 2.6 GFLOP per layer, and its cost is the weight stream (1.5 MB per layer and CTA, from L2) and latency, not MMA rate.
// The weights are packed on the host in B-FRAGMENT order (pack_frag in model.py): for (n-tile of 8, k-step of 16) the 32
// lanes' {hi b0, hi b1, lo b0, lo b1} = one coalesced 512-byte read per warp, straight to registers.
//
// Self-attention (600 x 600 x 8 heads per scene) is the one exchange between CTAs: after the QKV projection each CTA writes
// its 32 keys / values to global memory ALREADY in fragment order (K: B operand of Q K^T per 8 keys; V: B operand of P V
// per 16 keys) and raises a per-scene counter; a CTA starts its attention when the scene's counter shows all blocks of the
// layer.  One warp per head: S = Q K^T by MMA with the soft-max in two passes over the keys (maximum, then exp / sum / P V
// with P re-used from the accumulator registers as the A operand, flash-attention-2 style).  The launch is cooperative
// (all CTAs co-resident; B x ceil(Nq / 32) <= SM count per launch, larger batches are cut into scene groups).
//
// MSDA sampling (mmcv ms_deform_attn): 16 lanes per (query, head) = 4 bilinear corners x 4 channel quads, every lane
// issues the 12 (level, point) loads of its corner as independent LDG.128, the corners are summed by two shuffles.
#include "tc_common.cuh"

namespace ff3d {

constexpr int DS_R = 32;            // queries per CTA
constexpr int DS_NW = 16;           // warps per CTA: (head, 16-row tile) units in the attention, one n-tile each in the projections
constexpr int DS_NT = DS_NW * 32;
constexpr int DS_C = 128;           // hidden channels
constexpr int DS_HEADS = 8;
constexpr int DS_D = 16;
constexpr int DS_MAXL = 4;          // decoder layers per stage
constexpr int DS_LDA = DS_C + 8;    // halves per row of a K = 128 A operand (+8: conflict-free fragment loads)
constexpr int DS_LDY = DS_C + 4;    // floats per row of the fp32 scratch
constexpr int DS_FCH = 256;         // FFN hidden chunk
constexpr int DS_LDH = DS_FCH + 8;
constexpr int DS_LDKV = 2 * DS_C + 4;

struct DsLayer {
  const uint4* w_qkv; const float* b_qkv;     // [384 x 128] rows = (q | k | v)
  const uint4* w_o;   const float* b_o;
  const uint4* w_oa;  const float* b_oa;      // sampling offsets | attention weights, N padded to 16
  const uint4* w_op;  const float* b_op;
  const uint4* w_f1;  const float* b_f1;
  const uint4* w_f2;  const float* b_f2;
  const float* ln_g[3]; const float* ln_b[3];
};

struct DsP {
  int nq, n_layers, nblk;                     // queries per scene, layers, blocks per scene
  int L, P, n_oa;                             // MSDA levels / points, padded width of the offsets|weights projection
  int ffn;                                    // FFN hidden width (multiple of 256)
  int lvl_h[4], lvl_w[4], lvl_start[4];
  const float* x_in; const float* qpe; const float* q_pos;      // [scenes*nq, 128 | 128 | 2]
  float ref_w, ref_h;
  const float* value; int ldv; long long v_bstride;             // [scenes, n_tokens, ldv]; layer j reads columns j*128..
  DsLayer lay[DS_MAXL];
  const uint4* w_h1; const float* b_h1; int n_h1;               // prediction heads: 128 -> n_h1 (ReLU) -> n_pred
  const uint4* w_h2; const float* b_h2; int n_pred;             // n_h1 % 128 == 0, n_pred % 16 == 0 (zero padded)
  float* x_out; float* pred; int ld_pred, pred_cols;
  uint4* kf; uint4* vf;                       // [layer][scene][head][key block][lane] fragments
  int nkb16;                                  // 16-key blocks per scene (= 2 * nblk)
  int* counters;                              // [scenes], zero at launch
  int* overflow;
};

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// fp32 -> (hi, lo * 2^11) halves, saturating like split_f16x4
__device__ __forceinline__ void split1(float v, __half& h, __half& l, bool& ovf) {
  const float c = fminf(fmaxf(v, -F16_MAX), F16_MAX);
  ovf = ovf || c != v;
  h = __float2half_rn(c);
  l = __float2half_rn(fminf(fmaxf((c - __half2float(h)) * 2048.f, -F16_MAX), F16_MAX));
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& h, uint32_t& l, bool& ovf) {
  __half ha, la, hb, lb;
  split1(a, ha, la, ovf);
  split1(b, hb, lb, ovf);
  __half2 hh = __halves2half2(ha, hb), ll = __halves2half2(la, lb);
  h = *reinterpret_cast<uint32_t*>(&hh);
  l = *reinterpret_cast<uint32_t*>(&ll);
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// split of two values known to lie in the fp16 range (soft-max probabilities): packed conversions, no clamps
__device__ __forceinline__ void split2_unit(float a, float b, uint32_t& h, uint32_t& l) {
  const __half2 hh = __floats2half2_rn(a, b);
  const float2 f = __half22float2(hh);
  const __half2 ll = __floats2half2_rn((a - f.x) * 2048.f, (b - f.y) * 2048.f);
  h = *reinterpret_cast<const uint32_t*>(&hh);
  l = *reinterpret_cast<const uint32_t*>(&ll);
}

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// acc += A[32 x 128 halves of K, from column kofs] x W[k-steps ks0 .. ks0+8) for NT adjacent n-tiles starting at nt0.
// Ah / Al: hi / lo planes in shared memory, row stride lda halves.  wf: fragment-packed weights with ks_total k-steps per
// n-tile.  All 8 * NT weight fragments are requested before the first MMA (one 512-byte coalesced read each).
template <int NT>
__device__ __forceinline__ void gemm8(const __half* __restrict__ Ah, const __half* __restrict__ Al, int lda, int kofs,
                                      const uint4* __restrict__ wf, int ks_total, int nt0, int ks0, int lane,
                                      float (&acc)[2][NT][4], float (&accx)[2][NT][4]) {
  uint4 b[NT][8];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) b[nt][ks] = __ldg(wf + ((size_t)(nt0 + nt) * ks_total + ks0 + ks) * 32 + lane);
  // ldmatrix.x4: lanes 0-7 / 8-15 / 16-23 / 24-31 address the rows of the (rows 0-7, k 0-7) / (rows 8-15, k 0-7) /
  // (rows 0-7, k 8-15) / (rows 8-15, k 8-15) 8x8 blocks = the a0..a3 registers of the m16n8k16 A fragment
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lcol = (lane >> 4) * 8;
  const uint32_t ah_base = smem_u32(Ah + lrow * lda + kofs + lcol), al_base = smem_u32(Al + lrow * lda + kofs + lcol);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      ldmatrix_x4(ah[m], ah_base + (uint32_t)((m * 16 * lda + ks * 16) * 2));
      ldmatrix_x4(al[m], al_base + (uint32_t)((m * 16 * lda + ks * 16) * 2));
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        mma16816(acc[m][nt], ah[m], b[nt][ks].x, b[nt][ks].y);       // hi * hi
        mma16816(accx[m][nt], ah[m], b[nt][ks].z, b[nt][ks].w);      // hi * lo   (2^11 scale)
        mma16816(accx[m][nt], al[m], b[nt][ks].x, b[nt][ks].y);      // lo * hi   (2^11 scale)
      }
  }
}

template <int NT>
__device__ __forceinline__ void zero_acc(float (&acc)[2][NT][4], float (&accx)[2][NT][4]) {
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[m][nt][i] = 0.f; accx[m][nt][i] = 0.f; }
}

// visit the accumulator elements of a pair of n-tiles: f(row, col, value) with value = main + cross / 2048
template <int NT, typename F>
__device__ __forceinline__ void for_acc(const float (&acc)[2][NT][4], const float (&accx)[2][NT][4], int n0, int lane, F f) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int row = m * 16 + g + 8 * hrow, col = n0 + nt * 8 + 2 * t;
        f(row, col, fmaf(accx[m][nt][2 * hrow], 1.f / 2048.f, acc[m][nt][2 * hrow]),
          fmaf(accx[m][nt][2 * hrow + 1], 1.f / 2048.f, acc[m][nt][2 * hrow + 1]));
      }
}

// LayerNorm over the 128 columns of the 32 rows of ys (warp w: rows 2w, 2w+1; lane: 4 columns) -> xs (fp32) and, split,
// A (optionally + add[row][col], the positional embedding)
__device__ __forceinline__ void layer_norm_rows(const float* ys, float* xs, const float* __restrict__ gamma,
                                                const float* __restrict__ beta, __half* Ah, __half* Al, const float* add,
                                                int warp, int lane, bool& ovf) {
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + lane), bt = __ldg(reinterpret_cast<const float4*>(beta) + lane);
#pragma unroll
  for (int i = 0; i < DS_R / DS_NW; ++i) {
    const int row = warp * (DS_R / DS_NW) + i;
    const float4 v = *reinterpret_cast<const float4*>(ys + row * DS_LDY + lane * 4);
    float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / DS_C);
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    float q = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / DS_C) + 1e-5f);
    float4 o4 = make_float4(d0 * rstd * gm.x + bt.x, d1 * rstd * gm.y + bt.y, d2 * rstd * gm.z + bt.z, d3 * rstd * gm.w + bt.w);
    *reinterpret_cast<float4*>(xs + row * DS_C + lane * 4) = o4;
    if (add) {
      const float4 a = *reinterpret_cast<const float4*>(add + row * DS_C + lane * 4);
      o4.x += a.x; o4.y += a.y; o4.z += a.z; o4.w += a.w;
    }
    uint32_t h01, h23, l01, l23;
    split_f16x4(o4, h01, h23, l01, l23, ovf);
    *reinterpret_cast<uint2*>(Ah + row * DS_LDA + lane * 4) = make_uint2(h01, h23);
    *reinterpret_cast<uint2*>(Al + row * DS_LDA + lane * 4) = make_uint2(l01, l23);
  }
}

// MSDA: in-place pass over the offsets | weights block of one (query, head): offsets -> pixel coordinates of every
// (level, point) sample, logits -> soft-max weights.  One thread per (query, head).
__device__ __forceinline__ void msda_prepare(float* oas, int ldoa, const float* refs, const DsP& p, int tid) {
  if (tid >= DS_R * DS_HEADS) return;
  const int row = tid >> 3, h = tid & 7;
  const int LP = p.L * p.P, n_off = DS_HEADS * LP * 2;
  float* aw = oas + row * ldoa + n_off + h * LP;
  float* of = oas + row * ldoa + h * LP * 2;
  const float rx = refs[row * 2], ry = refs[row * 2 + 1];
  float mxw = -INFINITY;
  for (int i = 0; i < LP; ++i) mxw = fmaxf(mxw, aw[i]);
  float den = 0.f;
  for (int i = 0; i < LP; ++i) { const float e = expf(aw[i] - mxw); aw[i] = e; den += e; }
  const float inv = 1.f / den;
  for (int i = 0; i < LP; ++i) aw[i] *= inv;
  for (int lv = 0; lv < p.L; ++lv) {
    const float Hf = (float)p.lvl_h[lv], Wf = (float)p.lvl_w[lv];
    for (int pt = 0; pt < p.P; ++pt) {
      const int i = lv * p.P + pt;
      const float lx = rx + of[i * 2] / Wf, ly = ry + of[i * 2 + 1] / Hf;      // mmcv: reference + offset / (W, H)
      of[i * 2] = lx * Wf - 0.5f;                                              // grid_sample, align_corners = False
      of[i * 2 + 1] = ly * Hf - 0.5f;
    }
  }
}

// MSDA sampling of one (query, head) by 16 lanes = 4 bilinear corners x 4 channel quads: all LP loads of the lane's corner
// are issued before the first use (clamped address, zero weight outside the map).  LPC = compile-time L * P (0: generic).
template <int LPC>
__device__ __forceinline__ float4 msda_sample(const float* of, const float* aw, const float* __restrict__ vbase, const DsP& p,
                                              int dx, int dy) {
  float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (LPC > 0) {
    float4 v[LPC];
    float wg[LPC];
#pragma unroll
    for (int i = 0; i < LPC; ++i) {
      const int lv = i / (LPC / 3);                                           // LPC = 3 levels x P points
      const int H = p.lvl_h[lv], Wd = p.lvl_w[lv];
      const float2 pxy = *reinterpret_cast<const float2*>(of + i * 2);
      const float x0f = floorf(pxy.x), y0f = floorf(pxy.y);
      const int xi = (int)x0f + dx, yi = (int)y0f + dy;
      const float fx = pxy.x - x0f, fy = pxy.y - y0f;
      const bool inb = xi >= 0 && xi < Wd && yi >= 0 && yi < H;
      wg[i] = inb ? (dx ? fx : 1.f - fx) * (dy ? fy : 1.f - fy) * aw[i] : 0.f;
      const int xc = min(max(xi, 0), Wd - 1), yc = min(max(yi, 0), H - 1);
      v[i] = __ldg(reinterpret_cast<const float4*>(vbase + ((size_t)p.lvl_start[lv] + (size_t)yc * Wd + xc) * p.ldv));
    }
#pragma unroll
    for (int i = 0; i < LPC; ++i) {
      acc4.x = fmaf(wg[i], v[i].x, acc4.x); acc4.y = fmaf(wg[i], v[i].y, acc4.y);
      acc4.z = fmaf(wg[i], v[i].z, acc4.z); acc4.w = fmaf(wg[i], v[i].w, acc4.w);
    }
  } else {
    for (int lv = 0; lv < p.L; ++lv) {
      const int H = p.lvl_h[lv], Wd = p.lvl_w[lv];
      for (int pt = 0; pt < p.P; ++pt) {
        const int i = lv * p.P + pt;
        const float px = of[i * 2], py = of[i * 2 + 1];
        const float x0f = floorf(px), y0f = floorf(py);
        const int xi = (int)x0f + dx, yi = (int)y0f + dy;
        const float fx = px - x0f, fy = py - y0f;
        if (xi >= 0 && xi < Wd && yi >= 0 && yi < H) {
          const float wgt = (dx ? fx : 1.f - fx) * (dy ? fy : 1.f - fy) * aw[i];
          const float4 v = __ldg(reinterpret_cast<const float4*>(vbase + ((size_t)p.lvl_start[lv] + (size_t)yi * Wd + xi) * p.ldv));
          acc4.x = fmaf(wgt, v.x, acc4.x); acc4.y = fmaf(wgt, v.y, acc4.y);
          acc4.z = fmaf(wgt, v.z, acc4.z); acc4.w = fmaf(wgt, v.w, acc4.w);
        }
      }
    }
  }
  return acc4;
}

__global__ void __launch_bounds__(DS_NT, 1) decoder_stage_kernel(const DsP p) {
  extern __shared__ __align__(16) uint8_t ds_smem[];
  float* xs = reinterpret_cast<float*>(ds_smem);                         // [32][128]   layer input / residual
  float* qpes = xs + DS_R * DS_C;                                        // [32][128]
  float* ys = qpes + DS_R * DS_C;                                        // [32][132]   fp32 GEMM output before LayerNorm
  __half* A1h = reinterpret_cast<__half*>(ys + DS_R * DS_LDY);           // [32][136] x 2 planes, three operands
  __half* A1l = A1h + DS_R * DS_LDA;
  __half* A2h = A1l + DS_R * DS_LDA;
  __half* A2l = A2h + DS_R * DS_LDA;
  __half* Qh = A2l + DS_R * DS_LDA;
  __half* Ql = Qh + DS_R * DS_LDA;
  uint8_t* big = reinterpret_cast<uint8_t*>(Ql + DS_R * DS_LDA);          // K|V staging / offsets|weights / FFN chunk / head hidden
  float* refs = reinterpret_cast<float*>(big + 51200);                   // [32][2]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int scene = blockIdx.x / p.nblk, blk = blockIdx.x - scene * p.nblk;
  const int r0 = blk * DS_R;                                             // first query (scene-local) of this block
  const int n_valid = min(DS_R, p.nq - r0);
  const size_t grow0 = (size_t)scene * p.nq + r0;                        // global row of the block's first query
  bool ovf = false;

  // ---- load the query block
  for (int e = tid; e < DS_R * (DS_C / 4); e += DS_NT) {
    const int row = e >> 5, c4 = e & 31;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (row < n_valid) {
      a = __ldg(reinterpret_cast<const float4*>(p.x_in + (grow0 + row) * DS_C) + c4);
      b = __ldg(reinterpret_cast<const float4*>(p.qpe + (grow0 + row) * DS_C) + c4);
    }
    reinterpret_cast<float4*>(xs)[e] = a;
    reinterpret_cast<float4*>(qpes)[e] = b;
  }
  if (tid < DS_R) {
    float rx = 0.f, ry = 0.f;
    if (tid < n_valid) { rx = __ldg(p.q_pos + (grow0 + tid) * 2) / p.ref_w; ry = __ldg(p.q_pos + (grow0 + tid) * 2 + 1) / p.ref_h; }
    refs[tid * 2] = rx;
    refs[tid * 2 + 1] = ry;
  }
  __syncthreads();

  const int nkb8 = 2 * p.nkb16;
  const int n_scenes = gridDim.x / p.nblk;

  for (int l = 0; l < p.n_layers; ++l) {
    const DsLayer& W = p.lay[l];
    // ---- A1 = split(x + qpe), A2 = split(x)
    for (int e = tid; e < DS_R * (DS_C / 4); e += DS_NT) {
      const int row = e >> 5, c = (e & 31) * 4;
      const float4 a = reinterpret_cast<const float4*>(xs)[e], b = reinterpret_cast<const float4*>(qpes)[e];
      uint32_t h01, h23, l01, l23;
      split_f16x4(a, h01, h23, l01, l23, ovf);
      *reinterpret_cast<uint2*>(A2h + row * DS_LDA + c) = make_uint2(h01, h23);
      *reinterpret_cast<uint2*>(A2l + row * DS_LDA + c) = make_uint2(l01, l23);
      split_f16x4(make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w), h01, h23, l01, l23, ovf);
      *reinterpret_cast<uint2*>(A1h + row * DS_LDA + c) = make_uint2(h01, h23);
      *reinterpret_cast<uint2*>(A1l + row * DS_LDA + c) = make_uint2(l01, l23);
    }
    __syncthreads();
    // ---- q | k | v projection: 48 n-tiles, warp w owns q tile w, k tile w, v tile w; q,k read x + qpe, v reads x
    {
      float* kvst = reinterpret_cast<float*>(big);                       // [32][260]: k | v
#pragma unroll 1
      for (int i = 0; i < 3; ++i) {
        const int nt = i * DS_NW + warp, n0 = nt * 8;
        float acc[2][1][4], accx[2][1][4];
        zero_acc<1>(acc, accx);
        gemm8<1>(i < 2 ? A1h : A2h, i < 2 ? A1l : A2l, DS_LDA, 0, W.w_qkv, 8, nt, 0, lane, acc, accx);
        for_acc<1>(acc, accx, n0, lane, [&](int row, int col, float v0, float v1) {
          v0 += __ldg(W.b_qkv + col);
          v1 += __ldg(W.b_qkv + col + 1);
          if (i == 0) {                                                  // q, scaled by 1 / sqrt(d) = 1/4 (exact)
            uint32_t h, lo;
            split2(v0 * 0.25f, v1 * 0.25f, h, lo, ovf);
            *reinterpret_cast<uint32_t*>(Qh + row * DS_LDA + col) = h;
            *reinterpret_cast<uint32_t*>(Ql + row * DS_LDA + col) = lo;
          } else {
            *reinterpret_cast<float2*>(kvst + row * DS_LDKV + (col - DS_C)) = make_float2(v0, v1);
          }
        });
      }
      __syncthreads();
      // ---- this block's keys / values -> global, in fragment order (rows past nq are zero)
      uint4* kf = p.kf + (size_t)(l * n_scenes + scene) * DS_HEADS * nkb8 * 32;
      uint4* vf = p.vf + (size_t)(l * n_scenes + scene) * DS_HEADS * p.nkb16 * 64;
      for (int item = tid; item < 1024; item += DS_NT) {
        // K: item = (head, 8-key block kb of the 4 in this block, lane' = (g', t'))
        const int h = item >> 7, kb = (item >> 5) & 3, ln = item & 31, gg = ln >> 2, tt = ln & 3;
        const int rl = kb * 8 + gg;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (rl < n_valid) {
          const float* kr = kvst + rl * DS_LDKV + h * DS_D + 2 * tt;
          split2(kr[0], kr[1], o.x, o.z, ovf);
          split2(kr[8], kr[9], o.y, o.w, ovf);
        }
        kf[((size_t)h * nkb8 + blk * 4 + kb) * 32 + ln] = o;
      }
      for (int item = tid; item < 1024; item += DS_NT) {
        // V: item = (head, 16-key block kb of the 2 in this block, d n-tile, lane' = (g', t')): b0 = keys 2t', 2t'+1 ; b1 = +8
        const int h = item >> 7, kb = (item >> 6) & 1, nt = (item >> 5) & 1, ln = item & 31, gg = ln >> 2, tt = ln & 3;
        const int k0 = kb * 16 + 2 * tt;
        const float* vc = kvst + DS_C + h * DS_D + nt * 8 + gg;
        auto val = [&](int key) -> float { return key < n_valid ? vc[key * DS_LDKV] : 0.f; };
        uint4 o;
        split2(val(k0), val(k0 + 1), o.x, o.z, ovf);
        split2(val(k0 + 8), val(k0 + 9), o.y, o.w, ovf);
        vf[(((size_t)h * p.nkb16 + blk * 2 + kb) * 2 + nt) * 32 + ln] = o;
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        atomicAdd(p.counters + scene, 1);
        const int target = p.nblk * (l + 1);
        while (ld_acquire(p.counters + scene) < target) __nanosleep(32);
      }
      __syncthreads();
      // ---- self-attention: warp = (head, 16-row tile)
      {
        const int h = warp & 7, m = warp >> 3;
        const uint4* kfh = kf + (size_t)h * nkb8 * 32 + lane;
        const uint4* vfh = vf + (size_t)h * p.nkb16 * 64 + lane;
        uint32_t qh[4], ql[4];
        {
          const __half* ph = Qh + (m * 16 + g) * DS_LDA + h * DS_D + 2 * t;
          const __half* pl = Ql + (m * 16 + g) * DS_LDA + h * DS_D + 2 * t;
          qh[0] = *reinterpret_cast<const uint32_t*>(ph);
          qh[1] = *reinterpret_cast<const uint32_t*>(ph + 8 * DS_LDA);
          qh[2] = *reinterpret_cast<const uint32_t*>(ph + 8);
          qh[3] = *reinterpret_cast<const uint32_t*>(ph + 8 * DS_LDA + 8);
          ql[0] = *reinterpret_cast<const uint32_t*>(pl);
          ql[1] = *reinterpret_cast<const uint32_t*>(pl + 8 * DS_LDA);
          ql[2] = *reinterpret_cast<const uint32_t*>(pl + 8);
          ql[3] = *reinterpret_cast<const uint32_t*>(pl + 8 * DS_LDA + 8);
        }
        // scores of one 8-key tile: s[0..1] = row g, keys 2t, 2t+1 ; s[2..3] = row g+8
        auto scores = [&](const uint4& kq, float (&s)[4]) {
          float c[4] = {0.f, 0.f, 0.f, 0.f}, cx[4] = {0.f, 0.f, 0.f, 0.f};
          mma16816(c, qh, kq.x, kq.y);
          mma16816(cx, qh, kq.z, kq.w);
          mma16816(cx, ql, kq.x, kq.y);
#pragma unroll
          for (int i = 0; i < 4; ++i) s[i] = fmaf(cx[i], 1.f / 2048.f, c[i]);
        };
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll 8
        for (int kb = 0; kb < nkb8; ++kb) {
          const uint4 kq = __ldcg(kfh + (size_t)kb * 32);
          const int key = kb * 8 + 2 * t;
          float s[4];
          scores(kq, s);
          if (key < p.nq) { mx[0] = fmaxf(mx[0], s[0]); mx[1] = fmaxf(mx[1], s[2]); }
          if (key + 1 < p.nq) { mx[0] = fmaxf(mx[0], s[1]); mx[1] = fmaxf(mx[1], s[3]); }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
          mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
        }
        // exp(s - max) = 2^((s - max) * log2 e): one FFMA + MUFU.EX2 per score
        const float L2E = 1.4426950408889634f;
        const float mb[2] = {mx[0] * L2E, mx[1] * L2E};
        float sum[2] = {0.f, 0.f};
        float o[2][4], ox[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) { o[nt][i] = 0.f; ox[nt][i] = 0.f; }
#pragma unroll 4
        for (int kb = 0; kb < p.nkb16; ++kb) {
          const uint4 ka = __ldcg(kfh + (size_t)(2 * kb) * 32), kbq = __ldcg(kfh + (size_t)(2 * kb + 1) * 32);
          const uint4 v0 = __ldcg(vfh + (size_t)kb * 64), v1 = __ldcg(vfh + (size_t)kb * 64 + 32);
          const int key = kb * 16 + 2 * t;
          float sa[4], sb[4];
          scores(ka, sa);
          scores(kbq, sb);
          float pa[4], pb[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            pa[i] = (key + (i & 1) < p.nq) ? exp2f(fmaf(sa[i], L2E, -mb[i >> 1])) : 0.f;
            pb[i] = (key + 8 + (i & 1) < p.nq) ? exp2f(fmaf(sb[i], L2E, -mb[i >> 1])) : 0.f;
          }
          sum[0] += (pa[0] + pa[1]) + (pb[0] + pb[1]);
          sum[1] += (pa[2] + pa[3]) + (pb[2] + pb[3]);
          uint32_t ah[4], al[4];
          split2_unit(pa[0], pa[1], ah[0], al[0]);                       // probabilities are in [0, 1]
          split2_unit(pa[2], pa[3], ah[1], al[1]);
          split2_unit(pb[0], pb[1], ah[2], al[2]);
          split2_unit(pb[2], pb[3], ah[3], al[3]);
          mma16816(o[0], ah, v0.x, v0.y);
          mma16816(ox[0], ah, v0.z, v0.w);
          mma16816(ox[0], al, v0.x, v0.y);
          mma16816(o[1], ah, v1.x, v1.y);
          mma16816(ox[1], ah, v1.z, v1.w);
          mma16816(ox[1], al, v1.x, v1.y);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 1);
          sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 2);
        }
        // attention output (split) -> A1: the out-projection's operand
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int hr = 0; hr < 2; ++hr) {
            const float inv = 1.f / sum[hr];
            const float a = fmaf(ox[nt][2 * hr], 1.f / 2048.f, o[nt][2 * hr]) * inv;
            const float b = fmaf(ox[nt][2 * hr + 1], 1.f / 2048.f, o[nt][2 * hr + 1]) * inv;
            const int row = m * 16 + g + 8 * hr, col = h * DS_D + nt * 8 + 2 * t;
            uint32_t hh, ll;
            split2(a, b, hh, ll, ovf);
            *reinterpret_cast<uint32_t*>(A1h + row * DS_LDA + col) = hh;
            *reinterpret_cast<uint32_t*>(A1l + row * DS_LDA + col) = ll;
          }
      }
      __syncthreads();
    }
    // ---- attention out-projection + residual -> LayerNorm 0 -> x1 ; A1 = split(x1 + qpe)
    {
      float acc[2][1][4], accx[2][1][4];
      zero_acc<1>(acc, accx);
      gemm8<1>(A1h, A1l, DS_LDA, 0, W.w_o, 8, warp, 0, lane, acc, accx);
      for_acc<1>(acc, accx, warp * 8, lane, [&](int row, int col, float v0, float v1) {
        const float2 r = *reinterpret_cast<const float2*>(xs + row * DS_C + col);
        *reinterpret_cast<float2*>(ys + row * DS_LDY + col) = make_float2(v0 + __ldg(W.b_o + col) + r.x, v1 + __ldg(W.b_o + col + 1) + r.y);
      });
    }
    __syncthreads();
    layer_norm_rows(ys, xs, W.ln_g[0], W.ln_b[0], A1h, A1l, qpes, warp, lane, ovf);
    __syncthreads();
    // ---- sampling offsets | attention weights = (x1 + qpe) W_oa
    float* oas = reinterpret_cast<float*>(big);                          // [32][n_oa + 4]
    const int ldoa = p.n_oa + 4;
#pragma unroll 1
    for (int nt = warp; nt < p.n_oa / 8; nt += DS_NW) {
      float acc[2][1][4], accx[2][1][4];
      zero_acc<1>(acc, accx);
      gemm8<1>(A1h, A1l, DS_LDA, 0, W.w_oa, 8, nt, 0, lane, acc, accx);
      for_acc<1>(acc, accx, nt * 8, lane, [&](int row, int col, float v0, float v1) {
        *reinterpret_cast<float2*>(oas + row * ldoa + col) = make_float2(v0 + __ldg(W.b_oa + col), v1 + __ldg(W.b_oa + col + 1));
      });
    }
    __syncthreads();
    // ---- multi-scale deformable sampling -> A2 (split): 16 lanes per (query, head) = 4 corners x 4 channel quads
    msda_prepare(oas, ldoa, refs, p, tid);
    __syncthreads();
    {
      const int LP = p.L * p.P, n_off = DS_HEADS * LP * 2;
      const int sub = tid >> 4, l16 = tid & 15, corner = l16 >> 2, c4 = l16 & 3;
      const int dy = corner >> 1, dx = corner & 1;
      const bool fixed12 = p.L == 3 && p.P == 4;
      const float* vscene = p.value + (size_t)scene * p.v_bstride * p.ldv + l * DS_C + c4 * 4;
#pragma unroll 1
      for (int it = 0; it < (DS_R * DS_HEADS) / (DS_NT / 16); ++it) {
        const int pairi = it * (DS_NT / 16) + sub, row = pairi >> 3, h = pairi & 7;
        float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < n_valid) {
          const float* aw = oas + row * ldoa + n_off + h * LP;
          const float* of = oas + row * ldoa + h * LP * 2;
          acc4 = fixed12 ? msda_sample<12>(of, aw, vscene + h * DS_D, p, dx, dy) : msda_sample<0>(of, aw, vscene + h * DS_D, p, dx, dy);
        }
        // sum the four corners (lane bits 2, 3 of the 16-lane group)
#pragma unroll
        for (int o = 4; o <= 8; o <<= 1) {
          acc4.x += __shfl_xor_sync(0xffffffffu, acc4.x, o);
          acc4.y += __shfl_xor_sync(0xffffffffu, acc4.y, o);
          acc4.z += __shfl_xor_sync(0xffffffffu, acc4.z, o);
          acc4.w += __shfl_xor_sync(0xffffffffu, acc4.w, o);
        }
        if (corner == 0) {
          uint32_t h01, h23, l01, l23;
          split_f16x4(acc4, h01, h23, l01, l23, ovf);
          *reinterpret_cast<uint2*>(A2h + row * DS_LDA + h * DS_D + c4 * 4) = make_uint2(h01, h23);
          *reinterpret_cast<uint2*>(A2l + row * DS_LDA + h * DS_D + c4 * 4) = make_uint2(l01, l23);
        }
      }
    }
    __syncthreads();
    // ---- MSDA output projection + residual -> LayerNorm 1 -> x2 ; A1 = split(x2)
    {
      float acc[2][1][4], accx[2][1][4];
      zero_acc<1>(acc, accx);
      gemm8<1>(A2h, A2l, DS_LDA, 0, W.w_op, 8, warp, 0, lane, acc, accx);
      for_acc<1>(acc, accx, warp * 8, lane, [&](int row, int col, float v0, float v1) {
        const float2 r = *reinterpret_cast<const float2*>(xs + row * DS_C + col);
        *reinterpret_cast<float2*>(ys + row * DS_LDY + col) = make_float2(v0 + __ldg(W.b_op + col) + r.x, v1 + __ldg(W.b_op + col + 1) + r.y);
      });
    }
    __syncthreads();
    layer_norm_rows(ys, xs, W.ln_g[1], W.ln_b[1], A1h, A1l, nullptr, warp, lane, ovf);
    __syncthreads();
    // ---- FFN: 256-wide chunks of the hidden layer; the second GEMM's accumulators stay in registers across the chunks
    {
      __half* Hh = reinterpret_cast<__half*>(big);                       // [32][264] x 2 planes
      __half* Hl = Hh + DS_R * DS_LDH;
      float accy[2][1][4], accyx[2][1][4];
      zero_acc<1>(accy, accyx);
      const int n_chunks = p.ffn / DS_FCH, ks2 = p.ffn / 16;
#pragma unroll 1
      for (int ch = 0; ch < n_chunks; ++ch) {
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
          const int nt = i * DS_NW + warp;                               // n-tile inside the chunk: columns 8 nt ..
          float acc[2][1][4], accx[2][1][4];
          zero_acc<1>(acc, accx);
          gemm8<1>(A1h, A1l, DS_LDA, 0, W.w_f1, 8, ch * (DS_FCH / 8) + nt, 0, lane, acc, accx);
          for_acc<1>(acc, accx, nt * 8, lane, [&](int row, int col, float v0, float v1) {
            const int gc = ch * DS_FCH + col;
            uint32_t hh, ll;
            split2(fmaxf(v0 + __ldg(W.b_f1 + gc), 0.f), fmaxf(v1 + __ldg(W.b_f1 + gc + 1), 0.f), hh, ll, ovf);
            *reinterpret_cast<uint32_t*>(Hh + row * DS_LDH + col) = hh;
            *reinterpret_cast<uint32_t*>(Hl + row * DS_LDH + col) = ll;
          });
        }
        __syncthreads();
        gemm8<1>(Hh, Hl, DS_LDH, 0, W.w_f2, ks2, warp, ch * 16, lane, accy, accyx);
        gemm8<1>(Hh, Hl, DS_LDH, 128, W.w_f2, ks2, warp, ch * 16 + 8, lane, accy, accyx);
        __syncthreads();
      }
      for_acc<1>(accy, accyx, warp * 8, lane, [&](int row, int col, float v0, float v1) {
        const float2 r = *reinterpret_cast<const float2*>(xs + row * DS_C + col);
        *reinterpret_cast<float2*>(ys + row * DS_LDY + col) = make_float2(v0 + __ldg(W.b_f2 + col) + r.x, v1 + __ldg(W.b_f2 + col + 1) + r.y);
      });
    }
    __syncthreads();
    layer_norm_rows(ys, xs, W.ln_g[2], W.ln_b[2], A1h, A1l, nullptr, warp, lane, ovf);   // A1 = split(x): the heads' operand
    __syncthreads();
  }

  // ---- stage output: query features
  if (p.x_out) {
    for (int e = tid; e < DS_R * (DS_C / 4); e += DS_NT) {
      const int row = e >> 5, c4 = e & 31;
      if (row < n_valid) reinterpret_cast<float4*>(p.x_out + (grow0 + row) * DS_C)[c4] = reinterpret_cast<const float4*>(xs)[e];
    }
  }
  // ---- prediction heads: hh = relu(x W1 + b1) [32 x n_h1] ; pred = hh W2 + b2 (block-diagonal W2)
  {
    const int ldhh = p.n_h1 + 8;
    __half* HHh = reinterpret_cast<__half*>(big);
    __half* HHl = HHh + DS_R * ldhh;
#pragma unroll 1
    for (int nt = warp; nt < p.n_h1 / 8; nt += DS_NW) {
      float acc[2][1][4], accx[2][1][4];
      zero_acc<1>(acc, accx);
      gemm8<1>(A1h, A1l, DS_LDA, 0, p.w_h1, 8, nt, 0, lane, acc, accx);
      for_acc<1>(acc, accx, nt * 8, lane, [&](int row, int col, float v0, float v1) {
        uint32_t hh, ll;
        split2(fmaxf(v0 + __ldg(p.b_h1 + col), 0.f), fmaxf(v1 + __ldg(p.b_h1 + col + 1), 0.f), hh, ll, ovf);
        *reinterpret_cast<uint32_t*>(HHh + row * ldhh + col) = hh;
        *reinterpret_cast<uint32_t*>(HHl + row * ldhh + col) = ll;
      });
    }
    __syncthreads();
    const int ks2 = p.n_h1 / 16;
#pragma unroll 1
    for (int nt = warp; nt < p.n_pred / 8; nt += DS_NW) {
      float acc[2][1][4], accx[2][1][4];
      zero_acc<1>(acc, accx);
#pragma unroll 1
      for (int k8 = 0; k8 < ks2 / 8; ++k8) gemm8<1>(HHh, HHl, ldhh, k8 * 128, p.w_h2, ks2, nt, k8 * 8, lane, acc, accx);
      for_acc<1>(acc, accx, nt * 8, lane, [&](int row, int col, float v0, float v1) {
        if (row < n_valid) {
          float* pr = p.pred + (grow0 + row) * p.ld_pred;
          if (col < p.pred_cols) pr[col] = v0 + __ldg(p.b_h2 + col);
          if (col + 1 < p.pred_cols) pr[col + 1] = v1 + __ldg(p.b_h2 + col + 1);
        }
      });
    }
  }
  if (ovf && p.overflow) atomicOr(p.overflow, 1);
}

constexpr size_t DS_SMEM = (size_t)2 * DS_R * DS_C * 4 + (size_t)DS_R * DS_LDY * 4 + (size_t)6 * DS_R * DS_LDA * 2 + 51200 + DS_R * 2 * 4;

}  // namespace ff3d

extern "C" size_t ff3d_decoder_stage_workspace_bytes(int B, int nq, int n_layers) {
  const size_t nblk = (size_t)(nq + ff3d::DS_R - 1) / ff3d::DS_R;
  // K fragments: per (layer, scene, head) 4*nblk key blocks x 32 lanes x 16 bytes; V the same size; + the counters
  return 2 * (size_t)n_layers * B * ff3d::DS_HEADS * (4 * nblk) * 32 * 16 + 256 + (size_t)B * sizeof(int);
}

extern "C" int ff3d_decoder_stage(const ff3d_decoder_stage_desc* d, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(d != nullptr, "ff3d_decoder_stage: null descriptor");
  FF3D_REQUIRE(d->hidden == DS_C && d->heads == DS_HEADS, "ff3d_decoder_stage: built for 128 channels x 8 heads (got %d x %d)",
               d->hidden, d->heads);
  FF3D_REQUIRE(d->n_layers >= 1 && d->n_layers <= DS_MAXL, "ff3d_decoder_stage: 1..%d layers", DS_MAXL);
  FF3D_REQUIRE(d->n_levels >= 1 && d->n_levels <= 4 && d->n_points >= 1 && d->n_levels * d->n_points <= 32,
               "ff3d_decoder_stage: MSDA levels / points out of range");
  FF3D_REQUIRE(d->ffn > 0 && d->ffn % DS_FCH == 0, "ff3d_decoder_stage: FFN width must be a multiple of %d", DS_FCH);
  const int n_oa = (DS_HEADS * d->n_levels * d->n_points * 3 + 15) / 16 * 16;
  FF3D_REQUIRE((size_t)DS_R * (n_oa + 4) * 4 <= 51200, "ff3d_decoder_stage: offsets / weights block does not fit");
  FF3D_REQUIRE(d->n_h1 > 0 && d->n_h1 % 128 == 0 && (size_t)2 * DS_R * (d->n_h1 + 8) * 2 <= 51200,
               "ff3d_decoder_stage: head hidden width %d (multiple of 128, <= 384)", d->n_h1);
  FF3D_REQUIRE(d->n_pred > 0 && d->n_pred % 16 == 0 && d->pred_cols <= d->n_pred && d->ld_pred >= d->pred_cols,
               "ff3d_decoder_stage: prediction width must be padded to a multiple of 16");
  FF3D_REQUIRE(d->B >= 1 && d->nq >= 1 && d->x_in && d->qpe && d->q_pos && d->value && d->pred && d->workspace,
               "ff3d_decoder_stage: null tensor");
  FF3D_REQUIRE(d->ldv % 4 == 0 && (reinterpret_cast<uintptr_t>(d->value) & 15) == 0, "ff3d_decoder_stage: value rows 16-byte aligned");
  FF3D_REQUIRE(d->workspace_bytes >= ff3d_decoder_stage_workspace_bytes(d->B, d->nq, d->n_layers),
               "ff3d_decoder_stage: workspace too small");
  const int nblk = cdiv(d->nq, DS_R);
  FF3D_REQUIRE(nblk <= num_sms(), "ff3d_decoder_stage: %d queries per scene exceed one co-resident wave", d->nq);
  static const cudaError_t attr =
      cudaFuncSetAttribute(decoder_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DS_SMEM);
  if (attr != cudaSuccess) { set_error("ff3d_decoder_stage: cudaFuncSetAttribute: %s", cudaGetErrorString(attr)); return FF3D_ECUDA; }

  DsP p = {};
  p.nq = d->nq; p.n_layers = d->n_layers; p.nblk = nblk;
  p.L = d->n_levels; p.P = d->n_points; p.n_oa = n_oa; p.ffn = d->ffn;
  for (int i = 0; i < d->n_levels; ++i) { p.lvl_h[i] = d->lvl_h[i]; p.lvl_w[i] = d->lvl_w[i]; p.lvl_start[i] = d->lvl_start[i]; }
  p.ref_w = d->ref_w; p.ref_h = d->ref_h;
  p.ldv = d->ldv; p.v_bstride = d->v_bstride;
  for (int l = 0; l < d->n_layers; ++l) {
    const ff3d_decoder_layer_weights& s = d->layers[l];
    FF3D_REQUIRE(s.w_qkv && s.w_o && s.w_oa && s.w_op && s.w_f1 && s.w_f2, "ff3d_decoder_stage: layer %d weights missing", l);
    DsLayer& q = p.lay[l];
    q.w_qkv = static_cast<const uint4*>(s.w_qkv); q.b_qkv = s.b_qkv;
    q.w_o = static_cast<const uint4*>(s.w_o); q.b_o = s.b_o;
    q.w_oa = static_cast<const uint4*>(s.w_oa); q.b_oa = s.b_oa;
    q.w_op = static_cast<const uint4*>(s.w_op); q.b_op = s.b_op;
    q.w_f1 = static_cast<const uint4*>(s.w_f1); q.b_f1 = s.b_f1;
    q.w_f2 = static_cast<const uint4*>(s.w_f2); q.b_f2 = s.b_f2;
    for (int i = 0; i < 3; ++i) { q.ln_g[i] = s.ln_gamma[i]; q.ln_b[i] = s.ln_beta[i]; }
  }
  p.w_h1 = static_cast<const uint4*>(d->w_h1); p.b_h1 = d->b_h1; p.n_h1 = d->n_h1;
  p.w_h2 = static_cast<const uint4*>(d->w_h2); p.b_h2 = d->b_h2; p.n_pred = d->n_pred;
  p.ld_pred = d->ld_pred; p.pred_cols = d->pred_cols;
  p.nkb16 = 2 * nblk;
  p.overflow = d->overflow_dev;
  cudaStream_t st = as_stream(stream);
  // scenes are independent: groups of as many scenes as fit one co-resident wave (one CTA per SM)
  const int per_launch = num_sms() / nblk;
  uint8_t* ws = static_cast<uint8_t*>(d->workspace);
  const size_t frag_bytes = (size_t)d->n_layers * d->B * DS_HEADS * (4 * (size_t)nblk) * 32 * 16;
  int* counters = reinterpret_cast<int*>(ws + align_up(2 * frag_bytes, 256));
  if (cudaMemsetAsync(counters, 0, (size_t)d->B * sizeof(int), st) != cudaSuccess) return check_launch("ff3d_decoder_stage(memset)");
  for (int s0 = 0; s0 < d->B; s0 += per_launch) {
    const int ns = d->B - s0 < per_launch ? d->B - s0 : per_launch;
    const size_t row0 = (size_t)s0 * d->nq;
    p.x_in = d->x_in + row0 * DS_C; p.qpe = d->qpe + row0 * DS_C; p.q_pos = d->q_pos + row0 * 2;
    p.value = d->value + (size_t)s0 * d->v_bstride * d->ldv;
    p.x_out = d->x_out ? d->x_out + row0 * DS_C : nullptr;
    p.pred = d->pred + row0 * d->ld_pred;
    // fragment buffers of this group: [layer][scene in group][head][...]
    p.kf = reinterpret_cast<uint4*>(ws) + (size_t)d->n_layers * s0 * DS_HEADS * (4 * (size_t)nblk) * 32;
    p.vf = reinterpret_cast<uint4*>(ws + frag_bytes) + (size_t)d->n_layers * s0 * DS_HEADS * (4 * (size_t)nblk) * 32;
    p.counters = counters + s0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ns * nblk));
    cfg.blockDim = dim3(DS_NT);
    cfg.dynamicSmemBytes = DS_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, decoder_stage_kernel, p) != cudaSuccess) return check_launch("ff3d_decoder_stage");
  }
  return check_launch("ff3d_decoder_stage");
}
