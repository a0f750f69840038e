// Decoder-side kernels: sine position embedding, query self-attention core, multi-scale deformable sampling,
// ROI grid sampling, box-state update and final box decode.  Token-major layout everywhere: [B*Nq, C] rows.
#include "tc_common.cuh"

namespace ff3d {

// ------------------------------------------------------------------------------------------------------------
// gen_sineembed_for_position (projects/mmdet3d_plugin/models/utils/utils.py:40-66)
__global__ void sine_embed_kernel(const float* __restrict__ pos, float w, float h, const float* __restrict__ dim_t,
                                  float* __restrict__ out, int rows) {
  long long total = (long long)rows * 256;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int r = (int)(e >> 8);
    int j = (int)(e & 255);
    int jj = j & 127;
    // out[:, :128] embeds y, out[:, 128:] embeds x
    float ref = (j < 128) ? pos[r * 2 + 1] / h : pos[r * 2 + 0] / w;
    float a = ref * 6.283185307179586f / dim_t[jj];
    out[e] = (jj & 1) ? cosf(a) : sinf(a);
  }
}

// ------------------------------------------------------------------------------------------------------------
// softmax(q k^T / sqrt(d)) v, one warp per query, K/V of one (batch, head) staged in shared memory
constexpr int MHA_MAXQ = 1024;
// One LANE per query, the keys split over the four warps of the CTA (round 2; the round-1 kernel put one warp on a query
// with the lanes over the keys: two LDS per FMA, shared-memory-issue bound at ~90 us for 4 x 8 x 600 x 600).  The K / V
// rows of the (scene, head) sit in shared memory as 64-byte rows and every lane of a warp reads the SAME row -- one
// broadcast LDS.128 feeds 32 lanes x 4 values -- while q, the output accumulators and the soft-max state stay in registers.
// Each warp runs a two-pass soft-max (row maximum, then exp / sum / weighted values) over its quarter of the keys; the four
// partial (max, sum, acc) states of a query are merged through shared memory (the flash-attention rescaling, in fp32).
template <int D>
__global__ void __launch_bounds__(128) mha_core_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k,
                                                       int ldk, const float* __restrict__ v, int ldv,
                                                       float* __restrict__ out, int ldo, int Nq, int heads) {
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;                      // [Nq][D]
  float* Vs = sm + (size_t)Nq * D;
  float* part = Vs + (size_t)Nq * D;   // [4 warps][32 queries][D + 2]: partial acc, max, sum
  const int bh = blockIdx.x;
  const int b = bh / heads, h = bh - b * heads;
  const size_t row0 = (size_t)b * Nq;
  for (int e = threadIdx.x; e < Nq * (D / 4); e += blockDim.x) {
    const int j = e / (D / 4), c4 = e - j * (D / 4);
    reinterpret_cast<float4*>(Ks)[e] = __ldg(reinterpret_cast<const float4*>(k + (row0 + j) * ldk + h * D) + c4);
    reinterpret_cast<float4*>(Vs)[e] = __ldg(reinterpret_cast<const float4*>(v + (row0 + j) * ldv + h * D) + c4);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qi = blockIdx.y * 32 + lane;
  const bool qvalid = qi < Nq;
  const int per = (Nq + 3) / 4;
  const int j0 = warp * per, j1 = min(Nq, j0 + per);
  const float scale = rsqrtf((float)D);
  float qr[D];
#pragma unroll
  for (int c = 0; c < D; c += 4) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (qvalid) t = __ldg(reinterpret_cast<const float4*>(q + (row0 + qi) * ldq + h * D + c));
    qr[c] = t.x * scale; qr[c + 1] = t.y * scale; qr[c + 2] = t.z * scale; qr[c + 3] = t.w * scale;
  }
  auto score = [&](int j) -> float {
    const float4* kr = reinterpret_cast<const float4*>(Ks + (size_t)j * D);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int c = 0; c < D / 4; ++c) {
      const float4 t = kr[c];
      s0 = fmaf(qr[4 * c], t.x, s0); s1 = fmaf(qr[4 * c + 1], t.y, s1);
      s2 = fmaf(qr[4 * c + 2], t.z, s2); s3 = fmaf(qr[4 * c + 3], t.w, s3);
    }
    return (s0 + s1) + (s2 + s3);
  };
  float mx = -INFINITY;
#pragma unroll 4
  for (int j = j0; j < j1; ++j) mx = fmaxf(mx, score(j));
  float sum = 0.f;
  float acc[D];
#pragma unroll
  for (int c = 0; c < D; ++c) acc[c] = 0.f;
#pragma unroll 2
  for (int j = j0; j < j1; ++j) {
    const float pj = expf(score(j) - mx);
    sum += pj;
    const float4* vr = reinterpret_cast<const float4*>(Vs + (size_t)j * D);
#pragma unroll
    for (int c = 0; c < D / 4; ++c) {
      const float4 t = vr[c];
      acc[4 * c] = fmaf(pj, t.x, acc[4 * c]); acc[4 * c + 1] = fmaf(pj, t.y, acc[4 * c + 1]);
      acc[4 * c + 2] = fmaf(pj, t.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(pj, t.w, acc[4 * c + 3]);
    }
  }
  float* mine = part + ((size_t)warp * 32 + lane) * (D + 2);
#pragma unroll
  for (int c = 0; c < D; ++c) mine[c] = acc[c];
  mine[D] = mx;
  mine[D + 1] = sum;
  __syncthreads();
  if (warp == 0 && qvalid) {
    float m = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) m = fmaxf(m, part[((size_t)w * 32 + lane) * (D + 2) + D]);
    float tot = 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float* pw = part + ((size_t)w * 32 + lane) * (D + 2);
      const float f = pw[D] == -INFINITY ? 0.f : expf(pw[D] - m);      // a warp with no keys contributes nothing
      tot = fmaf(f, pw[D + 1], tot);
#pragma unroll
      for (int c = 0; c < D; ++c) acc[c] = fmaf(f, pw[c], acc[c]);
    }
    const float inv = 1.f / tot;
    float* op = out + (row0 + qi) * ldo + h * D;
#pragma unroll
    for (int c = 0; c < D; c += 4)
      *reinterpret_cast<float4*>(op + c) = make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
  }
}

// ------------------------------------------------------------------------------------------------------------
struct Levels {
  int L;
  int h[8], w[8], start[8];
};

__device__ __forceinline__ float bilinear_zero(const float* __restrict__ base, long long ld, int H, int W, float px,
                                               float py, int col) {
  // grid_sample(align_corners=False, padding_mode='zeros') on pixel coordinates (px, py); base = token (0,0)
  float x0f = floorf(px), y0f = floorf(py);
  int x0 = (int)x0f, y0 = (int)y0f;
  float fx = px - x0f, fy = py - y0f;
  float v = 0.f;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      int xi = x0 + dx, yi = y0 + dy;
      if (xi >= 0 && xi < W && yi >= 0 && yi < H) {
        float wgt = (dx ? fx : 1.f - fx) * (dy ? fy : 1.f - fy);
        v = fmaf(wgt, __ldg(base + ((long long)yi * W + xi) * ld + col), v);
      }
    }
  return v;
}

// thread = (b, q, head, channel); d consecutive lanes share one (b,q,head)
__global__ void msda_kernel(const float* __restrict__ value, int ldv, int v_col0, long long v_bstride, Levels lv, int P,
                            const float* __restrict__ ref, float ref_w, float ref_h, const float* __restrict__ offs,
                            int ldoffs, const float* __restrict__ attw, int ldattw, float* __restrict__ out, int B, int Nq, int heads,
                            int d) {
  long long total = (long long)B * Nq * heads * d;
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= total) return;
  int c = (int)(e % d);
  long long r = e / d;
  int h = (int)(r % heads);
  long long bq = r / heads;
  int b = (int)(bq / Nq);
  const int LP = lv.L * P;
  const float* aw = attw + bq * ldattw + h * LP;
  const float* of = offs + bq * ldoffs + h * LP * 2;
  float rx = ref[bq * 2 + 0] / ref_w, ry = ref[bq * 2 + 1] / ref_h;
  float mx = -INFINITY;
  for (int i = 0; i < LP; ++i) mx = fmaxf(mx, aw[i]);
  float den = 0.f;
  for (int i = 0; i < LP; ++i) den += expf(aw[i] - mx);
  float inv = 1.f / den;
  float acc = 0.f;
  const float* vb = value + (long long)b * v_bstride * ldv;
  for (int l = 0; l < lv.L; ++l) {
    int H = lv.h[l], W = lv.w[l];
    const float* base = vb + (long long)lv.start[l] * ldv;
    for (int pnt = 0; pnt < P; ++pnt) {
      int i = l * P + pnt;
      float lx = rx + of[i * 2 + 0] / (float)W;
      float ly = ry + of[i * 2 + 1] / (float)H;
      float px = lx * (float)W - 0.5f, py = ly * (float)H - 0.5f;
      float a = expf(aw[i] - mx) * inv;
      acc = fmaf(a, bilinear_zero(base, ldv, H, W, px, py, v_col0 + h * d + c), acc);
    }
  }
  out[bq * (heads * d) + h * d + c] = acc;
}

// ------------------------------------------------------------------------------------------------------------
struct RoiP {
  int box_ld, g, C;
  float expand, cell_x, cell_y, origin_x, origin_y, rx0, ry0, rx1, ry1;
};
// block = one query, threads = channels; out row = [L][g*g][C]
// bilinear sample of 4 consecutive channels (grid_sample align_corners=False, zero padding), base = token (0, 0)
__device__ __forceinline__ float4 bilinear_zero4(const float* __restrict__ base, long long ld, int H, int W, float px, float py,
                                                 int col) {
  const float x0f = floorf(px), y0f = floorf(py);
  const int x0 = (int)x0f, y0 = (int)y0f;
  const float fx = px - x0f, fy = py - y0f;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const int xi = x0 + dx, yi = y0 + dy;
      if (xi >= 0 && xi < W && yi >= 0 && yi < H) {
        const float wgt = (dx ? fx : 1.f - fx) * (dy ? fy : 1.f - fy);
        const float4 t = __ldg(reinterpret_cast<const float4*>(base + ((long long)yi * W + xi) * ld + col));
        v.x = fmaf(wgt, t.x, v.x); v.y = fmaf(wgt, t.y, v.y); v.z = fmaf(wgt, t.z, v.z); v.w = fmaf(wgt, t.w, v.w);
      }
    }
  return v;
}

// One warp per (query, grid point): the lanes cover the channels four at a time (one 512-byte row per corner and level),
// out row = [L][g*g][C].  SPLIT: the row is written as fp16 [hi(K) | lo(K)] (K = L*g*g*C) for the TMA-fed roi_mlp.0 GEMM.
template <bool SPLIT>
__global__ void __launch_bounds__(256) roi_sample_kernel(const float* __restrict__ qbox, RoiP p, const float* __restrict__ value,
                                                         int ldv, long long v_bstride, Levels lv, float* __restrict__ out,
                                                         __half* __restrict__ outs, int Nq, long long n_items, int* overflow) {
  const int lane = threadIdx.x & 31;
  const int G = p.g * p.g;
  const long long K = (long long)lv.L * G * p.C;
  bool ovf = false;
  for (long long item = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; item < n_items;
       item += ((long long)gridDim.x * blockDim.x) >> 5) {
    const long long bq = item / G;
    const int n = (int)(item - bq * G);
    const int b = (int)(bq / Nq);
    const float* qb = qbox + bq * p.box_ld;
    // decode_box (transfusion_bbox_coder.py:54-69) on (centre, expanded log-dims, sin/cos)
    const float cx = qb[0] * p.cell_x + p.origin_x;
    const float cy = qb[1] * p.cell_y + p.origin_y;
    const float w = expf(qb[3] * p.expand), l = expf(qb[4] * p.expand);
    const float yaw = atan2f(qb[6], qb[7]);
    const float sn = sinf(yaw), cs = cosf(yaw);
    const float* vb = value + (long long)b * v_bstride * ldv;
    const int i = n / p.g, j = n - i * p.g;
    const float gx = ((float)i + 0.5f) / (float)p.g * w - w / 2.f;
    const float gy = ((float)j + 0.5f) / (float)p.g * l - l / 2.f;
    // [upstream] mmdet3d v0.17.1 rotation_3d_in_axis(axis=2): x' = x cos + y sin, y' = -x sin + y cos
    const float wx = gx * cs + gy * sn + cx;
    const float wy = -gx * sn + gy * cs + cy;
    float nx = (wx - p.rx0) / (p.rx1 - p.rx0) * 2.f - 1.f;
    float ny = (wy - p.ry0) / (p.ry1 - p.ry0) * 2.f - 1.f;
    nx = fminf(fmaxf(nx, -2.f), 2.f);
    ny = fminf(fmaxf(ny, -2.f), 2.f);
    for (int lvl = 0; lvl < lv.L; ++lvl) {
      const int H = lv.h[lvl], W = lv.w[lvl];
      const float px = ((nx + 1.f) * (float)W - 1.f) / 2.f;
      const float py = ((ny + 1.f) * (float)H - 1.f) / 2.f;
      const float* base = vb + (long long)lv.start[lvl] * ldv;
      for (int c = lane * 4; c < p.C; c += 128) {
        const float4 v = bilinear_zero4(base, ldv, H, W, px, py, c);
        const long long k = ((long long)lvl * G + n) * p.C + c;
        if (SPLIT) {
          uint32_t h01, h23, l01, l23;
          split_f16x4(v, h01, h23, l01, l23, ovf);
          *reinterpret_cast<uint2*>(outs + bq * 2 * K + k) = make_uint2(h01, h23);
          *reinterpret_cast<uint2*>(outs + bq * 2 * K + K + k) = make_uint2(l01, l23);
        } else {
          *reinterpret_cast<float4*>(out + bq * K + k) = v;
        }
      }
    }
  }
  if (SPLIT && ovf && overflow) atomicOr(overflow, 1);
}

// ------------------------------------------------------------------------------------------------------------
// box-state update after the prediction heads (focal_decoder.py:945-957). pred row = [center2 height1 dim3 rot2 (vel2) cls..]
__global__ void head_update_kernel(float* __restrict__ pred, int ldp, float* __restrict__ query_pos,
                                   const float* __restrict__ prev, int ldprev, int rows) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float* pr = pred + (size_t)r * ldp;
  pr[0] += query_pos[r * 2 + 0];
  pr[1] += query_pos[r * 2 + 1];
  query_pos[r * 2 + 0] = pr[0];
  query_pos[r * 2 + 1] = pr[1];
  if (prev) {
    const float* pv = prev + (size_t)r * ldprev;
    pr[3] += pv[3];
    pr[4] += pv[4];
    pr[6] += pv[6];
    pr[7] += pv[7];
  }
}


// class-aware regression (focal_decoder.py:940-943): the heads emit one (centre, height, dim, rot) set per class,
// channel = class * k + d; every query keeps the set of its own class.  full rows = [g0: nc*k0 | g1: nc*k1 | ... | tail],
// out rows = [k0 | k1 | ... | tail].  One thread per (row, output column).
struct ClsSelP { int n_groups, k[8], tail, nc, ld_full, ld_out; };
__global__ void class_select_kernel(const float* __restrict__ full, const int* __restrict__ label, ClsSelP p,
                                    float* __restrict__ out, int rows, int out_cols) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)rows * out_cols) return;
  int r = (int)(e / out_cols), c = (int)(e - (long long)r * out_cols);
  int lab = min(max(label[r], 0), p.nc - 1);
  int src = 0, c0 = 0;
  bool done = false;
  for (int g = 0; g < p.n_groups; ++g) {
    if (!done && c < c0 + p.k[g]) { src += lab * p.k[g] + (c - c0); done = true; }
    if (!done) { src += p.nc * p.k[g]; c0 += p.k[g]; }
  }
  if (!done) src += c - c0;                                // tail (class logits): copied through
  out[(size_t)r * p.ld_out + c] = full[(size_t)r * p.ld_full + src];
}

struct DecP {
  int C, has_vel, cls_col, ldp;
  float cell_x, cell_y, origin_x, origin_y, pr[6];
};
__global__ void box_decode_kernel(const float* __restrict__ pred, const float* __restrict__ qscore,
                                  const int* __restrict__ qlabel, DecP p, float* __restrict__ boxes,
                                  float* __restrict__ scores, int* __restrict__ labels, unsigned char* __restrict__ keep,
                                  int rows) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* pr = pred + (size_t)r * p.ldp;
  int lab = qlabel[r];
  float s = 1.f / (1.f + expf(-pr[p.cls_col + lab]));
  s = s * qscore[(size_t)r * p.C + lab];
  // score = sigmoid(cls) * query_heatmap_score * one_hot(label); max/argmax over classes (first index on ties)
  int outl = lab;
  if (!(s > 0.f)) { s = 0.f; outl = 0; }
  float x = pr[0] * p.cell_x + p.origin_x;
  float y = pr[1] * p.cell_y + p.origin_y;
  float dw = expf(pr[3]), dl = expf(pr[4]), dh = expf(pr[5]);
  float z = pr[2] - dh * 0.5f;
  float yaw = atan2f(pr[6], pr[7]);
  int code = p.has_vel ? 9 : 7;
  float* bo = boxes + (size_t)r * code;
  bo[0] = x; bo[1] = y; bo[2] = z; bo[3] = dw; bo[4] = dl; bo[5] = dh; bo[6] = yaw;
  if (p.has_vel) { bo[7] = pr[8]; bo[8] = pr[9]; }
  scores[r] = s;
  labels[r] = outl;
  keep[r] = (x >= p.pr[0] && y >= p.pr[1] && z >= p.pr[2] && x <= p.pr[3] && y <= p.pr[4] && z <= p.pr[5]) ? 1 : 0;
}

static int fill_levels(Levels* lv, const int* h, const int* w, const int* s, int L) {
  if (L < 1 || L > 8) return -1;
  lv->L = L;
  for (int i = 0; i < L; ++i) { lv->h[i] = h[i]; lv->w[i] = w[i]; lv->start[i] = s[i]; }
  return 0;
}

}  // namespace ff3d

extern "C" int ff3d_sine_embed(const float* pos, float w, float h, const float* dim_t, float* out, int rows,
                               ff3d_stream_t stream) {
  using namespace ff3d;
  if (rows <= 0) return FF3D_OK;
  long long total = (long long)rows * 256;
  long long nb = (total + 255) / 256;
  long long cap = (long long)num_sms() * 32;
  sine_embed_kernel<<<(int)(nb > cap ? cap : nb), 256, 0, as_stream(stream)>>>(pos, w, h, dim_t, out, rows);
  return check_launch("ff3d_sine_embed");
}

extern "C" int ff3d_mha_core(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out,
                             int ldo, int B, int Nq, int heads, int d, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(d == 16 || d == 32, "mha_core: head dim %d unsupported (16, 32)", d);
  FF3D_REQUIRE(Nq >= 1 && Nq <= MHA_MAXQ, "mha_core: Nq=%d unsupported (<= %d)", Nq, MHA_MAXQ);
  size_t smem = ((size_t)2 * Nq * d + (size_t)4 * 32 * (d + 2)) * sizeof(float);
  FF3D_REQUIRE(smem <= 220 * 1024, "mha_core: K/V tile does not fit shared memory");
  FF3D_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 &&
                   ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                     reinterpret_cast<uintptr_t>(out)) & 15) == 0,
               "mha_core: q / k / v / out rows must be 16-byte aligned");
  // one lane per query, 32 queries per CTA, keys split over its 4 warps: B * heads * ceil(Nq / 32) CTAs (4 x 8 x 19 = 608)
  dim3 grid(B * heads, cdiv(Nq, 32));
  cudaStream_t st = as_stream(stream);
  // the shared-memory size depends on Nq: opt in to the device maximum once (thread-safe static initialisers)
  if (d == 16) {
    static const cudaError_t attr = cudaFuncSetAttribute(mha_core_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    (void)attr;
    mha_core_kernel<16><<<grid, 128, smem, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, Nq, heads);
  } else {
    static const cudaError_t attr = cudaFuncSetAttribute(mha_core_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    (void)attr;
    mha_core_kernel<32><<<grid, 128, smem, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, Nq, heads);
  }
  return check_launch("ff3d_mha_core");
}

extern "C" int ff3d_msda(const float* value, int ldv, int v_col0, long long v_bstride, const int* lvl_h,
                         const int* lvl_w, const int* lvl_start, int L, int P, const float* ref, float ref_w, float ref_h,
                         const float* offs, int ldoffs, const float* attw, int ldattw, float* out, int B, int Nq, int heads, int d,
                         ff3d_stream_t stream) {
  using namespace ff3d;
  Levels lv;
  FF3D_REQUIRE(fill_levels(&lv, lvl_h, lvl_w, lvl_start, L) == 0, "msda: 1..8 levels supported");
  long long total = (long long)B * Nq * heads * d;
  if (total <= 0) return FF3D_OK;
  msda_kernel<<<cdiv(total, 128), 128, 0, as_stream(stream)>>>(value, ldv, v_col0, v_bstride, lv, P, ref, ref_w, ref_h, offs,
                                                               ldoffs, attw, ldattw, out, B, Nq, heads, d);
  return check_launch("ff3d_msda");
}

extern "C" int ff3d_roi_sample(const float* query_box, int box_ld, const float* value, int ldv, long long v_bstride,
                               const int* lvl_h, const int* lvl_w, const int* lvl_start, int L, int C, int g,
                               float expand, float cell_x, float cell_y, float origin_x, float origin_y,
                               const float* roi_range4, float* out, int B, int Nq, ff3d_stream_t stream) {
  using namespace ff3d;
  Levels lv;
  FF3D_REQUIRE(fill_levels(&lv, lvl_h, lvl_w, lvl_start, L) == 0, "roi_sample: 1..8 levels supported");
  RoiP p{box_ld, g, C, expand, cell_x, cell_y, origin_x, origin_y, roi_range4[0], roi_range4[1], roi_range4[2],
         roi_range4[3]};
  if (B * Nq <= 0) return FF3D_OK;
  FF3D_REQUIRE(C % 4 == 0 && ldv % 4 == 0, "roi_sample: C and ldv must be multiples of 4");
  const long long items = (long long)B * Nq * g * g;
  long long nb = (items * 32 + 255) / 256, cap = (long long)num_sms() * 16;
  roi_sample_kernel<false><<<(int)(nb > cap ? cap : nb), 256, 0, as_stream(stream)>>>(query_box, p, value, ldv, v_bstride, lv, out,
                                                                                      nullptr, Nq, items, nullptr);
  return check_launch("ff3d_roi_sample");
}

// same sampling, output rows in split form [hi(K) | lo(K)] fp16 (K = L*g*g*C) for ff3d_tmagemm
extern "C" int ff3d_roi_sample_split(const float* query_box, int box_ld, const float* value, int ldv, long long v_bstride,
                                     const int* lvl_h, const int* lvl_w, const int* lvl_start, int L, int C, int g, float expand,
                                     float cell_x, float cell_y, float origin_x, float origin_y, const float* roi_range4,
                                     void* out_split, int B, int Nq, int* overflow_dev, ff3d_stream_t stream) {
  using namespace ff3d;
  Levels lv;
  FF3D_REQUIRE(fill_levels(&lv, lvl_h, lvl_w, lvl_start, L) == 0, "roi_sample: 1..8 levels supported");
  RoiP p{box_ld, g, C, expand, cell_x, cell_y, origin_x, origin_y, roi_range4[0], roi_range4[1], roi_range4[2],
         roi_range4[3]};
  if (B * Nq <= 0) return FF3D_OK;
  FF3D_REQUIRE(C % 4 == 0 && ldv % 4 == 0, "roi_sample: C and ldv must be multiples of 4");
  const long long items = (long long)B * Nq * g * g;
  long long nb = (items * 32 + 255) / 256, cap = (long long)num_sms() * 16;
  roi_sample_kernel<true><<<(int)(nb > cap ? cap : nb), 256, 0, as_stream(stream)>>>(
      query_box, p, value, ldv, v_bstride, lv, nullptr, static_cast<__half*>(out_split), Nq, items, overflow_dev);
  return check_launch("ff3d_roi_sample_split");
}

extern "C" int ff3d_head_update(float* pred, int ldp, float* query_pos, const float* prev, int ldprev, int rows,
                                ff3d_stream_t stream) {
  using namespace ff3d;
  if (rows <= 0) return FF3D_OK;
  head_update_kernel<<<cdiv(rows, 128), 128, 0, as_stream(stream)>>>(pred, ldp, query_pos, prev, ldprev, rows);
  return check_launch("ff3d_head_update");
}

extern "C" int ff3d_box_decode(const float* pred, int ldp, int cls_col, int has_vel, const float* query_score,
                               const int* query_label, int rows, int C, float cell_x, float cell_y, float origin_x,
                               float origin_y, const float* post_range6, float* boxes, float* scores, int* labels,
                               unsigned char* keep, ff3d_stream_t stream) {
  using namespace ff3d;
  if (rows <= 0) return FF3D_OK;
  DecP p;
  p.C = C; p.has_vel = has_vel; p.cls_col = cls_col; p.ldp = ldp;
  p.cell_x = cell_x; p.cell_y = cell_y; p.origin_x = origin_x; p.origin_y = origin_y;
  for (int i = 0; i < 6; ++i) p.pr[i] = post_range6[i];
  box_decode_kernel<<<cdiv(rows, 128), 128, 0, as_stream(stream)>>>(pred, query_score, query_label, p, boxes, scores,
                                                                    labels, keep, rows);
  return check_launch("ff3d_box_decode");
}

extern "C" int ff3d_class_select(const float* full, int ld_full, const int* label, const int* group_k, int n_groups, int tail,
                                 int num_classes, float* out, int ld_out, int rows, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(n_groups > 0 && n_groups <= 8 && num_classes > 0 && tail >= 0, "class_select: bad group description");
  if (rows <= 0) return FF3D_OK;
  ClsSelP p;
  p.n_groups = n_groups; p.tail = tail; p.nc = num_classes; p.ld_full = ld_full; p.ld_out = ld_out;
  int out_cols = tail;
  for (int g = 0; g < 8; ++g) { p.k[g] = g < n_groups ? group_k[g] : 0; out_cols += p.k[g]; }
  FF3D_REQUIRE(out_cols <= ld_out, "class_select: output row too narrow");
  class_select_kernel<<<cdiv((long long)rows * out_cols, 256), 256, 0, as_stream(stream)>>>(full, label, p, out, rows, out_cols);
  return check_launch("ff3d_class_select");
}
