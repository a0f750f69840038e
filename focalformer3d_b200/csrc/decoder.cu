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
template <int D>
__global__ void __launch_bounds__(256) mha_core_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k,
                                                       int ldk, const float* __restrict__ v, int ldv,
                                                       float* __restrict__ out, int ldo, int Nq, int heads) {
  extern __shared__ float sm[];
  float* Ks = sm;                      // [Nq][D+1]
  float* Vs = sm + (size_t)Nq * (D + 1);
  const int bh = blockIdx.x;
  const int b = bh / heads, h = bh - b * heads;
  const size_t row0 = (size_t)b * Nq;
  for (int e = threadIdx.x; e < Nq * D; e += blockDim.x) {
    int j = e / D, c = e - j * D;
    Ks[j * (D + 1) + c] = k[(row0 + j) * ldk + h * D + c];
    Vs[j * (D + 1) + c] = v[(row0 + j) * ldv + h * D + c];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const float scale = rsqrtf((float)D);
  for (int qi = blockIdx.y * nwarp + warp; qi < Nq; qi += gridDim.y * nwarp) {
    float qr[D];
#pragma unroll
    for (int c = 0; c < D; ++c) qr[c] = q[(row0 + qi) * ldq + h * D + c] * scale;
    float sc[MHA_MAXQ / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < MHA_MAXQ / 32; ++t) {
      int j = lane + t * 32;
      float s = -INFINITY;
      if (j < Nq) {
        s = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) s = fmaf(qr[c], Ks[j * (D + 1) + c], s);
      }
      sc[t] = s;
      mx = fmaxf(mx, s);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    float acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = 0.f;
#pragma unroll
    for (int t = 0; t < MHA_MAXQ / 32; ++t) {
      int j = lane + t * 32;
      if (j < Nq) {
        float pj = expf(sc[t] - mx);
        sum += pj;
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = fmaf(pj, Vs[j * (D + 1) + c], acc[c]);
      }
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
    for (int c = 0; c < D; ++c)
      for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    float inv = 1.f / sum;
    if (lane < D) {
      float r = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c)
        if (c == lane) r = acc[c];
      out[(row0 + qi) * ldo + h * D + lane] = r * inv;
    }
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
// SPLIT: the row is written as fp16 [hi(K) | lo(K)] (K = L*g*g*C) for the TMA-fed roi_mlp.0 GEMM instead of fp32
template <bool SPLIT>
__global__ void roi_sample_kernel(const float* __restrict__ qbox, RoiP p, const float* __restrict__ value, int ldv,
                                  long long v_bstride, Levels lv, float* __restrict__ out, __half* __restrict__ outs, int Nq,
                                  int* overflow) {
  long long bq = blockIdx.x;
  int b = (int)(bq / Nq);
  const float* qb = qbox + bq * p.box_ld;
  // decode_box (transfusion_bbox_coder.py:54-69) on (centre, expanded log-dims, sin/cos)
  float cx = qb[0] * p.cell_x + p.origin_x;
  float cy = qb[1] * p.cell_y + p.origin_y;
  float w = expf(qb[3] * p.expand), l = expf(qb[4] * p.expand);
  float yaw = atan2f(qb[6], qb[7]);
  float sn = sinf(yaw), cs = cosf(yaw);
  const float* vb = value + (long long)b * v_bstride * ldv;
  int G = p.g * p.g;
  const long long K = (long long)lv.L * G * p.C;
  float* orow = SPLIT ? nullptr : out + bq * K;
  __half* srow = SPLIT ? outs + bq * 2 * K : nullptr;
  bool ovf = false;
  for (int n = 0; n < G; ++n) {
    int i = n / p.g, j = n - i * p.g;
    float gx = ((float)i + 0.5f) / (float)p.g * w - w / 2.f;
    float gy = ((float)j + 0.5f) / (float)p.g * l - l / 2.f;
    // [upstream] mmdet3d v0.17.1 rotation_3d_in_axis(axis=2): x' = x cos + y sin, y' = -x sin + y cos
    float wx = gx * cs + gy * sn + cx;
    float wy = -gx * sn + gy * cs + cy;
    float nx = (wx - p.rx0) / (p.rx1 - p.rx0) * 2.f - 1.f;
    float ny = (wy - p.ry0) / (p.ry1 - p.ry0) * 2.f - 1.f;
    nx = fminf(fmaxf(nx, -2.f), 2.f);
    ny = fminf(fmaxf(ny, -2.f), 2.f);
    for (int lvl = 0; lvl < lv.L; ++lvl) {
      int H = lv.h[lvl], W = lv.w[lvl];
      float px = ((nx + 1.f) * (float)W - 1.f) / 2.f;
      float py = ((ny + 1.f) * (float)H - 1.f) / 2.f;
      const float* base = vb + (long long)lv.start[lvl] * ldv;
      for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
        const float v = bilinear_zero(base, ldv, H, W, px, py, c);
        const long long k = ((long long)lvl * G + n) * p.C + c;
        if (SPLIT) {
          const float cv = fminf(fmaxf(v, -F16_MAX), F16_MAX);
          ovf = ovf || cv != v;
          const __half hi = __float2half_rn(cv);
          srow[k] = hi;
          srow[K + k] = __float2half_rn(fminf(fmaxf((cv - __half2float(hi)) * 2048.f, -F16_MAX), F16_MAX));
        } else {
          orow[k] = v;
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
  size_t smem = (size_t)2 * Nq * (d + 1) * sizeof(float);
  FF3D_REQUIRE(smem <= 220 * 1024, "mha_core: K/V tile does not fit shared memory");
  // query chunks per (batch, head): ~64 queries each, but never more CTAs than fit in ONE wave (two 81 KB CTAs per SM)
  // -- 4 x 8 x 10 = 320 CTAs on 296 slots ran a nearly empty second wave
  int ny = cdiv(Nq, 64);
  const int one_wave = (2 * num_sms()) / (B * heads);
  if (ny > one_wave && one_wave >= 1) ny = one_wave;
  dim3 grid(B * heads, ny);
  cudaStream_t st = as_stream(stream);
  // the shared-memory size depends on Nq: opt in to the device maximum once (thread-safe static initialisers)
  FF3D_REQUIRE(smem <= 227 * 1024, "mha_core: Nq=%d needs %zu bytes of shared memory", Nq, smem);
  if (d == 16) {
    static const cudaError_t attr = cudaFuncSetAttribute(mha_core_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    (void)attr;
    mha_core_kernel<16><<<grid, 256, smem, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, Nq, heads);
  } else {
    static const cudaError_t attr = cudaFuncSetAttribute(mha_core_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    (void)attr;
    mha_core_kernel<32><<<grid, 256, smem, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, Nq, heads);
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
  roi_sample_kernel<false><<<B * Nq, 128, 0, as_stream(stream)>>>(query_box, p, value, ldv, v_bstride, lv, out, nullptr, Nq, nullptr);
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
  roi_sample_kernel<true><<<B * Nq, 128, 0, as_stream(stream)>>>(query_box, p, value, ldv, v_bstride, lv, nullptr,
                                                                static_cast<__half*>(out_split), Nq, overflow_dev);
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
