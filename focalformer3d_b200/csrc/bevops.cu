// Small HBM-bound NHWC helpers: depthwise 3x3, LayerNorm, row adds.
#include "tc_common.cuh"

namespace ff3d {

// thread = (4 consecutive pixels of a row, 4 * V channels); weights [9, C].  The 3 x 6 input window of the four outputs is
// walked row by row: 18 input vectors and 9 weight vectors per 4 outputs instead of 36 + 36 with one pixel per thread.
// SPLIT: the output is written in split form (fp16 [hi(C) | lo(C)] rows, for a TMA-fed 1x1 conv) instead of fp32; a
// thread then owns 8 channels so that both planes get 16-byte stores
template <bool SPLIT>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y, int ldy, int B, int H,
                                                        int W, int C, int act, __half* __restrict__ ys, int* overflow) {
  constexpr int V = SPLIT ? 2 : 1;                       // float4 vectors per thread
  constexpr int XB = 4;                                  // pixels per thread
  bool ovf = false;
  const int cvn = C / (4 * V);
  const int xbn = (W + XB - 1) / XB;
  const long long total = (long long)B * H * xbn * cvn;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(e % cvn);
    long long r = e / cvn;
    const int x0 = (int)(r % xbn) * XB;
    r /= xbn;
    const int yh = (int)(r % H);
    const int b = (int)(r / H);
    const int c = cv * 4 * V;
    float4 acc[XB][V];
#pragma unroll
    for (int u = 0; u < V; ++u) {
      const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias + c + 4 * u)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int px = 0; px < XB; ++px) acc[px][u] = bv;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = yh + ky - 1;
      if (iy < 0 || iy >= H) continue;
      const float* rowp = x + ((long long)b * H + iy) * W * ldx + c;
      float4 in[XB + 2][V];
#pragma unroll
      for (int j = 0; j < XB + 2; ++j) {
        const int ix = x0 + j - 1;
        const bool ok = ix >= 0 && ix < W;
#pragma unroll
        for (int u = 0; u < V; ++u)
          in[j][u] = ok ? __ldg(reinterpret_cast<const float4*>(rowp + (long long)ix * ldx + 4 * u)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
        for (int u = 0; u < V; ++u) {
          const float4 k = __ldg(reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C + c + 4 * u));
#pragma unroll
          for (int px = 0; px < XB; ++px) {
            const float4 v = in[px + kx][u];
            acc[px][u].x = fmaf(v.x, k.x, acc[px][u].x); acc[px][u].y = fmaf(v.y, k.y, acc[px][u].y);
            acc[px][u].z = fmaf(v.z, k.z, acc[px][u].z); acc[px][u].w = fmaf(v.w, k.w, acc[px][u].w);
          }
        }
      }
    }
#pragma unroll
    for (int px = 0; px < XB; ++px) {
      if (x0 + px >= W) break;
      const long long pix = ((long long)b * H + yh) * W + x0 + px;
#pragma unroll
      for (int u = 0; u < V; ++u) {
        acc[px][u].x = apply_act(acc[px][u].x, act); acc[px][u].y = apply_act(acc[px][u].y, act);
        acc[px][u].z = apply_act(acc[px][u].z, act); acc[px][u].w = apply_act(acc[px][u].w, act);
      }
      if constexpr (SPLIT) {
        uint32_t hw[4], lw[4];
        split_f16x4(acc[px][0], hw[0], hw[1], lw[0], lw[1], ovf);
        split_f16x4(acc[px][V - 1], hw[2], hw[3], lw[2], lw[3], ovf);
        *reinterpret_cast<uint4*>(ys + pix * 2 * C + c) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(ys + pix * 2 * C + C + c) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      } else {
        *reinterpret_cast<float4*>(y + pix * ldy + c) = acc[px][0];
      }
    }
  }
  if (SPLIT && ovf && overflow) atomicOr(overflow, 1);
}

// one warp per row, C <= 1024, C % 32 == 0 handled generally with a strided loop; two-pass (mean, then variance)
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float* __restrict__ y, int rows, int C, float eps) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + (size_t)warp * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float mean = s / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { float d = xr[c] - mean; v = fmaf(d, d, v); }
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  float rstd = rsqrtf(v / (float)C + eps);
  float* yr = y + (size_t)warp * C;
  for (int c = lane; c < C; c += 32) yr[c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
}

__global__ void add_rows_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ y,
                                long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 u = a[i], v = b[i];
    y[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
}

__global__ void add_bcast_rows_kernel(const float4* __restrict__ a, const float4* __restrict__ p, float4* __restrict__ y,
                                      int B, long long per4) {
  long long n4 = per4 * B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 u = a[i], v = p[i % per4];
    y[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
}


// Local-context attention (locatt_ops similar -> softmax -> weighting, encoder_utils.py:155-163) fused: one warp per
// pixel, lane l owns channels 4l..4l+3 of every 128-channel block.  K*K <= 96 scores live three per lane; neighbours
// outside the map score 0 and still take part in the soft-max (the reference pads the similarity with zeros), but
// contribute no value.  q/k/v rows of neighbouring pixels are re-read from L1/L2 (81x reuse inside a CTA's 8-pixel row).
template <int NV>
__global__ void __launch_bounds__(256) local_attention_kernel(const float* __restrict__ q, int ldq,
                                                              const float* __restrict__ k, int ldk,
                                                              const float* __restrict__ v, int ldv, float* __restrict__ y,
                                                              int ldy, int B, int H, int W, int K, float scale) {
  const int lane = threadIdx.x & 31;
  const long long pix = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (pix >= (long long)B * H * W) return;
  const int x0 = (int)(pix % W);
  const long long r = pix / W;
  const int y0 = (int)(r % H);
  const long long b = r / H;
  const int R = K >> 1, KK = K * K;
  float4 qv[NV];
#pragma unroll
  for (int u = 0; u < NV; ++u) qv[u] = __ldg(reinterpret_cast<const float4*>(q + pix * ldq + u * 128 + lane * 4));
  float sc[3] = {0.f, 0.f, 0.f};
  for (int kk = 0; kk < KK; ++kk) {
    const int ny = y0 + kk / K - R, nx = x0 + kk % K - R;
    float s = 0.f;
    if (ny >= 0 && ny < H && nx >= 0 && nx < W) {          // warp-uniform
      const float* kp = k + ((b * H + ny) * W + nx) * ldk + lane * 4;
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(kp + u * 128));
        s = fmaf(qv[u].x, t.x, s); s = fmaf(qv[u].y, t.y, s); s = fmaf(qv[u].z, t.z, s); s = fmaf(qv[u].w, t.w, s);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    if ((kk & 31) == lane) {                               // static register indices (no local-memory array)
      if (kk < 32) sc[0] = s * scale; else if (kk < 64) sc[1] = s * scale; else sc[2] = s * scale;
    }
  }
  float m = -INFINITY;
#pragma unroll
  for (int u = 0; u < 3; ++u) if (u * 32 + lane < KK) m = fmaxf(m, sc[u]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    sc[u] = (u * 32 + lane < KK) ? expf(sc[u] - m) : 0.f;
    sum += sc[u];
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  float4 acc[NV];
#pragma unroll
  for (int u = 0; u < NV; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int kk = 0; kk < KK; ++kk) {
    const float w0 = __shfl_sync(0xffffffffu, sc[0], kk & 31);
    const float w1 = __shfl_sync(0xffffffffu, sc[1], kk & 31);
    const float w2 = __shfl_sync(0xffffffffu, sc[2], kk & 31);
    const float w = (kk < 32 ? w0 : (kk < 64 ? w1 : w2)) * inv;
    const int ny = y0 + kk / K - R, nx = x0 + kk % K - R;
    if (ny >= 0 && ny < H && nx >= 0 && nx < W) {
      const float* vp = v + ((b * H + ny) * W + nx) * ldv + lane * 4;
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(vp + u * 128));
        acc[u].x = fmaf(w, t.x, acc[u].x); acc[u].y = fmaf(w, t.y, acc[u].y);
        acc[u].z = fmaf(w, t.z, acc[u].z); acc[u].w = fmaf(w, t.w, acc[u].w);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < NV; ++u) *reinterpret_cast<float4*>(y + pix * ldy + u * 128 + lane * 4) = acc[u];
}

// Tiled variant for C = 128 (the shipped configs): one CTA = 8x8 output pixels.  The (8+K-1)^2 key rows are staged once
// in shared memory (row stride 132 floats, tile width 17: the lane <-> neighbour mapping below is then bank-conflict
// free), zero-filled outside the map -- a zero key scores exactly 0, a zero value adds nothing, which IS the reference's
// border rule, so the inner loops carry no bounds checks.  Scores: lane l owns neighbours l, l+32, l+64 and runs the whole
// 128-channel dot product itself (q broadcast from shared memory; no shuffle reductions).  Then the same buffer is
// re-filled with the value rows and the weighted sum runs with lanes over channels.  Global traffic drops from
// 2*K*K rows per pixel to ~9 rows per pixel.
constexpr int LA_T = 8, LA_C = 128, LA_LD = 132;
__global__ void __launch_bounds__(512, 1) local_attention_tiled_kernel(const float* __restrict__ q, int ldq,
                                                                    const float* __restrict__ k, int ldk,
                                                                    const float* __restrict__ v, int ldv,
                                                                    float* __restrict__ y, int ldy, int H, int W, int K,
                                                                    float scale) {
  extern __shared__ __align__(16) float la_smem[];
  const int R = K >> 1, KK = K * K, HW = LA_T + K - 1, HWP = HW + 1;      // halo tile is HW x HWP (padded width)
  float* qs = la_smem;                                                   // [64][128]
  float* ks = la_smem + LA_T * LA_T * LA_C;                              // [HW * HWP][132]
  const int tiles_x = (W + LA_T - 1) / LA_T, tiles_y = (H + LA_T - 1) / LA_T;
  const int b = blockIdx.x / (tiles_x * tiles_y);
  const int trem = blockIdx.x - b * tiles_x * tiles_y;
  const int y0 = (trem / tiles_x) * LA_T, x0 = (trem % tiles_x) * LA_T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long img = (long long)b * H * W;
  auto stage_halo = [&](const float* src, int ld) {
    for (int e = tid; e < HW * HW * (LA_C / 4); e += blockDim.x) {
      const int c4 = e & (LA_C / 4 - 1), pix = e / (LA_C / 4);
      const int hy = pix / HW, hx = pix - hy * HW;
      const int gy = y0 + hy - R, gx = x0 + hx - R;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) t = __ldg(reinterpret_cast<const float4*>(src + (img + (long long)gy * W + gx) * ld + c4 * 4));
      *reinterpret_cast<float4*>(ks + (hy * HWP + hx) * LA_LD + c4 * 4) = t;
    }
  };
  for (int e = tid; e < LA_T * LA_T * (LA_C / 4); e += blockDim.x) {
    const int c4 = e & (LA_C / 4 - 1), pix = e / (LA_C / 4);
    const int gy = y0 + pix / LA_T, gx = x0 + pix % LA_T;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gy < H && gx < W) t = __ldg(reinterpret_cast<const float4*>(q + (img + (long long)gy * W + gx) * ldq + c4 * 4));
    *reinterpret_cast<float4*>(qs + pix * LA_C + c4 * 4) = t;
  }
  stage_halo(k, ldk);
  __syncthreads();
  // ---- scores + soft-max: 4 pixels per warp, 3 neighbours per lane
  constexpr int PPW = LA_T * LA_T / 16;
  float wgt[PPW][3];
  int nb[3];                                                             // halo-row offset of my neighbour, minus the pixel's
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int kk = r * 32 + lane;
    nb[r] = kk < KK ? ((kk / K) * HWP + kk % K) * LA_LD : 0;
  }
#pragma unroll
  for (int pp = 0; pp < PPW; ++pp) {
    const int p = warp * PPW + pp;
    const float* qp = qs + p * LA_C;
    const float* kp = ks + ((p / LA_T) * HWP + p % LA_T) * LA_LD;
    float a[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r) a[r][0] = a[r][1] = a[r][2] = a[r][3] = 0.f;
#pragma unroll 4
    for (int c = 0; c < LA_C; c += 4) {
      const float4 qv = *reinterpret_cast<const float4*>(qp + c);        // same address in every lane: broadcast
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float4 kv = *reinterpret_cast<const float4*>(kp + nb[r] + c);
        a[r][0] = fmaf(qv.x, kv.x, a[r][0]); a[r][1] = fmaf(qv.y, kv.y, a[r][1]);
        a[r][2] = fmaf(qv.z, kv.z, a[r][2]); a[r][3] = fmaf(qv.w, kv.w, a[r][3]);
      }
    }
    float s[3], m = -INFINITY;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      s[r] = (r * 32 + lane < KK) ? ((a[r][0] + a[r][1]) + (a[r][2] + a[r][3])) * scale : -INFINITY;
      m = fmaxf(m, s[r]);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) { s[r] = expf(s[r] - m); sum += s[r]; }   // expf(-inf) = 0 for the unused slots
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
#pragma unroll
    for (int r = 0; r < 3; ++r) wgt[pp][r] = s[r] * inv;
  }
  __syncthreads();
  stage_halo(v, ldv);
  __syncthreads();
  // ---- weighted sum of the value rows: lane l owns channels 4l .. 4l+3
#pragma unroll
  for (int pp = 0; pp < PPW; ++pp) {
    const int p = warp * PPW + pp;
    const int gy = y0 + p / LA_T, gx = x0 + p % LA_T;
    const float* vp = ks + ((p / LA_T) * HWP + p % LA_T) * LA_LD + lane * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int n = KK - r * 32 < 32 ? KK - r * 32 : 32;
      for (int l = 0; l < n; ++l) {
        const float w = __shfl_sync(0xffffffffu, wgt[pp][r], l);
        const int kk = r * 32 + l;
        const float4 t = *reinterpret_cast<const float4*>(vp + ((kk / K) * HWP + kk % K) * LA_LD);
        acc.x = fmaf(w, t.x, acc.x); acc.y = fmaf(w, t.y, acc.y); acc.z = fmaf(w, t.z, acc.z); acc.w = fmaf(w, t.w, acc.w);
      }
    }
    if (gy < H && gx < W) *reinterpret_cast<float4*>(y + (img + (long long)gy * W + gx) * ldy + lane * 4) = acc;
  }
}

static inline int grid_for(long long work, int threads) {
  long long nb = (work + threads - 1) / threads;
  long long cap = (long long)num_sms() * 32;
  return (int)(nb < 1 ? 1 : (nb > cap ? cap : nb));
}

}  // namespace ff3d

extern "C" int ff3d_dwconv3x3(const float* x, int ldx, const float* w, const float* bias, float* y, int ldy, int B, int H,
                              int W, int C, int act, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "dwconv3x3: C/ldx/ldy must be multiples of 4");
  long long total = (long long)B * H * ((W + 3) / 4) * (C / 4);
  dwconv3x3_kernel<false><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(x, ldx, w, bias, y, ldy, B, H, W, C, act, nullptr,
                                                                                nullptr);
  return check_launch("ff3d_dwconv3x3");
}

// same conv, output in split form only: ys [B*H*W, 2C] fp16 [hi | lo]
extern "C" int ff3d_dwconv3x3_split(const float* x, int ldx, const float* w, const float* bias, void* ys, int B, int H, int W,
                                    int C, int act, int* overflow_dev, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(C % 8 == 0 && ldx % 4 == 0, "dwconv3x3_split: C must be a multiple of 8, ldx of 4");
  long long total = (long long)B * H * ((W + 3) / 4) * (C / 8);
  dwconv3x3_kernel<true><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(x, ldx, w, bias, nullptr, 0, B, H, W, C, act,
                                                                               static_cast<__half*>(ys), overflow_dev);
  return check_launch("ff3d_dwconv3x3_split");
}

extern "C" int ff3d_layernorm(const float* x, const float* gamma, const float* beta, float* y, int rows, int C, float eps,
                              ff3d_stream_t stream) {
  using namespace ff3d;
  if (rows <= 0) return FF3D_OK;
  layernorm_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, as_stream(stream)>>>(x, gamma, beta, y, rows, C, eps);
  return check_launch("ff3d_layernorm");
}

extern "C" int ff3d_add_rows(const float* a, const float* b, float* y, long long n, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(n % 4 == 0, "add_rows: n must be a multiple of 4");
  add_rows_kernel<<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), reinterpret_cast<float4*>(y), n / 4);
  return check_launch("ff3d_add_rows");
}

extern "C" int ff3d_add_bcast_rows(const float* a, const float* p, float* y, int B, long long rows, int C,
                                   ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(C % 4 == 0, "add_bcast_rows: C must be a multiple of 4");
  long long per4 = rows * C / 4;
  add_bcast_rows_kernel<<<grid_for(per4 * B, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(p), reinterpret_cast<float4*>(y), B, per4);
  return check_launch("ff3d_add_bcast_rows");
}

// y = a + p in split form only: ys [B*rows, 2C] fp16 [hi | lo] (the value_proj input of the deformable decoder)
__global__ void add_bcast_rows_split_kernel(const float4* __restrict__ a, const float4* __restrict__ p, __half* __restrict__ ys,
                                            int B, long long per4, int C, int* overflow) {
  using namespace ff3d;
  long long n4 = per4 * B;
  bool ovf = false;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 u = a[i], v = p[i % per4];
    const float4 s = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
    uint32_t h01, h23, l01, l23;
    split_f16x4(s, h01, h23, l01, l23, ovf);
    const long long e = i * 4;
    const long long row = (e >> 31) ? e / C : (long long)((unsigned)e / (unsigned)C);      // 32-bit division when it fits
    const int col = (int)(e - row * C);
    *reinterpret_cast<uint2*>(ys + row * 2 * C + col) = make_uint2(h01, h23);
    *reinterpret_cast<uint2*>(ys + row * 2 * C + C + col) = make_uint2(l01, l23);
  }
  if (ovf && overflow) atomicOr(overflow, 1);
}

extern "C" int ff3d_add_bcast_rows_split(const float* a, const float* p, void* ys, int B, long long rows, int C,
                                         int* overflow_dev, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(C % 8 == 0, "add_bcast_rows_split: C must be a multiple of 8");
  long long per4 = rows * C / 4;
  add_bcast_rows_split_kernel<<<grid_for(per4 * B, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(p), static_cast<__half*>(ys), B, per4, C, overflow_dev);
  return check_launch("ff3d_add_bcast_rows_split");
}

extern "C" int ff3d_local_attention(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* y,
                                    int ldy, int B, int H, int W, int C, int K, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE((C == 128 || C == 256) && (K & 1) == 1 && K * K <= 96, "local_attention: C must be 128 or 256, K odd, K*K <= 96");
  FF3D_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldy % 4 == 0, "local_attention: row strides must be multiples of 4");
  long long pix = (long long)B * H * W;
  if (pix == 0) return FF3D_OK;
  float scale = 1.0f / sqrtf((float)C);
  int blocks = cdiv(pix * 32, 256);
  if (C == 128) {
    const int hw = LA_T + K - 1;
    const size_t smem = sizeof(float) * ((size_t)LA_T * LA_T * LA_C + (size_t)hw * (hw + 1) * LA_LD);
    if (smem <= 227 * 1024) {
      // one-time opt-in to > 48 KB dynamic shared memory: function-local static initialiser (thread-safe, C++11)
      static const cudaError_t attr =
          cudaFuncSetAttribute(local_attention_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      (void)attr;
      const int tiles = cdiv(H, LA_T) * cdiv(W, LA_T);
      local_attention_tiled_kernel<<<B * tiles, 512, smem, as_stream(stream)>>>(q, ldq, k, ldk, v, ldv, y, ldy, H, W, K, scale);
      return check_launch("ff3d_local_attention");
    }
    local_attention_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(q, ldq, k, ldk, v, ldv, y, ldy, B, H, W, K, scale);
  } else
    local_attention_kernel<2><<<blocks, 256, 0, as_stream(stream)>>>(q, ldq, k, ldk, v, ldv, y, ldy, B, H, W, K, scale);
  return check_launch("ff3d_local_attention");
}
