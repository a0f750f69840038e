// Small HBM-bound NHWC helpers: depthwise 3x3, LayerNorm, row adds.
#include "common.cuh"

namespace ff3d {

// thread = (pixel, 4 channels); weights [9, C]; 9 float4 loads per output float4 (neighbours hit L1/L2)
__global__ void dwconv3x3_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w,
                                 const float* __restrict__ bias, float* __restrict__ y, int ldy, int B, int H, int W,
                                 int C, int act) {
  int c4n = C >> 2;
  long long total = (long long)B * H * W * c4n;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int c4 = (int)(e % c4n);
    long long pix = e / c4n;
    int xw = (int)(pix % W);
    long long r = pix / W;
    int yh = (int)(r % H);
    int b = (int)(r / H);
    int c = c4 * 4;
    float4 acc = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      int iy = yh + ky - 1;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        int ix = xw + kx - 1;
        if (ix < 0 || ix >= W) continue;
        float4 v = __ldg(reinterpret_cast<const float4*>(x + (((long long)b * H + iy) * W + ix) * ldx + c));
        float4 k = __ldg(reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C + c));
        acc.x = fmaf(v.x, k.x, acc.x); acc.y = fmaf(v.y, k.y, acc.y);
        acc.z = fmaf(v.z, k.z, acc.z); acc.w = fmaf(v.w, k.w, acc.w);
      }
    }
    acc.x = apply_act(acc.x, act); acc.y = apply_act(acc.y, act);
    acc.z = apply_act(acc.z, act); acc.w = apply_act(acc.w, act);
    *reinterpret_cast<float4*>(y + pix * ldy + c) = acc;
  }
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
  long long total = (long long)B * H * W * (C / 4);
  dwconv3x3_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(x, ldx, w, bias, y, ldy, B, H, W, C, act);
  return check_launch("ff3d_dwconv3x3");
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
