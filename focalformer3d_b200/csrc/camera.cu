// Camera branch helpers (NHWC fp32): image repack, 3x3/s2 max-pool, FPN nearest top-down add, Lift-Splat-Shoot splat.
// The convolutions of ResNet-50 / FPN / depthnet / bevencode run through the implicit-GEMM kernels (tcgemm.cu).
#include "common.cuh"

namespace ff3d {

static inline int cam_grid(long long work, int threads) {
  long long nb = (work + threads - 1) / threads;
  long long cap = (long long)num_sms() * 32;
  return (int)(nb < 1 ? 1 : (nb > cap ? cap : nb));
}

// [n, C, H, W] planar -> [n, H, W, ld] interleaved, channels >= C zero-filled; thread = one pixel (planar reads are
// coalesced across the warp, the 32-byte interleaved rows are written whole)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int C, long long HW,
                                    int ld) {
  long long total = (long long)n * HW;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    long long img = e / HW, pix = e - img * HW;
    const float* src = x + img * C * HW + pix;
    float* dst = y + e * ld;
    for (int c = 0; c < ld; c += 4) {
      float4 v;
      v.x = (c + 0 < C) ? __ldg(src + (long long)(c + 0) * HW) : 0.f;
      v.y = (c + 1 < C) ? __ldg(src + (long long)(c + 1) * HW) : 0.f;
      v.z = (c + 2 < C) ? __ldg(src + (long long)(c + 2) * HW) : 0.f;
      v.w = (c + 3 < C) ? __ldg(src + (long long)(c + 3) * HW) : 0.f;
      *reinterpret_cast<float4*>(dst + c) = v;
    }
  }
}

// 3x3 stride-2 pad-1 max-pool (padding behaves as -inf, like torch); thread = (output pixel, 4 channels)
__global__ void maxpool3x3s2_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int H, int W, int Ho,
                                    int Wo, int C) {
  int c4n = C >> 2;
  long long total = (long long)n * Ho * Wo * c4n;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int c = (int)(e % c4n) * 4;
    long long pix = e / c4n;
    int ox = (int)(pix % Wo);
    long long r = pix / Wo;
    int oy = (int)(r % Ho);
    long long b = r / Ho;
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      int iy = oy * 2 - 1 + ky;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        int ix = ox * 2 - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        float4 v = __ldg(reinterpret_cast<const float4*>(x + ((b * H + iy) * W + ix) * C + c));
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    *reinterpret_cast<float4*>(y + pix * C + c) = m;
  }
}

// dst[n, y, x, :] += src[n, min(floor(y * Hs/Hd), Hs-1), min(floor(x * Ws/Wd), Ws-1), :]   (F.interpolate 'nearest')
__global__ void upsample_add_kernel(float* __restrict__ dst, const float* __restrict__ src, int n, int Hd, int Wd,
                                    int Hs, int Ws, int C, float sy, float sx) {
  int c4n = C >> 2;
  long long total = (long long)n * Hd * Wd * c4n;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int c = (int)(e % c4n) * 4;
    long long pix = e / c4n;
    int xd = (int)(pix % Wd);
    long long r = pix / Wd;
    int yd = (int)(r % Hd);
    long long b = r / Hd;
    int ys = min((int)floorf(__fmul_rn((float)yd, sy)), Hs - 1);
    int xs = min((int)floorf(__fmul_rn((float)xd, sx)), Ws - 1);
    float4 a = *reinterpret_cast<const float4*>(dst + pix * C + c);
    float4 v = __ldg(reinterpret_cast<const float4*>(src + ((b * Hs + ys) * Ws + xs) * C + c));
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    *reinterpret_cast<float4*>(dst + pix * C + c) = a;
  }
}

struct SplatParams {
  const float* dn; int ld;        // depthnet output [n_img, fH, fW, ld]: cols [0, 64) context, [64, 64 + D) depth logits
  const float* frustum;           // [D, fH, fW, 3] (u, v, depth) -- the checkpoint's LiftSplatShoot.frustum
  const float* rots;              // [n_img, 9] row-major inverse-projection rotation part
  const float* trans;             // [n_img, 3]
  float* bev;                     // [B, ny, nx, nz * 64]
  int n_img, cams, D, fH, fW;
  float lo[3], dx[3];             // voxel = trunc((p - lo) / dx), lo = bx - dx/2
  int nx, ny, nz;
};

constexpr int kCamC = 64;

// One warp per feature-map pixel.  Depth soft-max over D <= 64 logits held two per lane; then the warp walks the depth
// bins two at a time: half-warp h owns bin 2i+h, its 16 lanes each own 4 context channels -> one 256-byte vector
// reduction (red.global.add.v4.f32) per (pixel, bin) into channel block gz*64 of the BEV cell.  The geometry uses
// individually rounded fp32 mul/add in a fixed order (oracle/camera.py get_geometry) so the cell indices are
// bit-identical; only the order of the per-cell additions differs from the oracle.
__global__ void __launch_bounds__(256) lss_splat_kernel(SplatParams p) {
  int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  long long n_pix = (long long)p.n_img * p.fH * p.fW;
  if (warp >= n_pix) return;
  int hw = p.fH * p.fW;
  int img = (int)(warp / hw);
  int pix = (int)(warp - (long long)img * hw);
  int b = img / p.cams;
  const float* row = p.dn + warp * p.ld;
  float l0 = (lane < p.D) ? __ldg(row + kCamC + lane) : -INFINITY;
  float l1 = (lane + 32 < p.D) ? __ldg(row + kCamC + 32 + lane) : -INFINITY;
  float m = fmaxf(l0, l1);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float e0 = (lane < p.D) ? expf(l0 - m) : 0.f;
  float e1 = (lane + 32 < p.D) ? expf(l1 - m) : 0.f;
  float s = e0 + e1;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float p0 = e0 / s, p1 = e1 / s;
  int half = lane >> 4, q = lane & 15;
  float4 f = __ldg(reinterpret_cast<const float4*>(row + q * 4));
  const float* R = p.rots + img * 9;
  const float* T = p.trans + img * 3;
  float r00 = __ldg(R + 0), r01 = __ldg(R + 1), r02 = __ldg(R + 2), r10 = __ldg(R + 3), r11 = __ldg(R + 4),
        r12 = __ldg(R + 5), r20 = __ldg(R + 6), r21 = __ldg(R + 7), r22 = __ldg(R + 8);
  float t0 = __ldg(T + 0), t1 = __ldg(T + 1), t2 = __ldg(T + 2);
  int iters = (p.D + 1) >> 1;
  for (int i = 0; i < iters; ++i) {
    int d = 2 * i + half;
    float pa = __shfl_sync(0xffffffffu, p0, d & 31);
    float pb = __shfl_sync(0xffffffffu, p1, d & 31);
    if (d >= p.D) continue;
    float pd = d < 32 ? pa : pb;
    const float* fr = p.frustum + ((long long)d * hw + pix) * 3;
    float fu = __ldg(fr), fv = __ldg(fr + 1), fd = __ldg(fr + 2);
    float px = __fmul_rn(fu, fd), py = __fmul_rn(fv, fd), pz = fd;
    float gx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r00, px), __fmul_rn(r01, py)), __fmul_rn(r02, pz)), t0);
    float gy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r10, px), __fmul_rn(r11, py)), __fmul_rn(r12, pz)), t1);
    float gz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r20, px), __fmul_rn(r21, py)), __fmul_rn(r22, pz)), t2);
    float vx = __fdiv_rn(__fsub_rn(gx, p.lo[0]), p.dx[0]);
    float vy = __fdiv_rn(__fsub_rn(gy, p.lo[1]), p.dx[1]);
    float vz = __fdiv_rn(__fsub_rn(gz, p.lo[2]), p.dx[2]);
    // .long() truncates towards zero: (-1, 0) lands in cell 0 like the reference (lss.py:335-343)
    if (!(vx > -1.f && vx < (float)p.nx && vy > -1.f && vy < (float)p.ny && vz > -1.f && vz < (float)p.nz)) continue;
    int ix = (int)vx, iy = (int)vy, iz = (int)vz;
    float* cell = p.bev + ((((long long)b * p.ny + iy) * p.nx + ix) * p.nz + iz) * kCamC + q * 4;
    atomicAdd(reinterpret_cast<float4*>(cell), make_float4(pd * f.x, pd * f.y, pd * f.z, pd * f.w));
  }
}

}  // namespace ff3d

extern "C" int ff3d_nchw_to_nhwc(const float* x, float* y, int n, int C, int H, int W, int ld, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(ld % 4 == 0 && ld >= C && n >= 0, "nchw_to_nhwc: ld must be a multiple of 4 and >= C");
  long long total = (long long)n * H * W;
  if (total == 0) return FF3D_OK;
  nchw_to_nhwc_kernel<<<cam_grid(total, 256), 256, 0, as_stream(stream)>>>(x, y, n, C, (long long)H * W, ld);
  return check_launch("ff3d_nchw_to_nhwc");
}

extern "C" int ff3d_maxpool3x3s2(const float* x, float* y, int n, int H, int W, int C, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(C % 4 == 0, "maxpool3x3s2: C must be a multiple of 4");
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  long long total = (long long)n * Ho * Wo * (C / 4);
  if (total == 0) return FF3D_OK;
  maxpool3x3s2_kernel<<<cam_grid(total, 256), 256, 0, as_stream(stream)>>>(x, y, n, H, W, Ho, Wo, C);
  return check_launch("ff3d_maxpool3x3s2");
}

extern "C" int ff3d_upsample_add(float* dst, const float* src, int n, int Hd, int Wd, int Hs, int Ws, int C,
                                 ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(C % 4 == 0 && Hd > 0 && Wd > 0 && Hs > 0 && Ws > 0, "upsample_add: bad shape");
  long long total = (long long)n * Hd * Wd * (C / 4);
  if (total == 0) return FF3D_OK;
  upsample_add_kernel<<<cam_grid(total, 256), 256, 0, as_stream(stream)>>>(dst, src, n, Hd, Wd, Hs, Ws, C,
                                                                          (float)Hs / (float)Hd, (float)Ws / (float)Wd);
  return check_launch("ff3d_upsample_add");
}

extern "C" int ff3d_lss_splat(const float* dn, int ld, const float* frustum, const float* rots, const float* trans,
                              float* bev, int B, int cams, int D, int fH, int fW, const float* lo3, const float* dx3,
                              int nx, int ny, int nz, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(D > 0 && D <= 64 && ld >= kCamC + D && ld % 4 == 0, "lss_splat: need D <= 64 and ld >= 64 + D");
  FF3D_REQUIRE(B > 0 && cams > 0 && nx > 0 && ny > 0 && nz > 0, "lss_splat: bad shape");
  SplatParams p;
  p.dn = dn; p.ld = ld; p.frustum = frustum; p.rots = rots; p.trans = trans; p.bev = bev;
  p.n_img = B * cams; p.cams = cams; p.D = D; p.fH = fH; p.fW = fW;
  for (int i = 0; i < 3; ++i) { p.lo[i] = lo3[i]; p.dx[i] = dx3[i]; }
  p.nx = nx; p.ny = ny; p.nz = nz;
  cudaError_t e = cudaMemsetAsync(bev, 0, sizeof(float) * (size_t)B * ny * nx * nz * kCamC, as_stream(stream));
  if (e != cudaSuccess) { set_error("ff3d_lss_splat: memset: %s", cudaGetErrorString(e)); return FF3D_ECUDA; }
  long long warps = (long long)p.n_img * fH * fW;
  lss_splat_kernel<<<cdiv(warps * 32, 256), 256, 0, as_stream(stream)>>>(p);
  return check_launch("ff3d_lss_splat");
}
