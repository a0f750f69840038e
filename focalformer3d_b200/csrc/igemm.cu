// Implicit-GEMM (fp32 SIMT path): y[row(m), n] = act(sum_{t,c} x[src(m,t), c] * w[t,c,n] + bias[n] + res[m,n]).
// One kernel family for linear layers, NHWC conv2d windows and rulebook sparse conv (see include/ff3d.h).
// Register-prefetch double buffering: the global loads of iteration i+1 are in flight while iteration i is
// multiplied out of shared memory.  Output-stationary: every output row is written exactly once, no atomics.
#include "common.cuh"

namespace ff3d {

struct GemmP {
  int mode, M;
  const int* m_dev;
  int cin, cout, taps;
  const float* x; int ldx;
  const float* x2;
  const float* w; int ldw;
  const float* bias;
  const float* res; int ldres;
  float* y; int ldy;
  int act, res_after_act;
  int B, H, W, Ho, Wo, kh, kw, stride, pad;
  long long x_bstride, y_bstride, y_row0;
  int ux, uy, dx, dy;
  const int* nbr; int nbr_stride;
  const int* y_off;
};

constexpr int BM = 128, BN = 64, BK = 16, NT = 256, APAD = 4;

template <int MODE>
__global__ void __launch_bounds__(NT, 2) igemm_kernel(const GemmP p) {
  __shared__ __align__(16) float As[BK][BM + APAD];
  __shared__ __align__(16) float Bs[BK][BN];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  int Mv = p.M;
  if (p.m_dev) { int md = *p.m_dev; Mv = md < Mv ? md : Mv; }
  if (m0 >= Mv) return;

  // ---- A-load assignment: 2 float4 per thread: row = (tid + j*256) >> 2, kq = tid & 3
  const int kq = tid & 3;
  int arow[2];
  arow[0] = tid >> 2;
  arow[1] = (tid + NT) >> 2;
  int cb[2], coy[2], cox[2];  // conv2d decomposition of the two rows
  bool rvalid[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    int m = m0 + arow[j];
    rvalid[j] = m < Mv;
    cb[j] = coy[j] = cox[j] = 0;
    if (MODE == FF3D_GEMM_CONV2D && rvalid[j]) {
      int hw = p.Ho * p.Wo;
      cb[j] = m / hw;
      int r = m - cb[j] * hw;
      coy[j] = r / p.Wo;
      cox[j] = r - coy[j] * p.Wo;
    }
  }
  // ---- B-load assignment: 1 float4 per thread
  const int kb = tid >> 4;
  const int nb = (tid & 15) * 4;

  const int nchunks = (p.cin + BK - 1) / BK;
  const int iters = p.taps * nchunks;
  int cur_t = -1;
  long long src[2] = {-1, -1};
  float4 ra[2], rb;

  auto fetch = [&](int it) {
    int t = it / nchunks;
    int c0 = (it - t * nchunks) * BK;
    if (t != cur_t) {
      cur_t = t;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        long long s = -1;
        if (rvalid[j]) {
          int m = m0 + arow[j];
          if (MODE == FF3D_GEMM_ROWS) {
            s = m;
          } else if (MODE == FF3D_GEMM_CONV2D) {
            int ky = t / p.kw, kx = t - ky * p.kw;
            int iy = coy[j] * p.stride - p.pad + ky;
            int ix = cox[j] * p.stride - p.pad + kx;
            if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) s = cb[j] * p.x_bstride + (long long)iy * p.W + ix;
          } else {
            s = __ldg(p.nbr + (size_t)t * p.nbr_stride + m);
          }
        }
        src[j] = s;
      }
    }
    int k = c0 + kq * 4;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src[j] >= 0 && k < p.cin) {
        v = __ldg(reinterpret_cast<const float4*>(p.x + src[j] * p.ldx + k));
        if (MODE == FF3D_GEMM_ROWS && p.x2) {
          float4 u = __ldg(reinterpret_cast<const float4*>(p.x2 + src[j] * p.ldx + k));
          v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
      }
      ra[j] = v;
    }
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 + kb < p.cin && n0 + nb < p.ldw)
      rb = __ldg(reinterpret_cast<const float4*>(p.w + ((size_t)t * p.cin + c0 + kb) * p.ldw + n0 + nb));
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  fetch(0);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      As[kq * 4 + 0][arow[j]] = ra[j].x;
      As[kq * 4 + 1][arow[j]] = ra[j].y;
      As[kq * 4 + 2][arow[j]] = ra[j].z;
      As[kq * 4 + 3][arow[j]] = ra[j].w;
    }
    *reinterpret_cast<float4*>(&Bs[kb][nb]) = rb;
    __syncthreads();
    if (it + 1 < iters) fetch(it + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
  const int n = n0 + tx * 4;
  if (n >= p.cout) return;
  float bz[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < p.cout) bz[j] = __ldg(p.bias + n + j);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + ty * 8 + i;
    if (m >= Mv) continue;
    float* yp;
    if (MODE == FF3D_GEMM_CONV2D) {
      int hw = p.Ho * p.Wo;
      int b = m / hw;
      int r = m - b * hw;
      int oy = r / p.Wo, ox = r - (r / p.Wo) * p.Wo;
      long long row = b * p.y_bstride + p.y_row0 + (long long)(oy * p.uy + p.dy) * (p.Wo * p.ux) + ox * p.ux + p.dx;
      yp = p.y + row * p.ldy;
    } else if (MODE == FF3D_GEMM_SPARSE && p.y_off) {
      yp = p.y + __ldg(p.y_off + m);
    } else {
      yp = p.y + (long long)m * p.ldy;
    }
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bz[j];
    if (p.res_after_act) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], p.act);
    }
    if (p.res) {
      const float* rp = p.res + (long long)m * p.ldres + n;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < p.cout) v[j] += __ldg(rp + j);
    }
    if (!p.res_after_act) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], p.act);
    }
    if (n + 3 < p.cout && ((reinterpret_cast<uintptr_t>(yp + n) & 15) == 0)) {
      *reinterpret_cast<float4*>(yp + n) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < p.cout) yp[n + j] = v[j];
    }
  }
}

}  // namespace ff3d

extern "C" int ff3d_igemm(const ff3d_gemm_desc* d, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(d != nullptr, "ff3d_igemm: null descriptor");
  FF3D_REQUIRE(d->mode >= 0 && d->mode <= 2, "ff3d_igemm: bad mode %d", d->mode);
  FF3D_REQUIRE(d->cin > 0 && d->cin % 4 == 0, "ff3d_igemm: cin=%d must be a positive multiple of 4", d->cin);
  FF3D_REQUIRE(d->ldx % 4 == 0 && d->ldw % 4 == 0 && d->ldw >= d->cout, "ff3d_igemm: ldx=%d ldw=%d cout=%d", d->ldx,
               d->ldw, d->cout);
  FF3D_REQUIRE(d->x && d->w && d->y && d->cout > 0 && d->taps > 0, "ff3d_igemm: null operand or empty shape");
  FF3D_REQUIRE((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0,
               "ff3d_igemm: x and w must be 16-byte aligned");
  if (d->M <= 0) return FF3D_OK;
  GemmP p;
  p.mode = d->mode; p.M = d->M; p.m_dev = d->m_dev;
  p.cin = d->cin; p.cout = d->cout; p.taps = d->taps;
  p.x = d->x; p.ldx = d->ldx; p.x2 = d->x2; p.w = d->w; p.ldw = d->ldw; p.bias = d->bias;
  p.res = d->res; p.ldres = d->ldres; p.y = d->y; p.ldy = d->ldy; p.act = d->act; p.res_after_act = d->res_after_act;
  p.B = d->B; p.H = d->H; p.W = d->W; p.Ho = d->Ho; p.Wo = d->Wo; p.kh = d->kh; p.kw = d->kw;
  p.stride = d->stride; p.pad = d->pad;
  p.ux = d->ux > 0 ? d->ux : 1; p.uy = d->uy > 0 ? d->uy : 1; p.dx = d->dx; p.dy = d->dy;
  p.x_bstride = d->x_bstride; p.y_bstride = d->y_bstride; p.y_row0 = d->y_row0;
  p.nbr = d->nbr; p.nbr_stride = d->nbr_stride; p.y_off = d->y_off;
  if (d->mode == FF3D_GEMM_CONV2D) {
    FF3D_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Ho > 0 && d->Wo > 0 && d->kh > 0 && d->kw > 0 && d->stride > 0,
                 "ff3d_igemm: bad conv geometry");
    FF3D_REQUIRE(d->taps == d->kh * d->kw, "ff3d_igemm: taps != kh*kw");
    FF3D_REQUIRE((long long)d->B * d->Ho * d->Wo == d->M, "ff3d_igemm: M != B*Ho*Wo");
    if (p.x_bstride == 0) p.x_bstride = (long long)d->H * d->W;
    if (p.y_bstride == 0) p.y_bstride = (long long)d->Ho * p.uy * d->Wo * p.ux;
  } else if (d->mode == FF3D_GEMM_SPARSE) {
    FF3D_REQUIRE(d->nbr != nullptr && d->nbr_stride >= d->M, "ff3d_igemm: sparse mode needs nbr [taps, >=M]");
  } else {
    FF3D_REQUIRE(d->taps == 1, "ff3d_igemm: ROWS mode has a single tap");
  }
  dim3 grid(cdiv(d->M, BM), cdiv(d->cout, BN));
  cudaStream_t st = as_stream(stream);
  if (d->mode == FF3D_GEMM_ROWS) igemm_kernel<FF3D_GEMM_ROWS><<<grid, NT, 0, st>>>(p);
  else if (d->mode == FF3D_GEMM_CONV2D) igemm_kernel<FF3D_GEMM_CONV2D><<<grid, NT, 0, st>>>(p);
  else igemm_kernel<FF3D_GEMM_SPARSE><<<grid, NT, 0, st>>>(p);
  return check_launch("ff3d_igemm");
}
