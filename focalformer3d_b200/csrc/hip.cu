// One Hard-Instance-Probing stage (projects/mmdet3d_plugin/models/dense_heads/focal_decoder.py:631-782):
//   heat = sigmoid(logits) * acc_mask -> class-aware 3x3 local-max NMS (border ring never a maximum; exempt classes
//   use a 1x1 window) -> top-k over all (class, cell) -> query feature/pos/score/label gathers ->
//   acc_mask *= 1 - dilate3x3(selected)   (exempt classes: 1x1).
// The reference runs ~25 PyTorch kernels per stage incl. torch.topk over C*H*W; here: 2 grid-wide element passes
// + one CTA per scene doing an exact 64-bit radix select (value bits | inverted index => unique keys, canonical
// tie-break towards the lower flat index) and all gathers.  HBM traffic ~= logits + mask + heat once each.
#include "common.cuh"

namespace ff3d {

struct HipP {
  int B, C, H, W, k, nms_kernel, ex_lo, ex_hi;
};

__global__ void hip_heat_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ logits2, int ldl2,
                                const float* __restrict__ acc_mask, float* __restrict__ heat, HipP p) {
  long long total = (long long)p.B * p.C * p.H * p.W;
  int HW = p.H * p.W;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int pos = (int)(e % HW);
    long long r = e / HW;
    int c = (int)(r % p.C);
    int b = (int)(r / p.C);
    float l = __ldg(logits + ((long long)b * HW + pos) * ldl + c);
    float s = 1.f / (1.f + expf(-l));
    if (logits2) {   // single-stage head: (sigmoid(heatmap_head) + sigmoid(heatmap_head_img)) / 2  (focal_decoder.py:549)
      float l2 = __ldg(logits2 + ((long long)b * HW + pos) * ldl2 + c);
      s = (s + 1.f / (1.f + expf(-l2))) / 2.f;
    }
    heat[e] = s * acc_mask[e];
  }
}

__global__ void hip_nms_kernel(const float* __restrict__ heat, float* __restrict__ nms_heat,
                               unsigned long long* __restrict__ cand, int* __restrict__ cand_cnt, HipP p) {
  long long total = (long long)p.B * p.C * p.H * p.W;
  int HW = p.H * p.W;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int pos = (int)(e % HW);
    long long r = e / HW;
    int c = (int)(r % p.C);
    int b = (int)(r / p.C);
    int y = pos / p.W, x = pos - y * p.W;
    float h = heat[e];
    float v;
    bool exempt = (c >= p.ex_lo && c <= p.ex_hi) || p.nms_kernel == 1;
    if (exempt) {
      v = h;
    } else if (y == 0 || x == 0 || y == p.H - 1 || x == p.W - 1) {
      v = 0.f;  // local_max stays 0 on the border ring (focal_decoder.py:672-676)
    } else {
      const float* hp = heat + (e - pos);
      float m = h;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) m = fmaxf(m, hp[(y + dy) * p.W + x + dx]);
      v = (h == m) ? h : 0.f;
    }
    nms_heat[e] = v;
    // warp-aggregated append (one atomic per warp and scene instead of one per candidate); a warp's elements may
    // straddle two scenes, so aggregate per scene of the lane
    const bool cand_ok = v > 0.f;
    const unsigned active = __activemask();
    const unsigned same_b = __match_any_sync(active, b);
    const unsigned voters = __ballot_sync(active, cand_ok) & same_b;
    if (cand_ok) {
      const int lane = threadIdx.x & 31;
      const int leader = __ffs(voters) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&cand_cnt[b], __popc(voters));
      base = __shfl_sync(voters, base, leader);
      const int slot = base + __popc(voters & ((1u << lane) - 1u));
      unsigned int flat = (unsigned int)(c * HW + pos);
      cand[(size_t)b * p.C * HW + slot] =
          ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - flat);
    }
  }
}

constexpr int SEL_THREADS = 1024;
constexpr int SEL_MAXK = 1024;
constexpr int SEL_CAP = 16384;   // contenders staged in dynamic shared memory (128 KB)

__global__ void __launch_bounds__(SEL_THREADS) hip_select_kernel(
    const unsigned long long* __restrict__ cand, const int* __restrict__ cand_cnt, const float* __restrict__ nms_heat,
    float* __restrict__ acc_mask, const float* __restrict__ feat, int ldf, int Cf, const float* __restrict__ cls_w,
    const float* __restrict__ cls_b, HipP p, int q0, int nq_total, int* __restrict__ top_idx,
    float* __restrict__ query_feat, float* __restrict__ query_pos, float* __restrict__ query_score,
    int* __restrict__ query_label) {
  __shared__ unsigned long long sel[SEL_MAXK];
  __shared__ int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_remaining, s_count, s_scan[SEL_THREADS / 32], s_base;
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int HW = p.H * p.W;
  const int CHW = p.C * HW;
  const unsigned long long* keys = cand + (size_t)b * CHW;
  const int n = cand_cnt[b];
  const int k = p.k;

  for (int i = tid; i < SEL_MAXK; i += SEL_THREADS) sel[i] = 0ull;
  if (tid == 0) { s_prefix = 0ull; s_remaining = k; s_count = 0; }
  __syncthreads();

  if (n > k) {
    // exact k-th largest key by MSB-first radix select, 8 bits per pass.  Pass 0 reads the global candidate list;
    // the contenders (keys sharing the winning top byte, typically a few thousand) are then staged in shared memory
    // so passes 1..7 never touch global memory again; keys above the winning byte are selected on the spot.
    extern __shared__ unsigned long long cont[];        // [SEL_CAP]
    __shared__ int s_ncont;
    for (int i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
    if (tid == 0) s_ncont = 0;
    __syncthreads();
    for (int i = tid; i < n; i += SEL_THREADS) atomicAdd(&hist[(int)(keys[i] >> 56)], 1);
    __syncthreads();
    if (tid == 0) {
      int rem = s_remaining, d = 255;
      for (; d > 0; --d) {
        if (hist[d] >= rem) break;
        rem -= hist[d];
      }
      s_remaining = rem;
      s_prefix = (unsigned long long)d << 56;
    }
    __syncthreads();
    const int d0 = (int)(s_prefix >> 56);
    const int ncont_total = hist[d0];
    const bool staged = ncont_total <= SEL_CAP;
    __syncthreads();
    if (staged) {
      for (int i = tid; i < n; i += SEL_THREADS) {
        unsigned long long key = keys[i];
        int d = (int)(key >> 56);
        if (d > d0) {
          int sidx = atomicAdd(&s_count, 1);
          if (sidx < SEL_MAXK) sel[sidx] = key;
        } else if (d == d0) {
          cont[atomicAdd(&s_ncont, 1)] = key;
        }
      }
      __syncthreads();
    }
    const unsigned long long* src = staged ? cont : keys;
    const int nsrc = staged ? ncont_total : n;
    for (int pass = 1; pass < 8; ++pass) {
      int shift = 56 - 8 * pass;
      for (int i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
      __syncthreads();
      unsigned long long prefix = s_prefix;
      unsigned long long himask = ~0ull << (shift + 8);
      for (int i = tid; i < nsrc; i += SEL_THREADS) {
        unsigned long long key = src[i];
        if ((key & himask) == prefix) atomicAdd(&hist[(int)((key >> shift) & 0xFF)], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int rem = s_remaining;
        int d = 255;
        for (; d > 0; --d) {
          if (hist[d] >= rem) break;
          rem -= hist[d];
        }
        s_remaining = rem;
        s_prefix = prefix | ((unsigned long long)d << shift);
      }
      __syncthreads();
    }
    unsigned long long thr = s_prefix;  // the k-th largest key (keys are unique)
    if (staged) {
      for (int i = tid; i < nsrc; i += SEL_THREADS) {
        unsigned long long key = src[i];
        if (key >= thr) {
          int sidx = atomicAdd(&s_count, 1);
          if (sidx < SEL_MAXK) sel[sidx] = key;
        }
      }
    } else {
      for (int i = tid; i < n; i += SEL_THREADS) {
        unsigned long long key = keys[i];
        if (key >= thr) {
          int sidx = atomicAdd(&s_count, 1);
          if (sidx < SEL_MAXK) sel[sidx] = key;
        }
      }
    }
    __syncthreads();
  } else {
    // degenerate: fewer positive candidates than k -> take them all, then fill with the lowest-index zero cells
    for (int i = tid; i < n; i += SEL_THREADS) sel[i] = keys[i];
    if (tid == 0) { s_count = n; s_base = n; }
    __syncthreads();
    const float* nh = nms_heat + (size_t)b * CHW;
    for (int base = 0; base < CHW && s_base < k; base += SEL_THREADS) {
      int idx = base + tid;
      bool z = idx < CHW && !(nh[idx] > 0.f);
      unsigned int bal = __ballot_sync(0xffffffffu, z);
      int lane = tid & 31, wid = tid >> 5;
      if (lane == 0) s_scan[wid] = __popc(bal);
      __syncthreads();
      int off = 0;
      for (int w = 0; w < wid; ++w) off += s_scan[w];
      int rank = s_base + off + __popc(bal & ((1u << lane) - 1u));
      if (z && rank < k)
        sel[rank] = (0ull << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)idx);
      __syncthreads();
      if (tid == 0) {
        int t = 0;
        for (int w = 0; w < SEL_THREADS / 32; ++w) t += s_scan[w];
        s_base += t;
      }
      __syncthreads();
    }
  }

  // bitonic sort (descending) of sel[0 .. SEL_MAXK)
  for (int size = 2; size <= SEL_MAXK; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      int i = tid;
      int j = i ^ stride;
      if (j > i) {
        unsigned long long a = sel[i], c = sel[j];
        bool desc = (i & size) == 0;
        if (desc ? (a < c) : (a > c)) { sel[i] = c; sel[j] = a; }
      }
    }
  }
  __syncthreads();

  // outputs
  for (int q = tid; q < k; q += SEL_THREADS) {
    unsigned long long key = sel[q];
    unsigned int flat = 0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull);
    int cls = (int)(flat / HW);
    int pos = (int)(flat - (unsigned int)cls * HW);
    int y = pos / p.W, x = pos - y * p.W;
    size_t qi = (size_t)b * nq_total + q0 + q;
    top_idx[(size_t)b * k + q] = (int)flat;
    query_label[qi] = cls;
    query_pos[qi * 2 + 0] = (float)x + 0.5f;
    query_pos[qi * 2 + 1] = (float)y + 0.5f;
    for (int c = 0; c < p.C; ++c) query_score[qi * p.C + c] = nms_heat[((size_t)b * p.C + c) * HW + pos];
    // accumulated positive mask: zero the dilated neighbourhood of the selected (class, cell)
    float* am = acc_mask + ((size_t)b * p.C + cls) * HW;
    bool exempt = (cls >= p.ex_lo && cls <= p.ex_hi) || p.nms_kernel == 1;
    int rad = exempt ? 0 : 1;
    for (int dy = -rad; dy <= rad; ++dy)
      for (int dx = -rad; dx <= rad; ++dx) {
        int yy = y + dy, xx = x + dx;
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) am[yy * p.W + xx] = 0.f;
      }
  }
  // query feature = stage feature at the cell + class encoding (Conv1d(C -> Cf, 1) of the one-hot)
  for (int e = tid; e < k * Cf; e += SEL_THREADS) {
    int q = e / Cf, c = e - q * Cf;
    unsigned long long key = sel[q];
    unsigned int flat = 0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull);
    int cls = (int)(flat / HW);
    int pos = (int)(flat - (unsigned int)cls * HW);
    float v = feat[((size_t)b * HW + pos) * ldf + c] + cls_w[cls * Cf + c] + cls_b[c];
    query_feat[((size_t)b * nq_total + q0 + q) * Cf + c] = v;
  }
}

struct HipWs { size_t heat, cand, cnt, total; };
static HipWs hip_layout(int B, int C, int H, int W) {
  HipWs w;
  size_t n = (size_t)B * C * H * W;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  w.heat = take(n * sizeof(float));
  w.cand = take(n * sizeof(unsigned long long));
  w.cnt = take(sizeof(int) * (size_t)B);
  w.total = o;
  return w;
}

}  // namespace ff3d

extern "C" size_t ff3d_hip_workspace_bytes(int B, int C, int H, int W) { return ff3d::hip_layout(B, C, H, W).total; }

extern "C" int ff3d_hip_stage(const float* logits, int ldl, const float* logits2, int ldl2, float* acc_mask, float* nms_heat, const float* feat, int ldf,
                              int Cf, const float* cls_w, const float* cls_b, int B, int C, int H, int W, int k,
                              int nms_kernel, int exempt_lo, int exempt_hi, int q0, int nq_total, int* top_idx,
                              float* query_feat, float* query_pos, float* query_score, int* query_label,
                              void* workspace, size_t workspace_bytes, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(k >= 1 && k <= SEL_MAXK, "hip_stage: k=%d unsupported (1..%d)", k, SEL_MAXK);
  FF3D_REQUIRE(nms_kernel == 1 || nms_kernel == 3, "hip_stage: nms_kernel=%d unsupported", nms_kernel);
  FF3D_REQUIRE((long long)C * H * W < 0x7FFFFFFFLL && k <= C * H * W, "hip_stage: bad sizes");
  HipWs w = hip_layout(B, C, H, W);
  if (workspace_bytes < w.total) {
    set_error("hip_stage: workspace %zu < required %zu", workspace_bytes, w.total);
    return FF3D_EWORKSPACE;
  }
  HipP p{B, C, H, W, k, nms_kernel, exempt_lo, exempt_hi};
  char* base = static_cast<char*>(workspace);
  float* heat = reinterpret_cast<float*>(base + w.heat);
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(base + w.cand);
  int* cnt = reinterpret_cast<int*>(base + w.cnt);
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)B, st);
  long long total = (long long)B * C * H * W;
  int nb = (int)((total + 255) / 256);
  int cap = num_sms() * 16;
  if (nb > cap) nb = cap;
  hip_heat_kernel<<<nb, 256, 0, st>>>(logits, ldl, logits2, ldl2, acc_mask, heat, p);
  hip_nms_kernel<<<nb, 256, 0, st>>>(heat, nms_heat, cand, cnt, p);
  // one-time opt-in to > 48 KB dynamic shared memory: function-local static initialiser (thread-safe, C++11)
  static const cudaError_t attr = cudaFuncSetAttribute(hip_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SEL_CAP * 8);
  (void)attr;
  hip_select_kernel<<<B, SEL_THREADS, SEL_CAP * sizeof(unsigned long long), st>>>(cand, cnt, nms_heat, acc_mask, feat, ldf, Cf, cls_w, cls_b, p, q0,
                                               nq_total, top_idx, query_feat, query_pos, query_score, query_label);
  return check_launch("ff3d_hip_stage");
}
