// Output side of the head: per-task circle / rotated-BEV NMS of get_bboxes
// (projects/mmdet3d_plugin/models/dense_heads/focal_decoder.py:1333-1393) and the test-time-augmentation merge
// (core/post_processing/merge_augs.py:111-184: per-class rotated NMS + IoU-weighted box voting).
//
// Replaces [upstream] mmdet3d.core.circle_nms (numba, host), mmdet3d.ops.iou3d nms_gpu / boxes_iou_bev (iou3d_kernel.cu).
// At most ~1000 boxes per scene: one CTA per (scene, task) sorts its members by score in shared memory (bitonic), then
// walks them in order -- the walk is inherently sequential, the suppression test of each kept box against the rest is
// parallel over the CTA.  The rotated IoU is computed by clipping rectangle A against the four half-planes of rectangle B
// (Sutherland-Hodgman) and the shoelace formula, not by the reference's intersection-point / angular-sort construction.
#include "common.cuh"

namespace ff3d {

constexpr int NMS_MAX = 1024;      // boxes per (scene, task)
constexpr int NMS_THREADS = 256;
constexpr int NMS_MAX_TASKS = 8;
constexpr int NMS_MAX_CLASSES = 32;

struct NmsTasks {
  int n_tasks;
  unsigned int class_mask[NMS_MAX_TASKS];   // bit c = class c belongs to the task
  float radius[NMS_MAX_TASKS];              // <= 0: the task keeps every member
};

struct Pt { float x, y; };

// area of the intersection of two rotated rectangles given as (x1, y1, x2, y2, angle) -- the iou3d "xyxyr" BEV format
__device__ float rot_intersection(const float* a, const float* b) {
  // rectangle A as a polygon (counter-clockwise)
  Pt poly[8], tmp[8];
  {
    const float cx = 0.5f * (a[0] + a[2]), cy = 0.5f * (a[1] + a[3]);
    const float hx = 0.5f * (a[2] - a[0]), hy = 0.5f * (a[3] - a[1]);
    const float c = cosf(a[4]), s = sinf(a[4]);
    const float dx[4] = {-hx, hx, hx, -hx}, dy[4] = {-hy, -hy, hy, hy};
#pragma unroll
    for (int i = 0; i < 4; ++i) { poly[i].x = cx + dx[i] * c - dy[i] * s; poly[i].y = cy + dx[i] * s + dy[i] * c; }
  }
  int n = 4;
  // clip against the four edges of B: in B's frame the rectangle is |u| <= hx, |v| <= hy
  const float bcx = 0.5f * (b[0] + b[2]), bcy = 0.5f * (b[1] + b[3]);
  const float bhx = 0.5f * (b[2] - b[0]), bhy = 0.5f * (b[3] - b[1]);
  const float bc = cosf(b[4]), bs = sinf(b[4]);
  for (int i = 0; i < n; ++i) {                         // to B's frame
    const float px = poly[i].x - bcx, py = poly[i].y - bcy;
    poly[i].x = px * bc + py * bs;
    poly[i].y = -px * bs + py * bc;
  }
  for (int e = 0; e < 4 && n > 0; ++e) {
    // half-plane: sgn * coord <= lim   (e = 0: u <= hx, 1: -u <= hx, 2: v <= hy, 3: -v <= hy)
    const bool use_x = e < 2;
    const float sgn = (e & 1) ? -1.f : 1.f;
    const float lim = use_x ? bhx : bhy;
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const Pt p = poly[i], q = poly[(i + 1) % n];
      const float dp = sgn * (use_x ? p.x : p.y) - lim, dq = sgn * (use_x ? q.x : q.y) - lim;
      if (dp <= 0.f) tmp[m++] = p;
      if ((dp <= 0.f) != (dq <= 0.f)) {
        const float t = dp / (dp - dq);
        tmp[m].x = p.x + t * (q.x - p.x);
        tmp[m].y = p.y + t * (q.y - p.y);
        ++m;
      }
    }
    n = m;
    for (int i = 0; i < n; ++i) poly[i] = tmp[i];
  }
  if (n < 3) return 0.f;
  float area = 0.f;
  for (int i = 0; i < n; ++i) {
    const Pt p = poly[i], q = poly[(i + 1) % n];
    area += p.x * q.y - q.x * p.y;
  }
  return 0.5f * fabsf(area);
}

__device__ __forceinline__ float rot_iou(const float* a, const float* b) {
  const float sa = (a[2] - a[0]) * (a[3] - a[1]), sb = (b[2] - b[0]) * (b[3] - b[1]);
  const float inter = rot_intersection(a, b);
  return inter / fmaxf(sa + sb - inter, 1e-8f);        // iou3d's EPS
}

// (x, y, z, dx, dy, dz, yaw, ...) -> (x1, y1, x2, y2, yaw): LiDARInstance3DBoxes.bev + xywhr2xyxyr
__device__ __forceinline__ void box_to_xyxyr(const float* bx, float* o) {
  o[0] = bx[0] - 0.5f * bx[3]; o[1] = bx[1] - 0.5f * bx[4];
  o[2] = bx[0] + 0.5f * bx[3]; o[3] = bx[1] + 0.5f * bx[4];
  o[4] = bx[6];
}

// one CTA per (scene, task).  mode 0 = circle (squared centre distance <= radius suppresses), 1 = rotated IoU > radius
__global__ void __launch_bounds__(NMS_THREADS) nms_tasks_kernel(const float* __restrict__ boxes, int box_ld,
                                                                const float* __restrict__ scores, const int* __restrict__ labels,
                                                                const unsigned char* __restrict__ keep_in, int nq, NmsTasks tasks,
                                                                int mode, int pre_max, int post_max, unsigned char* keep_out) {
  __shared__ unsigned long long key[NMS_MAX];          // (orderable score bits << 32) | index: sorted descending
  __shared__ unsigned char sup[NMS_MAX];
  __shared__ int n_members, n_kept, stop;
  const int b = blockIdx.x, task = blockIdx.y, tid = threadIdx.x;
  const float* bx = boxes + (size_t)b * nq * box_ld;
  const float* sc = scores + (size_t)b * nq;
  const int* lb = labels + (size_t)b * nq;
  const unsigned char* kin = keep_in + (size_t)b * nq;
  unsigned char* kout = keep_out + (size_t)b * nq;
  const unsigned int cmask = tasks.class_mask[task];
  const float radius = tasks.radius[task];
  if (tid == 0) { n_members = 0; n_kept = 0; stop = 0; }
  for (int i = tid; i < NMS_MAX; i += NMS_THREADS) { key[i] = 0ull; sup[i] = 0; }
  __syncthreads();
  for (int i = tid; i < nq; i += NMS_THREADS) {
    const int l = lb[i];
    if (kin[i] && l >= 0 && l < NMS_MAX_CLASSES && ((cmask >> l) & 1u)) {
      if (radius <= 0.f) { kout[i] = 1; continue; }   // focal_decoder.py:1376: the task keeps all of its boxes
      const int slot = atomicAdd(&n_members, 1);
      unsigned int u = __float_as_uint(sc[i]);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone map float -> uint
      // ties: the higher index first (reversed stable ascending sort, as scores.argsort()[::-1])
      key[slot] = ((unsigned long long)u << 32) | (unsigned int)i;
    }
  }
  __syncthreads();
  const int n = n_members;
  if (radius <= 0.f || n == 0) return;
  // bitonic sort, descending, over the next power of two >= n (zero keys sort last)
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int k = 2; k <= np2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < np2; i += NMS_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = key[i], c = key[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (a < c) : (a > c)) { key[i] = c; key[ixj] = a; }
        }
      }
      __syncthreads();
    }
  const int n_use = (mode == 1 && pre_max > 0 && pre_max < n) ? pre_max : n;
  for (int i = 0; i < n_use; ++i) {
    if (stop) break;
    if (sup[i]) continue;                               // uniform: shared flag read after the barrier below
    const int bi = (int)(key[i] & 0xFFFFFFFFu);
    if (tid == 0) {
      kout[bi] = 1;
      if (++n_kept >= post_max && post_max > 0) stop = 1;
    }
    float ai[5];
    if (mode == 1) box_to_xyxyr(bx + (size_t)bi * box_ld, ai);
    const float xi = bx[(size_t)bi * box_ld], yi = bx[(size_t)bi * box_ld + 1];
    for (int j = i + 1 + tid; j < n_use; j += NMS_THREADS) {
      if (sup[j]) continue;
      const int bj = (int)(key[j] & 0xFFFFFFFFu);
      bool hit;
      if (mode == 0) {
        const float dx = xi - bx[(size_t)bj * box_ld], dy = yi - bx[(size_t)bj * box_ld + 1];
        hit = dx * dx + dy * dy <= radius;
      } else {
        float aj[5];
        box_to_xyxyr(bx + (size_t)bj * box_ld, aj);
        hit = rot_iou(ai, aj) > radius;
      }
      if (hit) sup[j] = 1;
    }
    __syncthreads();
  }
}

__global__ void boxes_iou_bev_kernel(const float* __restrict__ a, int lda, int n, const float* __restrict__ b, int ldb, int m,
                                     float* __restrict__ iou) {
  const long long total = (long long)n * m;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / m), j = (int)(e - (long long)i * m);
    float ai[5], bj[5];
    box_to_xyxyr(a + (size_t)i * lda, ai);
    box_to_xyxyr(b + (size_t)j * ldb, bj);
    iou[e] = rot_iou(ai, bj);
  }
}

// merge_augs.py:152-165: voted[i] = sum_j w_ij box_j / (sum_j w_ij + 1e-6), w = iou zeroed below vote_thresh; yaw from the
// weighted sine / cosine sums.  One warp per selected box.
__global__ void box_voting_kernel(const float* __restrict__ iou, int n_sel, int m, const float* __restrict__ boxes, int ld,
                                  int dim, float vote_thresh, float* __restrict__ out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_sel) return;
  float acc[12];
#pragma unroll
  for (int d = 0; d < 12; ++d) acc[d] = 0.f;
  float wsum = 0.f, ssum = 0.f, csum = 0.f;
  for (int j = lane; j < m; j += 32) {
    float v = iou[(size_t)w * m + j];
    if (v < vote_thresh) v = 0.f;
    wsum += v;
    const float* bx = boxes + (size_t)j * ld;
    for (int d = 0; d < dim; ++d) acc[d] += v * bx[d];
    ssum += v * sinf(bx[6]);
    csum += v * cosf(bx[6]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    csum += __shfl_xor_sync(0xffffffffu, csum, o);
    for (int d = 0; d < dim; ++d) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o);
  }
  if (lane == 0) {
    const float den = wsum + 1e-6f;
    for (int d = 0; d < dim; ++d) out[(size_t)w * dim + d] = acc[d] / den;
    out[(size_t)w * dim + 6] = atan2f(ssum / den, csum / den);
  }
}

// bbox3d_mapping_back ([upstream] mmdet3d) for LiDAR boxes: undo the flips, then scale by 1 / scale_factor
__global__ void boxes_map_back_kernel(float* boxes, int ld, int dim, int n, float inv_scale, int flip_h, int flip_v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* b = boxes + (size_t)i * ld;
  const float pi = 3.14159265358979323846f;
  if (flip_h) { b[1] = -b[1]; b[6] = -b[6] + pi; if (dim > 8) b[8] = -b[8]; }
  if (flip_v) { b[0] = -b[0]; b[6] = -b[6]; if (dim > 7) b[7] = -b[7]; }
  for (int d = 0; d < 6; ++d) b[d] *= inv_scale;
  for (int d = 7; d < dim; ++d) b[d] *= inv_scale;
}

}  // namespace ff3d

extern "C" int ff3d_nms_tasks(const float* boxes, int box_ld, const float* scores, const int* labels,
                              const unsigned char* keep_in, int B, int nq, int n_tasks, const unsigned int* class_masks_host,
                              const float* radius_host, int mode, int pre_max, int post_max, unsigned char* keep_out,
                              ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(nq >= 1 && nq <= NMS_MAX, "nms_tasks: nq=%d out of range (<= %d)", nq, NMS_MAX);
  FF3D_REQUIRE(n_tasks >= 1 && n_tasks <= NMS_MAX_TASKS, "nms_tasks: %d tasks (<= %d)", n_tasks, NMS_MAX_TASKS);
  FF3D_REQUIRE(mode == 0 || mode == 1, "nms_tasks: mode 0 (circle) or 1 (rotate)");
  FF3D_REQUIRE(box_ld >= 7, "nms_tasks: boxes are (x, y, z, dx, dy, dz, yaw, ...) rows");
  if (B <= 0) return FF3D_OK;
  NmsTasks t;
  t.n_tasks = n_tasks;
  for (int i = 0; i < n_tasks; ++i) { t.class_mask[i] = class_masks_host[i]; t.radius[i] = radius_host[i]; }
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(keep_out, 0, (size_t)B * nq, st);
  nms_tasks_kernel<<<dim3(B, n_tasks), NMS_THREADS, 0, st>>>(boxes, box_ld, scores, labels, keep_in, nq, t, mode, pre_max, post_max,
                                                            keep_out);
  return check_launch("ff3d_nms_tasks");
}

extern "C" int ff3d_boxes_iou_bev(const float* a, int lda, int n, const float* b, int ldb, int m, float* iou, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(lda >= 7 && ldb >= 7, "boxes_iou_bev: boxes are (x, y, z, dx, dy, dz, yaw, ...) rows");
  if (n <= 0 || m <= 0) return FF3D_OK;
  long long nb = ((long long)n * m + 127) / 128, cap = (long long)num_sms() * 16;
  boxes_iou_bev_kernel<<<(int)(nb > cap ? cap : nb), 128, 0, as_stream(stream)>>>(a, lda, n, b, ldb, m, iou);
  return check_launch("ff3d_boxes_iou_bev");
}

extern "C" int ff3d_box_voting(const float* iou, int n_sel, int m, const float* boxes, int ld, int dim, float vote_thresh,
                               float* out, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(dim >= 7 && dim <= 12 && ld >= dim, "box_voting: box dim %d unsupported (7..12)", dim);
  if (n_sel <= 0) return FF3D_OK;
  box_voting_kernel<<<cdiv((long long)n_sel * 32, 128), 128, 0, as_stream(stream)>>>(iou, n_sel, m, boxes, ld, dim, vote_thresh, out);
  return check_launch("ff3d_box_voting");
}

extern "C" int ff3d_boxes_map_back(float* boxes, int ld, int dim, int n, float scale_factor, int flip_horizontal,
                                   int flip_vertical, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(dim >= 7 && ld >= dim && scale_factor != 0.f, "boxes_map_back: bad arguments");
  if (n <= 0) return FF3D_OK;
  boxes_map_back_kernel<<<cdiv(n, 128), 128, 0, as_stream(stream)>>>(boxes, ld, dim, n, 1.f / scale_factor, flip_horizontal,
                                                                   flip_vertical);
  return check_launch("ff3d_boxes_map_back");
}
