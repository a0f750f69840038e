// Stable LSD radix sort of (uint32 key, int32 value) pairs whose count lives on the device.
//
// Used by the sparse-conv rulebook (spconv.cu): output rows are ordered by their tap-presence mask so that the 128-row
// tiles of the gather-GEMM are (nearly) homogeneous and whole taps can be skipped per tile.  9-bit digits, three
// kernels per pass:
//   rs_hist     per 4096-element block: digit histogram -> H[digit][block]
//   rs_rowscan  one warp per digit: exclusive scan of H[digit][*] (in place) + digit totals T[digit]
//   rs_scatter  per block: digit bases (scan of T, redone per block in shared memory) + H[digit][block] + an in-block
//               stable rank (warps own contiguous 512-element runs; __match_any_sync ranks equal digits by lane)
// Stable, deterministic, no global atomics.  HBM-bound integer work: 2 x 8 bytes per element and pass.
#include "common.cuh"

namespace ff3d {

constexpr int RS_BITS = 9;
constexpr int RS_BINS = 1 << RS_BITS;
constexpr int RS_WARPS = 8;
constexpr int RS_ROUNDS = 16;
constexpr int RS_WARP_ELEMS = 32 * RS_ROUNDS;            // 512 contiguous elements per warp
constexpr int RS_ELEMS = RS_WARPS * RS_WARP_ELEMS;       // 4096 per block

__global__ void __launch_bounds__(256) rs_hist_kernel(const uint32_t* __restrict__ keys, const int* __restrict__ n_dev,
                                                       int cap, int shift, int nblk, int* __restrict__ H) {
  __shared__ int h[RS_BINS];
  const int n = min(*n_dev, cap);
  for (int d = threadIdx.x; d < RS_BINS; d += blockDim.x) h[d] = 0;
  __syncthreads();
  const int base = blockIdx.x * RS_ELEMS;
  const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
  for (int k = threadIdx.x; k < RS_ELEMS; k += blockDim.x) {
    const int i = base + k;
    // warp-aggregated counting: tap masks are heavily skewed (a few digits dominate), and 32 lanes adding to the same
    // shared-memory word serialise; one add per distinct digit of the warp instead
    const int d = i < n ? (int)((keys[i] >> shift) & (RS_BINS - 1)) : RS_BINS + (int)(threadIdx.x & 31);
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if (i < n && (peers & lt) == 0u) atomicAdd(&h[d], __popc(peers));
  }
  __syncthreads();
  for (int d = threadIdx.x; d < RS_BINS; d += blockDim.x) H[(size_t)d * nblk + blockIdx.x] = h[d];
}

// one warp per digit row
__global__ void __launch_bounds__(256) rs_rowscan_kernel(int* __restrict__ H, int nblk, int* __restrict__ T) {
  const int d = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (d >= RS_BINS) return;
  int* row = H + (size_t)d * nblk;
  int run = 0;
  for (int b0 = 0; b0 < nblk; b0 += 32) {
    const int b = b0 + lane;
    const int v = b < nblk ? row[b] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (b < nblk) row[b] = run + inc - v;
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) T[d] = run;
}

__global__ void __launch_bounds__(256) rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const int* __restrict__ vals_in,
                                                          const int* __restrict__ n_dev, int cap, int shift, int nblk,
                                                          const int* __restrict__ H, const int* __restrict__ T,
                                                          uint32_t* __restrict__ keys_out, int* __restrict__ vals_out) {
  __shared__ int wh[RS_WARPS][RS_BINS];                  // per-warp digit counts, then running output cursors
  __shared__ int dbase[RS_BINS];
  const int n = min(*n_dev, cap);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int base = blockIdx.x * RS_ELEMS;
  if (base >= n) return;                                 // uniform for the block
  for (int i = tid; i < RS_WARPS * RS_BINS; i += blockDim.x) (&wh[0][0])[i] = 0;
  // exclusive scan of the 512 digit totals (each block redoes it: 2 values per thread + one warp pass over 8 partials)
  {
    const int a = T[2 * tid], b = T[2 * tid + 1];
    int inc = a + b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    __shared__ int wsum[8];
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    int off = 0;
    for (int k = 0; k < w; ++k) off += wsum[k];
    const int excl = off + inc - (a + b);
    dbase[2 * tid] = excl;
    dbase[2 * tid + 1] = excl + a;
  }
  __syncthreads();
  const int wbase = base + w * RS_WARP_ELEMS;
  uint32_t kreg[RS_ROUNDS];
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const int i = wbase + r * 32 + lane;
    kreg[r] = i < n ? keys_in[i] : 0u;
    const int d = i < n ? (int)((kreg[r] >> shift) & (RS_BINS - 1)) : RS_BINS + lane;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if (i < n && (peers & ((1u << lane) - 1u)) == 0u) wh[w][d] += __popc(peers);     // this warp's private row: no atomics
    __syncwarp();                                        // the next round's leader of digit d may be another lane
  }
  __syncthreads();
  for (int d = tid; d < RS_BINS; d += blockDim.x) {
    int run = dbase[d] + H[(size_t)d * nblk + blockIdx.x];
#pragma unroll
    for (int k = 0; k < RS_WARPS; ++k) {
      const int t = wh[k][d];
      wh[k][d] = run;
      run += t;
    }
  }
  __syncthreads();
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const int i = wbase + r * 32 + lane;
    const bool act = i < n;
    const int d = act ? (int)((kreg[r] >> shift) & (RS_BINS - 1)) : RS_BINS + lane;   // inactive lanes match nobody
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    int pos = 0;
    if (act) pos = wh[w][d] + __popc(peers & lt);
    __syncwarp();
    if (act && (peers & lt) == 0u) wh[w][d] += __popc(peers);
    __syncwarp();
    if (act) {
      keys_out[pos] = kreg[r];
      vals_out[pos] = vals_in ? vals_in[i] : i;
    }
  }
}

static inline int rs_nblk(int cap) { return cap > 0 ? (cap + RS_ELEMS - 1) / RS_ELEMS : 1; }

// layout of the sort workspace
struct SortWs {
  size_t keys_a, keys_b, vals_a, vals_b, H, T, total;
};
static SortWs sort_ws(int cap) {
  SortWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t c = (size_t)(cap > 0 ? cap : 1);
  w.keys_a = take(4 * c); w.keys_b = take(4 * c); w.vals_a = take(4 * c); w.vals_b = take(4 * c);
  w.H = take(sizeof(int) * (size_t)RS_BINS * rs_nblk(cap));
  w.T = take(sizeof(int) * RS_BINS);
  w.total = off;
  return w;
}

// sorts ascending by the low key_bits of the keys; vals_in == NULL means vals = 0..n-1 (the result is the permutation)
int sort_pairs(const uint32_t* keys_in, const int* vals_in, const int* n_dev, int cap, int key_bits, uint32_t* keys_out,
               int* vals_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (cap <= 0) return FF3D_OK;
  const SortWs w = sort_ws(cap);
  if (ws == nullptr || ws_bytes < w.total) {
    set_error("ff3d_sort_pairs: workspace too small (%zu < %zu)", ws_bytes, w.total);
    return FF3D_EWORKSPACE;
  }
  uint8_t* base = static_cast<uint8_t*>(ws);
  uint32_t* kbuf[2] = {reinterpret_cast<uint32_t*>(base + w.keys_a), reinterpret_cast<uint32_t*>(base + w.keys_b)};
  int* vbuf[2] = {reinterpret_cast<int*>(base + w.vals_a), reinterpret_cast<int*>(base + w.vals_b)};
  int* H = reinterpret_cast<int*>(base + w.H);
  int* T = reinterpret_cast<int*>(base + w.T);
  const int passes = key_bits <= 0 ? 1 : (key_bits + RS_BITS - 1) / RS_BITS;
  const int nblk = rs_nblk(cap);
  const uint32_t* kin = keys_in;
  const int* vin = vals_in;
  for (int p = 0; p < passes; ++p) {
    uint32_t* kout = (p == passes - 1) ? keys_out : kbuf[p & 1];
    int* vout = (p == passes - 1) ? vals_out : vbuf[p & 1];
    rs_hist_kernel<<<nblk, 256, 0, st>>>(kin, n_dev, cap, p * RS_BITS, nblk, H);
    rs_rowscan_kernel<<<RS_BINS / 8, 256, 0, st>>>(H, nblk, T);
    rs_scatter_kernel<<<nblk, 256, 0, st>>>(kin, vin, n_dev, cap, p * RS_BITS, nblk, H, T, kout, vout);
    kin = kout;
    vin = vout;
  }
  return check_launch("ff3d_sort_pairs");
}

size_t sort_workspace_bytes(int cap) { return sort_ws(cap).total; }

}  // namespace ff3d

extern "C" size_t ff3d_sort_workspace_bytes(int cap) { return ff3d::sort_workspace_bytes(cap); }

extern "C" int ff3d_sort_pairs(const uint32_t* keys_in, const int* vals_in, const int* n_dev, int cap, int key_bits,
                               uint32_t* keys_out, int* vals_out, void* workspace, size_t workspace_bytes,
                               ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(keys_in && n_dev && keys_out && vals_out, "ff3d_sort_pairs: null argument");
  FF3D_REQUIRE(key_bits >= 1 && key_bits <= 32, "ff3d_sort_pairs: key_bits=%d out of range", key_bits);
  return sort_pairs(keys_in, vals_in, n_dev, cap, key_bits, keys_out, vals_out, workspace, workspace_bytes,
                    as_stream(stream));
}
