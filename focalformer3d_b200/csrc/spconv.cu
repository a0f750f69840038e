// Sparse-conv rulebooks as output-stationary neighbour maps (see include/ff3d.h).
//
// The reference's [upstream] spconv-v1 builds, for every conv, per-offset (in,out) pair lists with a cuckoo hash
// and then runs 27 x (gather, SGEMM, scatter-add).  Here a level keeps ONE open-addressing hash; a neighbour map
// nbr[t][o] names the input row feeding output row o through tap t, so the conv itself is a single gather-GEMM with
// no atomics (igemm.cu, FF3D_GEMM_SPARSE).  SubM maps are shared by all SubM convs of a level (the reference
// rebuilds them per conv: SparseBasicBlock passes no indice_key).
#include "common.cuh"

namespace ff3d {

__device__ __forceinline__ uint32_t lin_key(int b, int z, int y, int x, int D, int H, int W) {
  // unsigned arithmetic (callers guarantee batch * D * H * W < 2^32, which can exceed a signed int)
  return (((uint32_t)b * (uint32_t)D + (uint32_t)z) * (uint32_t)H + (uint32_t)y) * (uint32_t)W + (uint32_t)x;
}

__global__ void sp_hash_build_kernel(const int* __restrict__ coors, const int* __restrict__ n_dev, int cap, int D,
                                     int H, int W, uint32_t* hkeys, int* hvals, int hmask) {
  int n = min(*n_dev, cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = reinterpret_cast<const int4*>(coors)[i];
    bool ins;
    int s = hash_insert(hkeys, hmask, lin_key(c.x, c.y, c.z, c.w, D, H, W), &ins);
    if (s >= 0) hvals[s] = i;
  }
}

__global__ void sp_subm_map_kernel(const int* __restrict__ coors, const int* __restrict__ n_dev, int cap, int D, int H,
                                   int W, const uint32_t* __restrict__ hkeys, const int* __restrict__ hvals, int hmask,
                                   int* nbr) {
  int n = min(*n_dev, cap);
  long long total = (long long)n * 27;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int t = (int)(e / n);
    int o = (int)(e - (long long)t * n);
    int4 c = reinterpret_cast<const int4*>(coors)[o];
    int kz = t / 9, ky = (t / 3) % 3, kx = t % 3;
    int z = c.y + kz - 1, y = c.z + ky - 1, x = c.w + kx - 1;
    int r = -1;
    if (t == 13) {
      r = o;
    } else if (z >= 0 && z < D && y >= 0 && y < H && x >= 0 && x < W) {
      int s = hash_find(hkeys, hmask, lin_key(c.x, z, y, x, D, H, W));
      if (s >= 0) r = hvals[s];
    }
    nbr[(size_t)t * cap + o] = r;
  }
}

struct Down {
  int k[3], s[3], p[3];
  int D, H, W, Do, Ho, Wo;
};

// every (input site, tap) proposes an output site; first proposer allocates the row
__global__ void sp_down_sites_kernel(const int* __restrict__ coors_in, const int* __restrict__ n_in_dev, int cap_in,
                                     Down g, int* coors_out, int* n_out_dev, int cap_out, uint32_t* hkeys_out,
                                     int* hvals_out, int hmask_out, volatile int* overflow) {
  int n = min(*n_in_dev, cap_in);
  int kvol = g.k[0] * g.k[1] * g.k[2];
  long long total = (long long)n * kvol;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int i = (int)(e / kvol);
    int t = (int)(e - (long long)i * kvol);
    int kz = t / (g.k[1] * g.k[2]), ky = (t / g.k[2]) % g.k[1], kx = t % g.k[2];
    int4 c = reinterpret_cast<const int4*>(coors_in)[i];
    int vz = c.y + g.p[0] - kz, vy = c.z + g.p[1] - ky, vx = c.w + g.p[2] - kx;
    if (vz < 0 || vy < 0 || vx < 0) continue;
    if (vz % g.s[0] || vy % g.s[1] || vx % g.s[2]) continue;
    int oz = vz / g.s[0], oy = vy / g.s[1], ox = vx / g.s[2];
    if (oz >= g.Do || oy >= g.Ho || ox >= g.Wo) continue;
    bool ins;
    if (*overflow) continue;   // capacity already exceeded: stop filling the table (keeps probing short)
    int s = hash_insert(hkeys_out, hmask_out, lin_key(c.x, oz, oy, ox, g.Do, g.Ho, g.Wo), &ins);
    if (s < 0) { *overflow = 1; continue; }
    // warp-aggregated row allocation: one atomicAdd per converged warp instead of one per new site
    const unsigned act = __activemask();
    const unsigned vot = __ballot_sync(act, ins);
    if (ins) {
      const int lane = threadIdx.x & 31;
      const int leader = __ffs(vot) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(n_out_dev, __popc(vot));
      base = __shfl_sync(vot, base, leader);
      const int row = base + __popc(vot & ((1u << lane) - 1u));
      if (row < cap_out) {
        reinterpret_cast<int4*>(coors_out)[row] = make_int4(c.x, oz, oy, ox);
        hvals_out[s] = row;
      } else {
        *overflow = 1;
        hvals_out[s] = -1;
      }
    }
  }
}

__global__ void sp_clamp_count_kernel(int* n_dev, int cap) {
  if (*n_dev > cap) *n_dev = cap;
}

__global__ void sp_down_map_kernel(const int* __restrict__ coors_out, const int* __restrict__ n_out_dev, int cap_out,
                                   Down g, const uint32_t* __restrict__ hkeys_in, const int* __restrict__ hvals_in,
                                   int hmask_in, int* nbr) {
  int n = min(*n_out_dev, cap_out);
  int kvol = g.k[0] * g.k[1] * g.k[2];
  long long total = (long long)n * kvol;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int t = (int)(e / n);
    int o = (int)(e - (long long)t * n);
    int kz = t / (g.k[1] * g.k[2]), ky = (t / g.k[2]) % g.k[1], kx = t % g.k[2];
    int4 c = reinterpret_cast<const int4*>(coors_out)[o];
    int z = c.y * g.s[0] - g.p[0] + kz, y = c.z * g.s[1] - g.p[1] + ky, x = c.w * g.s[2] - g.p[2] + kx;
    int r = -1;
    if (z >= 0 && z < g.D && y >= 0 && y < g.H && x >= 0 && x < g.W) {
      int s = hash_find(hkeys_in, hmask_in, lin_key(c.x, z, y, x, g.D, g.H, g.W));
      if (s >= 0) r = hvals_in[s];
    }
    nbr[(size_t)t * cap_out + o] = r;
  }
}

__global__ void sp_bev_offsets_kernel(const int* __restrict__ coors, const int* __restrict__ n_dev, int cap, int H, int W,
                                      int ld, int C, int* off) {
  int n = min(*n_dev, cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = reinterpret_cast<const int4*>(coors)[i];
    off[i] = ((c.x * H + c.z) * W + c.w) * ld + c.y * C;
  }
}

static inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
static inline int persistent_blocks(long long work, int threads) {
  long long nb = (work + threads - 1) / threads;
  long long cap = (long long)num_sms() * 16;
  return (int)(nb < 1 ? 1 : (nb > cap ? cap : nb));
}

}  // namespace ff3d

extern "C" int ff3d_sp_hash_build(const int* coors, const int* n_dev, int cap, int batch, int D, int H, int W,
                                  uint32_t* hkeys, int* hvals, int hsize, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(is_pow2(hsize) && hsize >= 2 * cap, "sp_hash_build: hsize=%d must be a power of two >= 2*cap=%d", hsize,
               2 * cap);
  FF3D_REQUIRE((long long)batch * D * H * W < 0xFFFFFFFFLL, "sp_hash_build: grid too large for 32-bit keys");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(hkeys, 0xFF, sizeof(uint32_t) * (size_t)hsize, st);
  sp_hash_build_kernel<<<persistent_blocks(cap, 256), 256, 0, st>>>(coors, n_dev, cap, D, H, W, hkeys, hvals, hsize - 1);
  return check_launch("ff3d_sp_hash_build");
}

extern "C" int ff3d_sp_subm_map(const int* coors, const int* n_dev, int cap, int batch, int D, int H, int W,
                                const uint32_t* hkeys, const int* hvals, int hsize, int* nbr, ff3d_stream_t stream) {
  using namespace ff3d;
  (void)batch;
  FF3D_REQUIRE(is_pow2(hsize), "sp_subm_map: hsize must be a power of two");
  sp_subm_map_kernel<<<persistent_blocks((long long)cap * 27, 256), 256, 0, as_stream(stream)>>>(
      coors, n_dev, cap, D, H, W, hkeys, hvals, hsize - 1, nbr);
  return check_launch("ff3d_sp_subm_map");
}

extern "C" int ff3d_sp_down_build(const int* coors_in, const int* n_in_dev, int cap_in, int batch, int D, int H, int W,
                                  const uint32_t* hkeys_in, const int* hvals_in, int hsize_in, const int* k3,
                                  const int* s3, const int* p3, int* coors_out, int* n_out_dev, int cap_out, int Do,
                                  int Ho, int Wo, uint32_t* hkeys_out, int* hvals_out, int hsize_out, int* nbr_out,
                                  int* overflow_dev, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(is_pow2(hsize_in) && is_pow2(hsize_out) && hsize_out >= 2 * cap_out, "sp_down_build: bad hash sizes");
  FF3D_REQUIRE((long long)batch * Do * Ho * Wo < 0xFFFFFFFFLL, "sp_down_build: grid too large for 32-bit keys");
  Down g;
  for (int a = 0; a < 3; ++a) { g.k[a] = k3[a]; g.s[a] = s3[a]; g.p[a] = p3[a]; }
  g.D = D; g.H = H; g.W = W; g.Do = Do; g.Ho = Ho; g.Wo = Wo;
  int kvol = k3[0] * k3[1] * k3[2];
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(hkeys_out, 0xFF, sizeof(uint32_t) * (size_t)hsize_out, st);
  cudaMemsetAsync(n_out_dev, 0, sizeof(int), st);
  sp_down_sites_kernel<<<persistent_blocks((long long)cap_in * kvol, 256), 256, 0, st>>>(
      coors_in, n_in_dev, cap_in, g, coors_out, n_out_dev, cap_out, hkeys_out, hvals_out, hsize_out - 1, overflow_dev);
  sp_clamp_count_kernel<<<1, 1, 0, st>>>(n_out_dev, cap_out);
  sp_down_map_kernel<<<persistent_blocks((long long)cap_out * kvol, 256), 256, 0, st>>>(
      coors_out, n_out_dev, cap_out, g, hkeys_in, hvals_in, hsize_in - 1, nbr_out);
  return check_launch("ff3d_sp_down_build");
}

extern "C" int ff3d_sp_bev_offsets(const int* coors, const int* n_dev, int cap, int H, int W, int ld, int C, int* off,
                                   ff3d_stream_t stream) {
  using namespace ff3d;
  sp_bev_offsets_kernel<<<persistent_blocks(cap, 256), 256, 0, as_stream(stream)>>>(coors, n_dev, cap, H, W, ld, C, off);
  return check_launch("ff3d_sp_bev_offsets");
}
