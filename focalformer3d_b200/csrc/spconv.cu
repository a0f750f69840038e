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

// ---- site creation without a global row counter (round 2) -----------------------------------------------------------
// sp_down_sites_kernel above hands out rows with one atomicAdd per warp on a single counter: ~250 us for 1.5 M proposals
// (ncu), the largest rulebook kernel.  Row numbers of a new level are arbitrary (the level is re-stored in tap-mask
// order right afterwards), so: (1) insert the proposed keys, nothing else; (2) count the occupied slots per 2048-slot
// block; (3) scan the block counts; (4) number the occupied slots in slot order and decode their coordinates from the key.
// Deterministic, no contended atomics.
constexpr int SITE_BLK = 2048;     // hash slots per block (256 threads x 8)

__global__ void sp_sites_insert_kernel(const int* __restrict__ coors_in, const int* __restrict__ n_in_dev, int cap_in, Down g,
                                       uint32_t* hkeys_out, int hmask_out, volatile int* overflow) {
  // one thread per input site: per axis the (<= 3) taps whose output coordinate is integral and in range, then the
  // (<= 27, typically 1-8) combinations -- no 64-bit index arithmetic, no wasted (site, tap) iterations
  const int n = min(*n_in_dev, cap_in);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int4 c = reinterpret_cast<const int4*>(coors_in)[i];
    const int in3[3] = {c.y, c.z, c.w};
    const int lim[3] = {g.Do, g.Ho, g.Wo};
    int o[3][3], cnt[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      cnt[a] = 0;
      for (int k = 0; k < g.k[a]; ++k) {
        const int v = in3[a] + g.p[a] - k;
        if (v < 0 || v % g.s[a]) continue;
        const int ov = v / g.s[a];
        if (ov < lim[a] && cnt[a] < 3) o[a][cnt[a]++] = ov;
      }
    }
    for (int a0 = 0; a0 < cnt[0]; ++a0)
      for (int a1 = 0; a1 < cnt[1]; ++a1)
        for (int a2 = 0; a2 < cnt[2]; ++a2) {
          bool ins;
          if (hash_insert(hkeys_out, hmask_out, lin_key(c.x, o[0][a0], o[1][a1], o[2][a2], g.Do, g.Ho, g.Wo), &ins) < 0) *overflow = 1;
        }
  }
}

__global__ void __launch_bounds__(256) sp_sites_count_kernel(const uint32_t* __restrict__ hkeys, int hsize, int* bcount) {
  __shared__ int sh[8];
  const int base = blockIdx.x * SITE_BLK + threadIdx.x * 8;
  int c = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) c += (base + k < hsize && hkeys[base + k] != kEmptyKey) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    bcount[blockIdx.x] = t;
  }
}

// single block: exclusive scan of the block counts in place, total -> bcount[nb]
__global__ void __launch_bounds__(1024) sp_sites_scan_kernel(int* bcount, int nb) {
  __shared__ int sh[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? bcount[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    const int incl = sh[threadIdx.x];
    if (i < nb) bcount[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) bcount[nb] = carry;
}

__global__ void __launch_bounds__(256) sp_sites_assign_kernel(const uint32_t* __restrict__ hkeys, int hsize,
                                                              const int* __restrict__ bcount, int nb, Down g, int* coors_out,
                                                              int* hvals, int* n_out_dev, int cap_out, int* overflow) {
  __shared__ int sh[256];
  const int base = blockIdx.x * SITE_BLK + threadIdx.x * 8;
  uint32_t key[8];
  int c = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    key[k] = base + k < hsize ? hkeys[base + k] : kEmptyKey;
    c += key[k] != kEmptyKey ? 1 : 0;
  }
  sh[threadIdx.x] = c;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  int row = bcount[blockIdx.x] + sh[threadIdx.x] - c;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (key[k] == kEmptyKey) continue;
    if (row < cap_out) {
      uint32_t r = key[k];
      const int x = (int)(r % (uint32_t)g.Wo); r /= (uint32_t)g.Wo;
      const int y = (int)(r % (uint32_t)g.Ho); r /= (uint32_t)g.Ho;
      const int z = (int)(r % (uint32_t)g.Do); r /= (uint32_t)g.Do;
      reinterpret_cast<int4*>(coors_out)[row] = make_int4((int)r, z, y, x);
      hvals[base + k] = row;
    } else {
      hvals[base + k] = -1;
    }
    ++row;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int total = bcount[nb];
    *n_out_dev = total < cap_out ? total : cap_out;
    if (total > cap_out) *overflow = 1;
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


// ---- mask-sorted rulebooks ------------------------------------------------------------------------------------------
// The gather-GEMM is output-stationary: a 128-row tile runs every tap in which ANY of its rows has a neighbour.  In
// hash-allocation order a tile's rows are spatially random and the OR of their tap masks is always full, so 2-9x more
// (row, tap) pairs are gathered, split and multiplied than exist (measured on the bench clouds: SubM levels 4.1 / 2.4 /
// 2.1 / 2.0x, strided convs 9.5 / 5.7 / 3.9x).  Sorting the rows by their tap-presence mask (rarest tap = most
// significant key bit, so the bits that vary inside a tile are the ones that are nearly always set) makes the tiles
// homogeneous: 1.5 / 1.2 / 1.2 / 1.2x and 1.8 / 1.5 / 1.5x.  SubM levels are PHYSICALLY stored in mask order (all
// SubM convs of the level share it); strided convs write through a row map.
struct TapOrder { unsigned char bit[27]; };     // key bit of tap t (natural tap index t = (kz*k1 + ky)*k2 + kx)

__device__ __forceinline__ uint32_t probe_taps(int4 c, const Down& g, const uint32_t* __restrict__ hkeys,
                                               const int* __restrict__ hvals, int hmask, int kvol, int* rows /*nullable*/) {
  uint32_t m = 0;
  int t = 0;
  for (int kz = 0; kz < g.k[0]; ++kz)
    for (int ky = 0; ky < g.k[1]; ++ky)
      for (int kx = 0; kx < g.k[2]; ++kx, ++t) {
        const int z = c.y * g.s[0] - g.p[0] + kz, y = c.z * g.s[1] - g.p[1] + ky, x = c.w * g.s[2] - g.p[2] + kx;
        int r = -1;
        if (z >= 0 && z < g.D && y >= 0 && y < g.H && x >= 0 && x < g.W) {
          const int s = hash_find(hkeys, hmask, lin_key(c.x, z, y, x, g.D, g.H, g.W));
          if (s >= 0) r = hvals[s];
        }
        if (r >= 0) m |= 1u << t;
        if (rows) rows[t] = r;
      }
  (void)kvol;
  return m;
}

// sort key of every output row: its tap mask with the bits permuted into rarity order
__global__ void sp_tap_keys_kernel(const int* __restrict__ coors_out, const int* __restrict__ n_out_dev, int cap_out,
                                   Down g, TapOrder ord, const uint32_t* __restrict__ hkeys_in,
                                   const int* __restrict__ hvals_in, int hmask_in, uint32_t* __restrict__ keys,
                                   int* __restrict__ nbr_u) {
  const int n = min(*n_out_dev, cap_out);
  const int kvol = g.k[0] * g.k[1] * g.k[2];
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
    const int4 c = reinterpret_cast<const int4*>(coors_out)[o];
    int rows[27];
    const uint32_t m = probe_taps(c, g, hkeys_in, hvals_in, hmask_in, kvol, nbr_u ? rows : nullptr);
    uint32_t key = 0;
    for (int t = 0; t < kvol; ++t) key |= ((m >> t) & 1u) << ord.bit[t];
    keys[o] = key;
    // the probe results are kept (rows in the CURRENT order): ff3d_sp_nbr_permute re-orders them after the sort instead of
    // probing the hash a second time
    if (nbr_u)
      for (int t = 0; t < kvol; ++t) nbr_u[(size_t)t * cap_out + o] = rows[t];
  }
}

// store a level in sorted order: coors_out[i] = coors_in[perm[i]]; the level's hash now answers with the NEW row index
__global__ void sp_level_permute_kernel(const int* __restrict__ coors_in, const int* __restrict__ perm,
                                        const int* __restrict__ n_dev, int cap, int D, int H, int W, int* coors_out,
                                        const uint32_t* __restrict__ hkeys, int* hvals, int hmask, int* __restrict__ inv) {
  const int n = min(*n_dev, cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int old = perm[i];
    const int4 c = reinterpret_cast<const int4*>(coors_in)[old];
    reinterpret_cast<int4*>(coors_out)[i] = c;
    const int s = hash_find(hkeys, hmask, lin_key(c.x, c.y, c.z, c.w, D, H, W));
    if (s >= 0) hvals[s] = i;
    if (inv) inv[old] = i;                                   // old row -> new row
  }
}

// Neighbour map in sorted tile order from the rows probed by sp_tap_keys_kernel (no second round of hash probes):
// nbr[t][j] = remap(nbr_u[t][perm[j]]), remap = inv (SubM: the level itself was re-ordered) or identity; plus the tile
// masks and the output row map like sp_nbr_build_kernel.  One 128-thread block per 128-row tile.
__global__ void __launch_bounds__(128) sp_nbr_permute_kernel(const int* __restrict__ nbr_u, const int* __restrict__ perm,
                                                             const int* __restrict__ inv, const int* __restrict__ coors_out,
                                                             const int* __restrict__ n_out_dev, int cap_out, int kvol,
                                                             int* __restrict__ nbr, uint32_t* __restrict__ tile_mask,
                                                             int* __restrict__ y_off, int y_mode, int ldy, int Hb, int Wb, int Cc) {
  __shared__ uint32_t wm[4];
  const int n = min(*n_out_dev, cap_out);
  const int n_tiles = (n + 127) >> 7;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int j = tile * 128 + threadIdx.x;
    uint32_t m = 0;
    if (j < n) {
      const int o = perm ? perm[j] : j;
      for (int t = 0; t < kvol; ++t) {
        int r = nbr_u[(size_t)t * cap_out + o];
        if (r >= 0) {
          if (inv) r = inv[r];
          m |= 1u << t;
        }
        nbr[(size_t)t * cap_out + j] = r;
      }
      if (y_mode == 1) y_off[j] = o;
      else if (y_mode == 2) {
        const int4 c = reinterpret_cast<const int4*>(coors_out)[o];
        y_off[j] = ((c.x * Hb + c.z) * Wb + c.w) * ldy + c.y * Cc;
      }
    }
    m = __reduce_or_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) tile_mask[tile] = wm[0] | wm[1] | wm[2] | wm[3];
    __syncthreads();
  }
}

// neighbour map in (sorted) tile order + the OR of the tap masks of every 128-row tile + the output row map.
// One 128-thread block per 128-row tile.  y_mode 0: no row map; 1: y_off[j] = source ROW (a row map); 2: NHWC BEV element
// offset ((b*Hb + y)*Wb + x)*ld + z*C of the source row's coordinates.
__global__ void __launch_bounds__(128) sp_nbr_build_kernel(const int* __restrict__ coors_out, const int* __restrict__ perm,
                                                           const int* __restrict__ n_out_dev, int cap_out, Down g,
                                                           const uint32_t* __restrict__ hkeys_in,
                                                           const int* __restrict__ hvals_in, int hmask_in, int* __restrict__ nbr,
                                                           uint32_t* __restrict__ tile_mask, int* __restrict__ y_off, int y_mode,
                                                           int ldy, int Cc) {
  __shared__ uint32_t wm[4];
  const int n = min(*n_out_dev, cap_out);
  const int kvol = g.k[0] * g.k[1] * g.k[2];
  const int n_tiles = (n + 127) >> 7;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int j = tile * 128 + threadIdx.x;
    uint32_t m = 0;
    if (j < n) {
      const int o = perm ? perm[j] : j;
      const int4 c = reinterpret_cast<const int4*>(coors_out)[o];
      int rows[27];
      m = probe_taps(c, g, hkeys_in, hvals_in, hmask_in, kvol, rows);
      for (int t = 0; t < kvol; ++t) nbr[(size_t)t * cap_out + j] = rows[t];
      if (y_mode == 1) y_off[j] = o;                 // output ROW of tile position j
      else if (y_mode == 2) y_off[j] = ((c.x * g.Ho + c.z) * g.Wo + c.w) * ldy + c.y * Cc;
    }
    m = __reduce_or_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) tile_mask[tile] = wm[0] | wm[1] | wm[2] | wm[3];
    __syncthreads();
  }
}

// dst[i, 0:cols] = src[perm[i], 0:cols]  (cols % 4 == 0, 16-byte aligned rows): level-1 voxel features into mask order
__global__ void sp_gather_rows_kernel(const float* __restrict__ src, int ld_src, const int* __restrict__ perm,
                                      const int* __restrict__ n_dev, int cap, float* __restrict__ dst, int ld_dst, int cols4) {
  const int n = min(*n_dev, cap);
  const long long total = (long long)n * cols4;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / cols4), c = (int)(e - (long long)i * cols4);
    reinterpret_cast<float4*>(dst + (size_t)i * ld_dst)[c] = __ldg(reinterpret_cast<const float4*>(src + (size_t)perm[i] * ld_src) + c);
  }
}

static TapOrder tap_order(const int* k3) {
  // rarest-first static order: taps far from the kernel centre are present least often (measured on LiDAR clouds:
  // centre 1.0, faces 0.36-0.64, edges 0.20-0.47, corners 0.13-0.37), vertical offsets rarer than horizontal ones.
  // rank 0 (rarest) gets the MOST significant key bit.
  TapOrder ord;
  const int kvol = k3[0] * k3[1] * k3[2];
  int score[27], idx[27];
  for (int t = 0; t < kvol; ++t) {
    const int kz = t / (k3[1] * k3[2]), ky = (t / k3[2]) % k3[1], kx = t % k3[2];
    const int dz = 2 * kz - (k3[0] - 1), dy = 2 * ky - (k3[1] - 1), dx = 2 * kx - (k3[2] - 1);   // doubled offsets
    const int az = dz < 0 ? -dz : dz, ay = dy < 0 ? -dy : dy, ax = dx < 0 ? -dx : dx;
    score[t] = (az + ay + ax) * 8 + az;
    idx[t] = t;
  }
  for (int a = 0; a < kvol; ++a)                      // stable selection sort, descending score
    for (int b = a + 1; b < kvol; ++b)
      if (score[idx[b]] > score[idx[a]]) { int tmp = idx[a]; idx[a] = idx[b]; idx[b] = tmp; }
  for (int r = 0; r < kvol; ++r) ord.bit[idx[r]] = (unsigned char)(kvol - 1 - r);
  for (int t = kvol; t < 27; ++t) ord.bit[t] = 0;
  return ord;
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

// ---- mask-sorted rulebooks (see the kernel comments above) ---------------------------------------------------------
static int fill_down(ff3d::Down& g, const int* k3, const int* s3, const int* p3, int D, int H, int W, int Do, int Ho, int Wo) {
  for (int a = 0; a < 3; ++a) { g.k[a] = k3[a]; g.s[a] = s3[a]; g.p[a] = p3[a]; }
  g.D = D; g.H = H; g.W = W; g.Do = Do; g.Ho = Ho; g.Wo = Wo;
  return k3[0] * k3[1] * k3[2];
}

extern "C" int ff3d_sp_tap_keys(const int* coors_out, const int* n_out_dev, int cap_out, int D, int H, int W,
                                const uint32_t* hkeys_in, const int* hvals_in, int hsize_in, const int* k3, const int* s3,
                                const int* p3, uint32_t* keys, int* nbr_unsorted, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(is_pow2(hsize_in), "sp_tap_keys: hsize must be a power of two");
  Down g;
  const int kvol = fill_down(g, k3, s3, p3, D, H, W, 0, 0, 0);
  FF3D_REQUIRE(kvol >= 1 && kvol <= 27, "sp_tap_keys: kernel volume %d not in 1..27", kvol);
  sp_tap_keys_kernel<<<persistent_blocks(cap_out, 128), 128, 0, as_stream(stream)>>>(
      coors_out, n_out_dev, cap_out, g, tap_order(k3), hkeys_in, hvals_in, hsize_in - 1, keys, nbr_unsorted);
  return check_launch("ff3d_sp_tap_keys");
}

extern "C" int ff3d_sp_level_permute(const int* coors_in, const int* perm, const int* n_dev, int cap, int D, int H, int W,
                                     int* coors_out, const uint32_t* hkeys, int* hvals, int hsize, int* inv,
                                     ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(is_pow2(hsize) && coors_in != coors_out, "sp_level_permute: bad arguments");
  sp_level_permute_kernel<<<persistent_blocks(cap, 256), 256, 0, as_stream(stream)>>>(coors_in, perm, n_dev, cap, D, H, W,
                                                                                    coors_out, hkeys, hvals, hsize - 1, inv);
  return check_launch("ff3d_sp_level_permute");
}

extern "C" int ff3d_sp_nbr_permute(const int* nbr_unsorted, const int* perm, const int* inv, const int* coors_out,
                                   const int* n_out_dev, int cap_out, int kvol, int* nbr, uint32_t* tile_mask, int* y_off,
                                   int y_mode, int ldy, int bev_h, int bev_w, int bev_c, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(nbr_unsorted && nbr && tile_mask && kvol >= 1 && kvol <= 27, "sp_nbr_permute: bad arguments");
  FF3D_REQUIRE(y_mode == 0 || y_off != nullptr, "sp_nbr_permute: y_off missing");
  FF3D_REQUIRE(y_mode != 2 || coors_out != nullptr, "sp_nbr_permute: BEV offsets need the output coordinates");
  const int tiles = cdiv(cap_out, 128);
  const int cap_blocks = num_sms() * 16;
  sp_nbr_permute_kernel<<<tiles < cap_blocks ? (tiles < 1 ? 1 : tiles) : cap_blocks, 128, 0, as_stream(stream)>>>(
      nbr_unsorted, perm, inv, coors_out, n_out_dev, cap_out, kvol, nbr, tile_mask, y_off, y_mode, ldy, bev_h, bev_w, bev_c);
  return check_launch("ff3d_sp_nbr_permute");
}

extern "C" int ff3d_sp_nbr_build(const int* coors_out, const int* perm, const int* n_out_dev, int cap_out, int D, int H,
                                 int W, const uint32_t* hkeys_in, const int* hvals_in, int hsize_in, const int* k3,
                                 const int* s3, const int* p3, int* nbr, uint32_t* tile_mask, int* y_off, int y_mode,
                                 int ldy, int bev_h, int bev_w, int bev_c, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(is_pow2(hsize_in) && nbr && tile_mask, "sp_nbr_build: bad arguments");
  FF3D_REQUIRE(y_mode == 0 || y_off != nullptr, "sp_nbr_build: y_off missing");
  Down g;
  const int kvol = fill_down(g, k3, s3, p3, D, H, W, 0, bev_h, bev_w);
  FF3D_REQUIRE(kvol >= 1 && kvol <= 27, "sp_nbr_build: kernel volume %d not in 1..27", kvol);
  const int tiles = cdiv(cap_out, 128);
  const int cap_blocks = num_sms() * 16;
  sp_nbr_build_kernel<<<tiles < cap_blocks ? (tiles < 1 ? 1 : tiles) : cap_blocks, 128, 0, as_stream(stream)>>>(
      coors_out, perm, n_out_dev, cap_out, g, hkeys_in, hvals_in, hsize_in - 1, nbr, tile_mask, y_off, y_mode, ldy, bev_c);
  return check_launch("ff3d_sp_nbr_build");
}

extern "C" int ff3d_sp_down_sites(const int* coors_in, const int* n_in_dev, int cap_in, int batch, int D, int H, int W,
                                  const int* k3, const int* s3, const int* p3, int* coors_out, int* n_out_dev, int cap_out,
                                  int Do, int Ho, int Wo, uint32_t* hkeys_out, int* hvals_out, int hsize_out,
                                  int* overflow_dev, int* scratch, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(is_pow2(hsize_out) && hsize_out >= 2 * cap_out, "sp_down_sites: bad hash size");
  FF3D_REQUIRE((long long)batch * Do * Ho * Wo < 0xFFFFFFFFLL, "sp_down_sites: grid too large for 32-bit keys");
  FF3D_REQUIRE(scratch != nullptr, "sp_down_sites: scratch (ff3d_sp_down_sites_scratch_ints(hsize) ints) is missing");
  Down g;
  const int kvol = fill_down(g, k3, s3, p3, D, H, W, Do, Ho, Wo);
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(hkeys_out, 0xFF, sizeof(uint32_t) * (size_t)hsize_out, st);
  sp_sites_insert_kernel<<<persistent_blocks(cap_in, 128), 128, 0, st>>>(coors_in, n_in_dev, cap_in, g, hkeys_out,
                                                                                         hsize_out - 1, overflow_dev);
  const int nb = cdiv(hsize_out, SITE_BLK);
  sp_sites_count_kernel<<<nb, 256, 0, st>>>(hkeys_out, hsize_out, scratch);
  sp_sites_scan_kernel<<<1, 1024, 0, st>>>(scratch, nb);
  sp_sites_assign_kernel<<<nb, 256, 0, st>>>(hkeys_out, hsize_out, scratch, nb, g, coors_out, hvals_out, n_out_dev, cap_out,
                                            overflow_dev);
  return check_launch("ff3d_sp_down_sites");
}

extern "C" int ff3d_sp_down_sites_scratch_ints(int hsize) { return (hsize + ff3d::SITE_BLK - 1) / ff3d::SITE_BLK + 2; }

extern "C" int ff3d_sp_gather_rows(const float* src, int ld_src, const int* perm, const int* n_dev, int cap, float* dst,
                                   int ld_dst, int cols, ff3d_stream_t stream) {
  using namespace ff3d;
  FF3D_REQUIRE(cols % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0, "sp_gather_rows: cols and row strides must be multiples of 4");
  sp_gather_rows_kernel<<<persistent_blocks((long long)cap * (cols / 4), 256), 256, 0, as_stream(stream)>>>(
      src, ld_src, perm, n_dev, cap, dst, ld_dst, cols / 4);
  return check_launch("ff3d_sp_gather_rows");
}
