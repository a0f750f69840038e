// Shared helpers for libff3d.so (sm_100a).  No torch headers anywhere in this library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/ff3d.h"

namespace ff3d {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return FF3D_ECUDA;
  }
  return FF3D_OK;
}

#define FF3D_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ff3d::set_error(__VA_ARGS__);        \
      return FF3D_EINVAL;                  \
    }                                      \
  } while (0)

inline cudaStream_t as_stream(ff3d_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == FF3D_ACT_RELU) return fmaxf(v, 0.f);
  if (act == FF3D_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
  return v;
}

// ---- open-addressing hash over linearised voxel coordinates (uint32 keys, 0xFFFFFFFF = empty)
constexpr uint32_t kEmptyKey = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t hash_u32(uint32_t k) {
  k ^= k >> 16; k *= 0x7feb352dU; k ^= k >> 15; k *= 0x846ca68bU; k ^= k >> 16;
  return k;
}
// returns slot of key (inserting if absent); *inserted tells whether this call created it.
// Probing is bounded by the table size: -1 means the table is full (capacity overflow), never a hang.
__device__ __forceinline__ int hash_insert(uint32_t* keys, int mask, uint32_t key, bool* inserted) {
  uint32_t s = hash_u32(key) & mask;
  *inserted = false;
  for (int probe = 0; probe <= mask; ++probe) {
    uint32_t prev = atomicCAS(&keys[s], kEmptyKey, key);
    if (prev == kEmptyKey) { *inserted = true; return (int)s; }
    if (prev == key) return (int)s;
    s = (s + 1) & mask;
  }
  return -1;
}
__device__ __forceinline__ int hash_find(const uint32_t* __restrict__ keys, int mask, uint32_t key) {
  uint32_t s = hash_u32(key) & mask;
  for (int probe = 0; probe <= mask; ++probe) {
    uint32_t k = keys[s];
    if (k == key) return (int)s;
    if (k == kEmptyKey) return -1;
    s = (s + 1) & mask;
  }
  return -1;
}

}  // namespace ff3d
