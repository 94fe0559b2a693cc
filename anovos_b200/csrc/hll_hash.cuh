// Spark's XXH64 (seed 42) per-type value hashing + the HyperLogLog++ register update, shared by the stand-alone HLL kernel
// (hll.cu) and the run-summary kernel of the sort (sort.cu), which hashes only the DISTINCT values of the sorted keys.
#pragma once
#include "common.cuh"

namespace anv {

constexpr uint64_t XP1 = 0x9E3779B185EBCA87ull, XP2 = 0xC2B2AE3D27D4EB4Full, XP3 = 0x165667B19E3779F9ull,
                   XP4 = 0x85EBCA77C2B2AE63ull, XP5 = 0x27D4EB2F165667C5ull;
constexpr uint64_t HLL_SEED = 42;

__host__ __device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t h) {
  h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
  return h;
}
__host__ __device__ __forceinline__ uint64_t xxh64_int(uint32_t v) {
  uint64_t h = HLL_SEED + XP5 + 4ull;
  h ^= (uint64_t)v * XP1;
  h = rotl64(h, 23) * XP2 + XP3;
  return fmix64(h);
}
__host__ __device__ __forceinline__ uint64_t xxh64_long(uint64_t v) {
  uint64_t h = HLL_SEED + XP5 + 8ull;
  h ^= rotl64(v * XP2, 31) * XP1;
  h = rotl64(h, 27) * XP1 + XP4;
  return fmix64(h);
}

template <typename T> __device__ __forceinline__ uint64_t spark_hash(T x);
template <> __device__ __forceinline__ uint64_t spark_hash<int32_t>(int32_t x) { return xxh64_int((uint32_t)x); }
template <> __device__ __forceinline__ uint64_t spark_hash<int64_t>(int64_t x) { return xxh64_long((uint64_t)x); }
template <> __device__ __forceinline__ uint64_t spark_hash<float>(float x) {
  // floatToIntBits (NaN canonical 0x7fc00000), -0.0 normalised to 0.0
  const uint32_t b = (x != x) ? 0x7fc00000u : ((x == 0.0f) ? 0u : __float_as_uint(x));
  return xxh64_int(b);
}
template <> __device__ __forceinline__ uint64_t spark_hash<double>(double x) {
  const uint64_t b = (x != x) ? 0x7ff8000000000000ull : ((x == 0.0) ? 0ull : (uint64_t)__double_as_longlong(x));
  return xxh64_long(b);
}


// (register index, rho) of a hash for precision p
__device__ __forceinline__ void hll_slot(uint64_t h, int p, uint32_t& idx, uint32_t& rho) {
  idx = (uint32_t)(h >> (64 - p));
  const uint64_t w = (h << p) | (1ull << (p - 1));
  rho = (uint32_t)__clzll((long long)w) + 1u;
}

}  // namespace anv
