// Synthetic column generator for bench.py / tests: counter-based
// Philox4x32-10, key = (seed, column), counter = (row/4, stream).  Any chunk of any
// column is reproducible independently of the launch geometry.
#include "common.cuh"

namespace anv {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t b) { return ((b >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)

__device__ __forceinline__ void normals4(const uint4& r, float (&z)[4]) {
  const float r0 = sqrtf(-2.0f * __logf(u01(r.x))), r1 = sqrtf(-2.0f * __logf(u01(r.z)));
  float s0, c0, s1, c1;
  __sincosf(6.2831853071795865f * u01(r.y), &s0, &c0);
  __sincosf(6.2831853071795865f * u01(r.w), &s1, &c1);
  z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
}

template <typename OutT, typename F>
__device__ __forceinline__ void synth_loop(OutT* data, uint32_t* validity, int64_t n_rows, int64_t row0, uint2 key, float null_rate, F gen) {
  const int64_t n4 = (n_rows + 3) / 4;
  const int64_t n4_pad = (n4 + 31) & ~(int64_t)31;  // whole warps so the bitmap words are assembled uniformly
  const uint32_t null_thr = (uint32_t)fminf(null_rate * 4294967296.0f, 4294967040.0f);
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n4_pad; j += (int64_t)gridDim.x * blockDim.x) {
    OutT v[4];
    const int64_t g = j + (row0 >> 2);  // counter of the GLOBAL row group: chunk [row0, row0+n) of a frame == the frame's rows
    const uint4 r = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), 0u, 0u), key);
    gen(r, v);
    const int64_t row = j * 4;
    if (row + 3 < n_rows) {
      if (sizeof(OutT) == 4) *reinterpret_cast<uint4*>(data + row) = *reinterpret_cast<uint4*>(v);
    } else {
      for (int i = 0; i < 4; ++i) if (row + i < n_rows) data[row + i] = v[i];
    }
    if (validity) {
      const uint4 nr = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), 1u, 0u), key);
      uint32_t nib = 0;
      nib |= (nr.x >= null_thr && row + 0 < n_rows) ? 1u : 0u;
      nib |= (nr.y >= null_thr && row + 1 < n_rows) ? 2u : 0u;
      nib |= (nr.z >= null_thr && row + 2 < n_rows) ? 4u : 0u;
      nib |= (nr.w >= null_thr && row + 3 < n_rows) ? 8u : 0u;
      uint32_t w = nib << ((threadIdx.x & 7) * 4);
      w |= __shfl_xor_sync(ANV_FULL, w, 1);
      w |= __shfl_xor_sync(ANV_FULL, w, 2);
      w |= __shfl_xor_sync(ANV_FULL, w, 4);
      if ((threadIdx.x & 7) == 0 && row < n_rows) validity[row >> 5] = w;
    }
  }
}

__global__ void __launch_bounds__(256) synth_f32_kernel(float* data, uint32_t* validity, int64_t n_rows, int64_t row0, uint2 key, int family,
                                                        float a, float b, float null_rate) {
  synth_loop<float>(data, validity, n_rows, row0, key, null_rate, [=](const uint4& r, float (&v)[4]) {
    if (family == 0) {
      normals4(r, v);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = fmaf(v[i], b, a);
    } else if (family == 1) {
      normals4(r, v);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = __expf(fmaf(v[i], b, a));
    } else if (family == 2) {
      v[0] = fmaf(u01(r.x), b - a, a); v[1] = fmaf(u01(r.y), b - a, a);
      v[2] = fmaf(u01(r.z), b - a, a); v[3] = fmaf(u01(r.w), b - a, a);
    } else {
      const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // low 8 bits choose zero-inflation (70% exact zeros), high 24 bits the exponential
        const bool zero = (rr[i] & 0xffu) < 179u;
        v[i] = zero ? 0.0f : -b * __logf(u01(rr[i]));
      }
    }
  });
}

__global__ void __launch_bounds__(256) synth_codes_kernel(int32_t* data, uint32_t* validity, int64_t n_rows, int64_t row0, uint2 key,
                                                          int card, float zipf_s, float null_rate) {
  const float one_minus_s = 1.0f - zipf_s;
  const float span = __powf((float)card + 1.0f, one_minus_s) - 1.0f;
  const float inv = 1.0f / one_minus_s;
  synth_loop<int32_t>(data, validity, n_rows, row0, key, null_rate, [=](const uint4& r, int32_t (&v)[4]) {
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // inverse CDF of the continuous power law on [1, card+1): Zipf-like ranks
      const float x = __powf(fmaf(u01(rr[i]), span, 1.0f), inv);
      v[i] = min(max((int)x - 1, 0), card - 1);
    }
  });
}

static uint2 make_key(uint64_t seed, uint32_t column) {
  const uint64_t k = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(column + 1));
  return make_uint2((uint32_t)k, (uint32_t)(k >> 32));
}

}  // namespace anv

extern "C" int anv_synth_f32_rows(float* data, uint32_t* validity, int64_t n_rows, int64_t row0, uint64_t seed, uint32_t column,
                                  int family, float a, float b, float null_rate, void* stream) {
  if (!data || n_rows < 0 || row0 < 0 || (row0 & 31) || family < 0 || family > 3 || ((uintptr_t)data & 15)) {
    anv::set_error("anv_synth_f32: bad arguments (row0 must be a multiple of 32)");
    return ANV_ERR_INVALID;
  }
  if (n_rows == 0) return ANV_OK;
  const int64_t n4 = (n_rows + 3) / 4;
  const int blocks = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
  anv::synth_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(data, validity, n_rows, row0, anv::make_key(seed, column), family,
                                                                  a, b, null_rate);
  ANV_CUDA(cudaGetLastError());
  return ANV_OK;
}

extern "C" int anv_synth_f32(float* data, uint32_t* validity, int64_t n_rows, uint64_t seed, uint32_t column, int family,
                             float a, float b, float null_rate, void* stream) {
  return anv_synth_f32_rows(data, validity, n_rows, 0, seed, column, family, a, b, null_rate, stream);
}

extern "C" int anv_synth_codes_rows(int32_t* data, uint32_t* validity, int64_t n_rows, int64_t row0, uint64_t seed,
                                    uint32_t column, int cardinality, float zipf_s, float null_rate, void* stream) {
  if (!data || n_rows < 0 || row0 < 0 || (row0 & 31) || cardinality < 1 || zipf_s <= 1.0f || ((uintptr_t)data & 15)) {
    anv::set_error("anv_synth_codes: bad arguments (zipf_s must be > 1)");
    return ANV_ERR_INVALID;
  }
  if (n_rows == 0) return ANV_OK;
  const int64_t n4 = (n_rows + 3) / 4;
  const int blocks = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
  anv::synth_codes_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(data, validity, n_rows, row0, anv::make_key(seed, column),
                                                                    cardinality, zipf_s, null_rate);
  ANV_CUDA(cudaGetLastError());
  return ANV_OK;
}

extern "C" int anv_synth_codes(int32_t* data, uint32_t* validity, int64_t n_rows, uint64_t seed, uint32_t column,
                               int cardinality, float zipf_s, float null_rate, void* stream) {
  return anv_synth_codes_rows(data, validity, n_rows, 0, seed, column, cardinality, zipf_s, null_rate, stream);
}
