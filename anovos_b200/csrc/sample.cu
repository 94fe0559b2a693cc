// Spark's Bernoulli row sampler on the device: the default path of drift_detector.statistics
// (use_sampling=True -> data_sample -> idf.sample(False, fraction, seed) / stat.sampleBy; reference
// /root/reference/src/main/anovos/data_ingest/data_sampling.py:122-149, drift_stability/drift_detector.py:187-211).
//
// Spark (un-vendored; org.apache.spark.util.random) draws, per partition, one nextDouble() per row from
// XORShiftRandom(seed + partitionIndex) and keeps the row when x < fraction (BernoulliCellSampler with lb = 0; the
// rand(seed) column of DataFrameStatFunctions.sampleBy is the same generator, compared with the stratum's
// fraction).  XORShiftRandom: state = hashSeed(seed) (scala MurmurHash3.bytesHash of the 8 big-endian seed bytes,
// twice); next(bits): s ^= s << 21; s ^= s >>> 35; s ^= s << 4; return low `bits` bits;
// nextDouble() = ((next(26) << 27) + next(27)) * 2^-53.
//
// The recurrence is linear over GF(2), so the stream is generated in PARALLEL: thread t owns rows
// [t*R, (t+1)*R) and jumps to step 2*t*R by applying the precomputed matrices L^(2^k) (64 column images each) for
// the set bits of its step index, then iterates.  x < f  <=>  the 53-bit integer < ceil(f * 2^53): the caller passes
// integer thresholds, one per stratum, so no floating point is involved and the kept set is bit-exact.
#include "common.cuh"

namespace anv {

constexpr int SAMPLE_ROWS_PER_THREAD = 1024;   // multiple of 32: whole bitmap words per thread
constexpr int JUMP_POWERS = 44;                // steps < 2^44 (2 steps per row)

__constant__ uint64_t c_jump[JUMP_POWERS][64];

__host__ __device__ __forceinline__ uint64_t xorshift_step(uint64_t s) {
  s ^= s << 21;
  s ^= s >> 35;
  s ^= s << 4;
  return s;
}

__global__ void __launch_bounds__(128) spark_sample_kernel(int64_t n_rows, uint64_t s0, const int32_t* __restrict__ strata,
                                                           const uint64_t* __restrict__ thresholds, int n_strata,
                                                           uint32_t* __restrict__ keep) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r0 = t * SAMPLE_ROWS_PER_THREAD;
  if (r0 >= n_rows) return;
  uint64_t s = s0;
  uint64_t j = 2ull * (uint64_t)r0;            // next() calls consumed by the rows before r0
  for (int k = 0; j != 0 && k < JUMP_POWERS; ++k, j >>= 1) {
    if (j & 1ull) {
      uint64_t r = 0;
#pragma unroll 8
      for (int i = 0; i < 64; ++i) r ^= ((s >> i) & 1ull) ? c_jump[k][i] : 0ull;
      s = r;
    }
  }
  const uint64_t thr0 = thresholds[0];
  for (int w = 0; w < SAMPLE_ROWS_PER_THREAD / 32; ++w) {
    const int64_t base = r0 + (int64_t)w * 32;
    if (base >= n_rows) break;
    uint32_t word = 0;
#pragma unroll 4
    for (int b = 0; b < 32; ++b) {
      const int64_t row = base + b;
      if (row < n_rows) {
        s = xorshift_step(s);
        const uint64_t hi = s & ((1ull << 26) - 1ull);
        s = xorshift_step(s);
        const uint64_t lo = s & ((1ull << 27) - 1ull);
        const uint64_t x = (hi << 27) + lo;
        uint64_t thr = thr0;
        if (strata) {
          const int32_t g = strata[row];
          thr = (g >= 0 && g < n_strata) ? thresholds[g] : 0ull;   // unknown stratum: fraction 0.0 (fractions.getOrElse)
        }
        word |= (x < thr) ? (1u << b) : 0u;
      }
    }
    keep[base >> 5] = word;
  }
}

static uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
// scala.util.hashing.MurmurHash3.bytesHash over exactly 8 bytes
static uint32_t murmur3_bytes8(const uint8_t* d, uint32_t seed) {
  uint32_t h = seed;
  for (int i = 0; i < 8; i += 4) {
    uint32_t k = (uint32_t)d[i] | ((uint32_t)d[i + 1] << 8) | ((uint32_t)d[i + 2] << 16) | ((uint32_t)d[i + 3] << 24);
    k *= 0xcc9e2d51u; k = rotl32(k, 15); k *= 0x1b873593u;
    h ^= k; h = rotl32(h, 13); h = h * 5u + 0xe6546b64u;
  }
  h ^= 8u;
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}

}  // namespace anv

using namespace anv;

// XORShiftRandom.hashSeed(seed): the generator's initial state.
extern "C" uint64_t anv_spark_hash_seed(int64_t seed) {
  uint8_t bytes[8];
  for (int i = 0; i < 8; ++i) bytes[i] = (uint8_t)((uint64_t)seed >> (56 - 8 * i));   // ByteBuffer.putLong: big-endian
  const uint32_t low = murmur3_bytes8(bytes, 0x3c074a61u);                           // MurmurHash3.arraySeed
  const uint32_t high = murmur3_bytes8(bytes, low);
  return ((uint64_t)high << 32) | (uint64_t)low;
}

extern "C" int anv_spark_sample_mask(int64_t n_rows, int64_t seed, const int32_t* strata, const uint64_t* thresholds,
                                     int n_strata, uint32_t* keep, void* stream) {
  if (n_rows < 0 || n_strata < 1 || !thresholds || (n_rows > 0 && !keep)) { set_error("anv_spark_sample_mask: bad arguments"); return ANV_ERR_INVALID; }
  if (n_rows >= ((int64_t)1 << 42)) { set_error("anv_spark_sample_mask: n_rows >= 2^42"); return ANV_ERR_UNSUPPORTED; }
  if (n_rows == 0) return ANV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // L^(2^k) as column images; squaring: (A*A) e_i = A (A e_i)
  static uint64_t jump[JUMP_POWERS][64];
  static bool ready = false;
  if (!ready) {
    for (int i = 0; i < 64; ++i) jump[0][i] = xorshift_step(1ull << i);
    for (int k = 1; k < JUMP_POWERS; ++k)
      for (int i = 0; i < 64; ++i) {
        const uint64_t v = jump[k - 1][i];
        uint64_t r = 0;
        for (int b = 0; b < 64; ++b) if ((v >> b) & 1ull) r ^= jump[k - 1][b];
        jump[k][i] = r;
      }
    ready = true;   // idempotent: a concurrent first call computes the same table
  }
  ANV_CUDA(cudaMemcpyToSymbolAsync(c_jump, jump, sizeof(jump), 0, cudaMemcpyHostToDevice, st));
  const int64_t threads = (n_rows + SAMPLE_ROWS_PER_THREAD - 1) / SAMPLE_ROWS_PER_THREAD;
  const unsigned blocks = (unsigned)((threads + 127) / 128);
  spark_sample_kernel<<<blocks, 128, 0, st>>>(n_rows, anv_spark_hash_seed(seed), strata, thresholds, n_strata, keep);
  ANV_CUDA(cudaGetLastError());
  return ANV_OK;
}
