// Device-side initial latent posterior with jax.random's bit stream (threefry2x32, 20 rounds).
//
// Replaces reference core.py:571-583 (init_latent_posterior): posterior = uniform(key, (T, K)) * random_scale,
// row-normalised; jax 0.4.26 defaults (non-partitionable threefry): the n = T*K counters 0..n-1 (padded to
// even) are split in halves, block b encrypts (b, b + half) and element e of the output is word 0 of block e
// for e < half, word 1 of block e - half otherwise; float = bitcast((bits >> 9) | 0x3f800000) - 1.
// Host restatement + known-answer tests: poor_man_gplvm_b200/jaxprng.py.
#include <cuda_fp16.h>

#include "pmg_common.cuh"

namespace pmg {

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__device__ __forceinline__ void threefry2x32_20(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0];
  x1 += ks[1];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    if (i % 2 == 0) {
      x0 += x1; x1 = rotl32(x1, 13); x1 ^= x0;
      x0 += x1; x1 = rotl32(x1, 15); x1 ^= x0;
      x0 += x1; x1 = rotl32(x1, 26); x1 ^= x0;
      x0 += x1; x1 = rotl32(x1, 6); x1 ^= x0;
    } else {
      x0 += x1; x1 = rotl32(x1, 17); x1 ^= x0;
      x0 += x1; x1 = rotl32(x1, 29); x1 ^= x0;
      x0 += x1; x1 = rotl32(x1, 16); x1 ^= x0;
      x0 += x1; x1 = rotl32(x1, 24); x1 ^= x0;
    }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
  }
}

// element e of jax.random.bits(key, (n,)) as a float in [0,1)
__device__ __forceinline__ float jax_uniform_at(uint32_t k0, uint32_t k1, uint64_t e, uint64_t n, uint64_t half) {
  uint32_t x0, x1;
  const bool first = e < half;
  const uint64_t b = first ? e : e - half;
  x0 = (uint32_t)b;
  const uint64_t c1 = b + half;
  x1 = c1 < n ? (uint32_t)c1 : 0u;              // pad element of an odd-sized counter array
  threefry2x32_20(k0, k1, x0, x1);
  const uint32_t bits = first ? x0 : x1;
  return __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
}

// one warp per local row; rows [t_offset, t_offset + T) of a global [T_total, K] draw
template <int MAXQ>
__global__ void threefry_posterior_init_kernel(int64_t T, int K, int64_t t_offset, int64_t T_total, uint32_t k0,
                                               uint32_t k1, float random_scale, float* __restrict__ post,
                                               int64_t ldp, float* __restrict__ logpost, int64_t ldl,
                                               __half* __restrict__ g16, int64_t ldg, int64_t piece_stride,
                                               double* __restrict__ tw) {
  extern __shared__ double tw_s[];               // [K] per-CTA partial column sums
  for (int k = threadIdx.x; k < K; k += blockDim.x) tw_s[k] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const uint64_t n = (uint64_t)T_total * (uint64_t)K;
  const uint64_t half = (n + 1) / 2;
  for (int64_t t = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); t < T; t += (int64_t)gridDim.x * wpb) {
    const uint64_t e0 = (uint64_t)(t_offset + t) * (uint64_t)K;
    float v[MAXQ];
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < MAXQ; ++q) {
      const int k = lane + 32 * q;
      v[q] = 0.f;
      if (k < K) {
        v[q] = jax_uniform_at(k0, k1, e0 + k, n, half) * random_scale;
        s += v[q];
      }
    }
    s = warp_sum(s);
#pragma unroll
    for (int q = 0; q < MAXQ; ++q) {
      const int k = lane + 32 * q;
      if (k < K) {
        const float pv = v[q] / s;
        if (post) post[(size_t)t * ldp + k] = pv;
        if (logpost) logpost[(size_t)t * ldl + k] = logf(pv);      // log(0) = -inf = the reference's -1e40 in fp32
        if (g16) {
          const __half h = __float2half_rn(pv);
          g16[(size_t)t * ldg + k] = h;
          g16[(size_t)piece_stride + (size_t)t * ldg + k] = __float2half_rn(pv - __half2float(h));
        }
        if (tw) atomicAdd(&tw_s[k], (double)pv);
      }
    }
  }
  __syncthreads();
  if (tw)
    for (int k = threadIdx.x; k < K; k += blockDim.x) atomicAdd(&tw[k], tw_s[k]);
}

}  // namespace pmg

extern "C" int pmg_threefry_posterior_init(int64_t T, int K, int64_t t_offset, int64_t T_total, uint32_t key0,
                                           uint32_t key1, float random_scale, float* post, int64_t ldp,
                                           float* logpost, int64_t ldl, void* g16, int64_t ldg,
                                           int64_t piece_stride, double* tw, pmg_stream_t stream) {
  if (T <= 0 || K <= 0 || t_offset < 0 || T_total < t_offset + T) return PMG_ERR_BAD_ARG;
  if ((post && ldp < K) || (logpost && ldl < K) || (g16 && ldg < K)) return PMG_ERR_BAD_ARG;
  if ((uint64_t)T_total * (uint64_t)K > 0xFFFFFFFFull) return PMG_ERR_UNSUPPORTED_SHAPE;   // one 32-bit counter block
  if (K > 32 * 64) return PMG_ERR_UNSUPPORTED_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  if (tw) PMG_CUDA_CHECK(cudaMemsetAsync(tw, 0, sizeof(double) * K, st));
  const int threads = 256;
  int64_t blocks = (T + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const size_t smem = sizeof(double) * K;
#define PMG_PRNG_CASE(Q)                                                                                        \
  if (K <= 32 * Q) {                                                                                            \
    pmg::threefry_posterior_init_kernel<Q><<<(int)blocks, threads, smem, st>>>(                                 \
        T, K, t_offset, T_total, key0, key1, random_scale, post, ldp, logpost, ldl, (__half*)g16, ldg,          \
        piece_stride, tw);                                                                                      \
    PMG_LAUNCH_CHECK();                                                                                         \
    return PMG_OK;                                                                                              \
  }
  PMG_PRNG_CASE(4) PMG_PRNG_CASE(8) PMG_PRNG_CASE(16) PMG_PRNG_CASE(32) PMG_PRNG_CASE(64)
#undef PMG_PRNG_CASE
  return PMG_ERR_UNSUPPORTED_SHAPE;
}
