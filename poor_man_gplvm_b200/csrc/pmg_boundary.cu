// Boundary messages of a time-sharded pass, packed and unpacked by one small kernel each.
//
// A rank owns a contiguous block of bins (reference decoder.py:258-332 walks its chunks sequentially; the
// block boundary is a chunk boundary of that walk).  After a forward pass it sends its first true message to the
// left neighbour and, to the right one, its last true message plus this pass's message at the bin where the
// neighbour's boundary chain starts its next warm-up; what arrives becomes seam truth, warm start and the row behind
// the block.  Host-side these were ~20 tiny tensor ops per pass -- at 125 000 bins per rank the host, not the GPU,
// bounded the EM iteration.  Messages travel in one fixed-size buffer per rank (all-gathered over the ranks):
//   forward  buffer [8K]: to_left = [first (2K) | unused (2K)], to_right = [last (2K) | warm (2K)]
#include "pmg_common.cuh"

namespace pmg {

constexpr int BD_THREADS = 256;

__device__ __forceinline__ float bd_block_max(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float m = red[0];
  for (int w = 1; w < BD_THREADS / 32; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  return m;
}
__device__ __forceinline__ float bd_block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < BD_THREADS / 32; ++w) s += red[w];
  __syncthreads();
  return s;
}

// warm_mode 0: no warm message (zeros); 1: warm_src is the [2K] message; 2: compact row of the filtered posterior:
// alpha[0,:] = ax_row[0..K), alpha[1,x] = ax_row[K] * exp2(sc2 * (ll_row[x] - max ll_row))
__global__ void __launch_bounds__(BD_THREADS)
boundary_pack_fwd_kernel(int K, const float* __restrict__ first, const float* __restrict__ last, int warm_mode,
                         const float* __restrict__ warm_src, const float* __restrict__ ax_row,
                         const float* __restrict__ ll_row, float sc2, float* __restrict__ out) {
  __shared__ float red[BD_THREADS / 32];
  const int K2 = 2 * K;
  for (int i = threadIdx.x; i < K2; i += BD_THREADS) {
    out[i] = first ? first[i] : 0.f;
    out[K2 + i] = 0.f;
    out[2 * K2 + i] = last ? last[i] : 0.f;
  }
  float* warm = out + 3 * K2;
  if (warm_mode == 1) {
    for (int i = threadIdx.x; i < K2; i += BD_THREADS) warm[i] = warm_src[i];
  } else if (warm_mode == 2) {
    float m = -INFINITY;
    for (int x = threadIdx.x; x < K; x += BD_THREADS) m = fmaxf(m, ll_row[x]);
    m = bd_block_max(m, red);
    const float a1s = ax_row[K];
    for (int x = threadIdx.x; x < K; x += BD_THREADS) {
      warm[x] = ax_row[x];
      warm[K + x] = a1s * exp2f((ll_row[x] - m) * sc2);
    }
  } else {
    for (int i = threadIdx.x; i < K2; i += BD_THREADS) warm[i] = 0.f;
  }
}

// from_left  [4K] = the left neighbour's to_right:  seam truth in front of this block, warm start of chain 0
// from_right [4K] = the right neighbour's to_left:  its first true message = the row behind this block
//   general layout: alpha_stop[2K] = message;  compact layout: ax_stop[0..K) = message[0,:] and
//   ax_stop[K] = sum(message[1,:]) / sum_x exp2(sc2 * (ll_stop[x] - max ll_stop))  (the scalar the kernels expand)
__global__ void __launch_bounds__(BD_THREADS)
boundary_unpack_fwd_kernel(int K, const float* __restrict__ from_left, const float* __restrict__ from_right,
                           float* __restrict__ fwd_end0, float* __restrict__ fwarm0, int compact,
                           float* __restrict__ ax_stop, const float* __restrict__ ll_stop, float sc2,
                           float* __restrict__ alpha_stop) {
  __shared__ float red[BD_THREADS / 32];
  const int K2 = 2 * K;
  if (from_left) {
    for (int i = threadIdx.x; i < K2; i += BD_THREADS) {
      fwd_end0[i] = from_left[i];
      if (fwarm0) fwarm0[i] = from_left[K2 + i];
    }
  }
  if (from_right) {
    if (!compact) {
      for (int i = threadIdx.x; i < K2; i += BD_THREADS) alpha_stop[i] = from_right[i];
    } else {
      float m = -INFINITY;
      for (int x = threadIdx.x; x < K; x += BD_THREADS) m = fmaxf(m, ll_stop[x]);
      m = bd_block_max(m, red);
      float se = 0.f, s1 = 0.f;
      for (int x = threadIdx.x; x < K; x += BD_THREADS) {
        se += exp2f((ll_stop[x] - m) * sc2);
        s1 += from_right[K + x];
        ax_stop[x] = from_right[x];
      }
      se = bd_block_sum(se, red);
      s1 = bd_block_sum(s1, red);
      if (threadIdx.x == 0) ax_stop[K] = s1 / se;
    }
  }
}

}  // namespace pmg

extern "C" int pmg_boundary_pack_fwd(int K, const float* first, const float* last, int warm_mode,
                                     const float* warm_src, const float* ax_row, const float* ll_row,
                                     float likelihood_scale, float* out, pmg_stream_t stream) {
  if (K <= 0 || !out || warm_mode < 0 || warm_mode > 2) return PMG_ERR_BAD_ARG;
  if (warm_mode == 1 && !warm_src) return PMG_ERR_BAD_ARG;
  if (warm_mode == 2 && (!ax_row || !ll_row)) return PMG_ERR_BAD_ARG;
  pmg::boundary_pack_fwd_kernel<<<1, pmg::BD_THREADS, 0, (cudaStream_t)stream>>>(
      K, first, last, warm_mode, warm_src, ax_row, ll_row, likelihood_scale * 1.4426950408889634f, out);
  PMG_LAUNCH_CHECK();
  return PMG_OK;
}

extern "C" int pmg_boundary_unpack_fwd(int K, const float* from_left, const float* from_right, float* fwd_end0,
                                       float* fwarm0, int compact, float* ax_stop, const float* ll_stop,
                                       float likelihood_scale, float* alpha_stop, pmg_stream_t stream) {
  if (K <= 0) return PMG_ERR_BAD_ARG;
  if (from_left && !fwd_end0) return PMG_ERR_BAD_ARG;
  if (from_right && (compact ? (!ax_stop || !ll_stop) : !alpha_stop)) return PMG_ERR_BAD_ARG;
  if (!from_left && !from_right) return PMG_OK;
  pmg::boundary_unpack_fwd_kernel<<<1, pmg::BD_THREADS, 0, (cudaStream_t)stream>>>(
      K, from_left, from_right, fwd_end0, fwarm0, compact, ax_stop, ll_stop,
      likelihood_scale * 1.4426950408889634f, alpha_stop);
  PMG_LAUNCH_CHECK();
  return PMG_OK;
}
