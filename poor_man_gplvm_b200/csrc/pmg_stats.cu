// Time-reduction GEMM  C[m,n] = sum_t A[t,m] * B[t,n]  and the transition-count epilogue.
//
// Replaces reference poor_man_gplvm/fit_tuning_helper.py:28-42 (get_statistics:
// A = posterior over latent bins, B = spike counts) and the per-step
// logaddexp accumulation of decoder.py:215-221 through the identity
//   sum_t xi_t[d,d',x,x'] = M[d,d'] P_{d'}[x,x'] sum_t alpha_t[d,x] r_{t+1}[d',x']   (SURVEY S4).
#include "pmg_common.cuh"

namespace pmg {

constexpr int AT_BM = 64, AT_BN = 64, AT_BK = 16;

// CUDA-core fp32 tiles, deterministic split over time: partial[z][m][n]
__global__ void __launch_bounds__(256) atb_simt_kernel(int64_t T, int M, int N, const float* __restrict__ A,
                                                       int64_t lda, const float* __restrict__ B, int64_t ldb,
                                                       float* __restrict__ partial, int64_t t_per_split) {
  __shared__ float As[AT_BK][AT_BM + 4];
  __shared__ float Bs[AT_BK][AT_BN + 4];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.x * AT_BM, n0 = blockIdx.y * AT_BN;
  const int64_t tb = (int64_t)blockIdx.z * t_per_split;
  int64_t te = tb + t_per_split;
  if (te > T) te = T;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t t0 = tb; t0 < te; t0 += AT_BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int r = idx >> 6, cc = idx & 63;
      const int64_t t = t0 + r;
      As[r][cc] = (t < te && m0 + cc < M) ? A[(size_t)t * lda + m0 + cc] : 0.f;
      Bs[r][cc] = (t < te && n0 + cc < N) ? B[(size_t)t * ldb + n0 + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < AT_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = partial + (size_t)blockIdx.z * M * N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) out[(size_t)m * N + n] = acc[i][j];
    }
  }
}

// fixed-order reduction over the time splits (fp64 accumulate): run-to-run deterministic
__global__ void split_reduce_kernel(int splits, int M, int N, const float* __restrict__ partial,
                                    float* __restrict__ C, int64_t ldc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  double s = 0.0;
  for (int z = 0; z < splits; ++z) s += (double)partial[(size_t)z * M * N + i];
  C[(size_t)(i / N) * ldc + (i % N)] = (float)s;
}

__global__ void xi_finalize_kernel(int K, const float* __restrict__ G, const float* __restrict__ logP,
                                   float lm00, float lm01, float lm10, float lm11,
                                   float* __restrict__ log_acc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t KK = (int64_t)K * K;
  if (i >= 4 * KK) return;
  const int d = (int)(i / (2 * KK));
  const int dn = (int)((i / KK) & 1);
  const int x = (int)((i % KK) / K), xn = (int)(i % K);
  const float g = G[((size_t)d * K + x) * (2 * K) + (size_t)dn * K + xn];
  const float lm = d == 0 ? (dn == 0 ? lm00 : lm01) : (dn == 0 ? lm10 : lm11);
  // floor keeps rows of masked latent bins finite (the reference's log-space accumulator never reaches -inf)
  log_acc[i] = lm + logP[(size_t)dn * KK + (size_t)x * K + xn] + logf(fmaxf(g, 1e-37f));
}

static int atb_splits(int64_t T, int M, int N) {
  const int tiles = cdiv(M, AT_BM) * cdiv(N, AT_BN);
  int splits = (148 * 4 + tiles - 1) / tiles;
  const int64_t max_splits = (T + 255) / 256;
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  if (splits > 512) splits = 512;
  return splits;
}

}  // namespace pmg

int pmg_atb_tc_launch(int64_t T, int M, int N, const float* A, int64_t lda, const float* B, int64_t ldb,
                      float* C, int64_t ldc, void* ws, int64_t ws_bytes, cudaStream_t st);  // pmg_gemm_tc.cu
int64_t pmg_atb_tc_workspace_bytes(int64_t T, int M, int N);

namespace pmg {
// dst[0] = sum_i src[i * stride] in fp64, fixed summation order: per-block partials, the last block to finish adds
// them up in block order and re-arms the ticket.  (The log marginal of a pass = sum_t lmr_t, decoder.py:170,186.)
constexpr int SS_BLOCKS = 128;
__global__ void __launch_bounds__(256) strided_sum_kernel(int64_t n, const float* __restrict__ src, int64_t stride,
                                                          float* __restrict__ dst, unsigned* ticket, double* part) {
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += (double)src[(size_t)i * stride];
  acc = warp_sum_d(acc);
  __shared__ double sm[8];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += sm[w];
    part[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += ((volatile double*)part)[b];
    dst[0] = (float)s;
    *ticket = 0u;
  }
}
}  // namespace pmg

extern "C" int64_t pmg_strided_sum_workspace_bytes(void) { return 16 + 8 * pmg::SS_BLOCKS; }

extern "C" int pmg_strided_sum(int64_t n, const float* src, int64_t stride, float* dst, void* workspace,
                               pmg_stream_t stream) {
  if (n <= 0 || !src || !dst || !workspace || stride < 1 || ((uintptr_t)workspace & 7)) return PMG_ERR_BAD_ARG;
  int grid = (int)((n + 255) / 256);
  if (grid > pmg::SS_BLOCKS) grid = pmg::SS_BLOCKS;
  pmg::strided_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, src, stride, dst, (unsigned*)workspace,
                                                                 (double*)((char*)workspace + 16));
  PMG_LAUNCH_CHECK();
  return PMG_OK;
}

extern "C" int64_t pmg_atb_workspace_bytes(int64_t T, int M, int N, int impl) {
  if (T <= 0 || M <= 0 || N <= 0) return 0;
  int64_t simt = (int64_t)pmg::atb_splits(T, M, N) * M * N * (int64_t)sizeof(float);
  if (impl == 1) return simt;
  int64_t tc = pmg_atb_tc_workspace_bytes(T, M, N);
  return tc > simt ? tc : simt;
}

extern "C" int pmg_atb(int64_t T, int M, int N, const float* A, int64_t lda, const float* B, int64_t ldb,
                       float* C, int64_t ldc, void* workspace, int64_t workspace_bytes, int impl,
                       pmg_stream_t stream) {
  if (T <= 0 || M <= 0 || N <= 0 || !A || !B || !C || lda < M || ldb < N || ldc < N) return PMG_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == 0) {
    int rc = pmg_atb_tc_launch(T, M, N, A, lda, B, ldb, C, ldc, workspace, workspace_bytes, st);
    if (rc != PMG_ERR_UNSUPPORTED_SHAPE) return rc;
  }
  const int splits = pmg::atb_splits(T, M, N);
  if (!workspace || workspace_bytes < (int64_t)splits * M * N * (int64_t)sizeof(float)) return PMG_ERR_WORKSPACE;
  const int64_t tps = ((T + splits - 1) / splits + pmg::AT_BK - 1) / pmg::AT_BK * pmg::AT_BK;
  dim3 grid(pmg::cdiv(M, pmg::AT_BM), pmg::cdiv(N, pmg::AT_BN), splits);
  pmg::atb_simt_kernel<<<grid, 256, 0, st>>>(T, M, N, A, lda, B, ldb, (float*)workspace, tps);
  PMG_LAUNCH_CHECK();
  pmg::split_reduce_kernel<<<pmg::cdiv((int64_t)M * N, 256), 256, 0, st>>>(splits, M, N, (const float*)workspace, C, ldc);
  PMG_LAUNCH_CHECK();
  return PMG_OK;
}

extern "C" int pmg_xi_finalize(int K, const float* G, const float* logP, const float* logM, float* log_acc,
                               pmg_stream_t stream) {
  if (K <= 0 || !G || !logP || !logM || !log_acc) return PMG_ERR_BAD_ARG;
  pmg::xi_finalize_kernel<<<pmg::cdiv(4 * (int64_t)K * K, 256), 256, 0, (cudaStream_t)stream>>>(
      K, G, logP, logM[0], logM[1], logM[2], logM[3], log_acc);
  PMG_LAUNCH_CHECK();
  return PMG_OK;
}
