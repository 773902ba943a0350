// fp32 CUDA-core GEMM tile used (a) for shapes TMA cannot address (the toy
// dims of tests/bahdanau.ini: 7, 9, 11, 14 ...) and (b) inside the per-step
// recurrent kernels, where exact fp32 accumulation is kept on purpose
// (hard part (b)).  256 threads, BK = 16, register tile
// TM x TN per thread, operands addressed through (row, col) strides so all four
// transpose combinations share one body.
#pragma once
#include "common.cuh"

namespace nm {

constexpr int SIMT_BK = 16;
constexpr int SIMT_THREADS = 256;

template <int BM, int BN, int TM, int TN>
struct SimtSmem {
  float a[SIMT_BK][BM + 4];
  float b[SIMT_BK][BN + 4];
};

// acc[i][j] += sum_k A(m0 + ty*TM + i, k) * B(k, n0 + tx*TN + j)
// A(m,k) = A[m*sAm + k*sAk], B(k,n) = B[k*sBk + n*sBn]; out-of-range reads are 0.
template <int BM, int BN, int TM, int TN>
__device__ __forceinline__ void simt_tile(const float* __restrict__ A, int64_t sAm, int64_t sAk,
                                          const float* __restrict__ B, int64_t sBk, int64_t sBn,
                                          int64_t M, int64_t N, int64_t K, int64_t m0, int64_t n0,
                                          float (&acc)[TM][TN], SimtSmem<BM, BN, TM, TN>& sm) {
  static_assert((BM / TM) * (BN / TN) == SIMT_THREADS, "tile/thread mismatch");
  const int t = threadIdx.x;
  const int tx = t % (BN / TN), ty = t / (BN / TN);
  const bool a_kcontig = (sAk == 1);
  const bool b_ncontig = (sBn == 1);
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = 0; k0 < K; k0 += SIMT_BK) {
#pragma unroll
    for (int i = 0; i < (BM * SIMT_BK) / SIMT_THREADS; ++i) {
      const int idx = t + i * SIMT_THREADS;
      int m, k;
      if (a_kcontig) { m = idx / SIMT_BK; k = idx % SIMT_BK; }
      else           { k = idx / BM;      m = idx % BM; }
      const int64_t gm = m0 + m, gk = k0 + k;
      sm.a[k][m] = (gm < M && gk < K) ? A[gm * sAm + gk * sAk] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < (BN * SIMT_BK) / SIMT_THREADS; ++i) {
      const int idx = t + i * SIMT_THREADS;
      int n, k;
      if (b_ncontig) { k = idx / BN;      n = idx % BN; }
      else           { n = idx / SIMT_BK; k = idx % SIMT_BK; }
      const int64_t gn = n0 + n, gk = k0 + k;
      sm.b[k][n] = (gn < N && gk < K) ? B[gk * sBk + gn * sBn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SIMT_BK; ++k) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = sm.a[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = sm.b[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
}

// Generic kernel: epilogue functor gets (m, n, value) for every in-range element.
template <int BM, int BN, int TM, int TN, class Epi>
__global__ void __launch_bounds__(SIMT_THREADS)
simt_gemm_kernel(const float* __restrict__ A, int64_t sAm, int64_t sAk, const float* __restrict__ B,
                 int64_t sBk, int64_t sBn, int64_t M, int64_t N, int64_t K, Epi epi) {
  __shared__ SimtSmem<BM, BN, TM, TN> sm;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  float acc[TM][TN];
  simt_tile<BM, BN, TM, TN>(A, sAm, sAk, B, sBk, sBn, M, N, K, m0, n0, acc, sm);
  const int tx = threadIdx.x % (BN / TN), ty = threadIdx.x / (BN / TN);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + tx * TN + j;
      if (n < N) epi(m, n, acc[i][j]);
    }
  }
}

// Launch helper: picks 64x64 tiles when that still fills the chip, else 32x32.
template <class Epi>
static inline void simt_gemm_launch(const float* A, int64_t sAm, int64_t sAk, const float* B,
                                    int64_t sBk, int64_t sBn, int64_t M, int64_t N, int64_t K,
                                    Epi epi, cudaStream_t s) {
  const int64_t big = ceil_div(M, 64) * ceil_div(N, 64);
  if (big >= 2 * (int64_t)sm_count()) {
    dim3 grid((unsigned)ceil_div(N, 64), (unsigned)ceil_div(M, 64));
    simt_gemm_kernel<64, 64, 4, 4, Epi><<<grid, SIMT_THREADS, 0, s>>>(A, sAm, sAk, B, sBk, sBn, M, N,
                                                                     K, epi);
  } else {
    dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(M, 32));
    simt_gemm_kernel<32, 32, 2, 2, Epi><<<grid, SIMT_THREADS, 0, s>>>(A, sAm, sAk, B, sBk, sBn, M, N,
                                                                     K, epi);
  }
}

}  // namespace nm
