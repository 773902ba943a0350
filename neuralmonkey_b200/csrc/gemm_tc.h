// Host interface of the tcgen05 GEMM (gemm_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nm {

enum TcEpiMode {
  TC_EPI_DENSE = 0,     // C = act(acc + bias) + beta*C
  TC_EPI_XENT_FWD = 1,  // per-(row, n-tile) softmax partials; optional logits store
  TC_EPI_XENT_BWD = 2,  // C = (exp(x - lse) - onehot) * weights * scale
};

constexpr int TC_XENT_BN = 256;  // N tile used by the xent epilogues (sizes `part`)

struct TcEpilogue {
  int mode;
  float* C;  // dense: output; xent_fwd: logits (may be null); xent_bwd: dlogits
  int64_t ldc;
  const float* bias;  // [N] or null
  int act;
  float beta;
  // xent modes: x = acc + bias[n] + (n == unk_index ? -1e9 : 0)
  int64_t unk_index;       // < 0: none
  const int64_t* targets;  // [M] or null
  const float* weights;    // [M] or null (xent_bwd)
  const float* lse;        // [M] (xent_bwd)
  const float* scale;      // device scalar (xent_bwd)
  float4* part;            // [M][2*ceil(N/TC_XENT_BN)] (xent_fwd): (max, sumexp, argmax bits,
                           //   target logit or -inf) per (row, n-tile, epilogue half)
};

// True when the operands can be addressed by TMA (16-byte aligned rows and bases).
bool tc_gemm_supported(int transA, int transB, int64_t M, int64_t N, int64_t K, int64_t lda,
                       int64_t ldb, int64_t ldc, const void* A, const void* B, const void* C);

// op(A)[M,K] . op(B)[K,N] with the given epilogue.  Same operand conventions as nm_gemm.
int tc_gemm_launch(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A,
                   int64_t lda, const float* B, int64_t ldb, const TcEpilogue& epi, cudaStream_t s);

}  // namespace nm
