// Host interface of the tcgen05 GEMM (gemm_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nm {

enum TcEpiMode {
  TC_EPI_DENSE = 0,     // C = act(acc + bias) + beta*C
  TC_EPI_XENT_FWD = 1,  // per-(row, n-tile) softmax partials; optional logits store
  TC_EPI_XENT_BWD = 2,  // C = (exp(x - lse) - onehot) * weights * scale
  TC_EPI_XENT_BWD16 = 3,  // fp16 operands: C16 (and C16T) = half((exp(x - lse) - onehot) * weights)
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

// Extra arguments of the fp16-operand instances (kind::f16, K-major operands only); a separate
// trailing kernel parameter so the TF32 instances keep their parameter layout.
struct TcExt {
  void* C16;               // xent_bwd16: [M,N] fp16, row pitch ldc16 (elements)
  int64_t ldc16;
  void* C16T;              // xent_bwd16: the same matrix transposed, [N,M] fp16, row pitch ldc16t; may be null
  int64_t ldc16t;
  const float* alpha;      // dense: device scalar multiplying the accumulator (null: 1)
  const float* row_scale;  // dense: [M] per-row factor (null: 1)
  int transposed;          // dense: store D^T - element (m, n) goes to C[n * ldc + m]
};

// True when the operands can be addressed by TMA (16-byte aligned rows and bases).
bool tc_gemm_supported(int transA, int transB, int64_t M, int64_t N, int64_t K, int64_t lda,
                       int64_t ldb, int64_t ldc, const void* A, const void* B, const void* C);

// op(A)[M,K] . op(B)[K,N] with the given epilogue.  Same operand conventions as nm_gemm.
int tc_gemm_launch(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A,
                   int64_t lda, const float* B, int64_t ldb, const TcEpilogue& epi, cudaStream_t s);

// The same product with fp16 operands, both K-major: A is [M,K] (row pitch lda), B is [N,K] (row pitch
// ldb), pitches multiples of 8 elements, bases 16-byte aligned.  epi.mode: TC_EPI_DENSE (with `ext`),
// TC_EPI_XENT_FWD or TC_EPI_XENT_BWD16.
int tc_gemm16_launch(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
                     int64_t ldb, const TcEpilogue& epi, const TcExt& ext, cudaStream_t s);

// The same with both operands MN-major (reduction dimension strided): A is [K,M] (row pitch lda), B is
// [K,N] (row pitch ldb) - weight-gradient products X^T . dY without transposed copies.  Dense epilogue.
int tc_gemm16_mn_launch(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
                        int64_t ldb, const TcEpilogue& epi, const TcExt& ext, cudaStream_t s);

// CTA-pair tiles (cta_group::2, 256 x BN per pair): -1 = the library's policy, 0 = never, 1 = wherever the shape
// allows.  Returns the previous mode.  Overrides the NMB200_TC_PAIR environment variable.
int tc_gemm_set_pair_mode(int mode);

}  // namespace nm
