// tcgen05 GEMM for sm_100a: D[M,N] = op(A)[M,K] . op(B)[K,N], fp32 operands consumed as
// TF32 (kind::tf32), fp32 accumulation in TMEM.
//
//   * persistent: one CTA per SM walks 128 x BN output tiles (static round-robin);
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM allocator,
//     warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4);
//   * STAGES-deep smem ring of {A 128x32, B BNx32} fp32 tiles in the 128-byte swizzle
//     (16-byte atoms for K-major operands, 32-byte atoms for MN-major ones),
//     filled by cp.async.bulk.tensor (OOB rows/cols arrive as zeros: no tail code);
//   * two TMEM accumulator buffers (2*BN columns) so the epilogue of tile i overlaps
//     the main loop of tile i+1;
//   * both operand majors: "K-major" (reduction dim contiguous) and "MN-major"
//     (reduction dim strided), so forward (X.W), input-gradient (dY.W^T) and
//     weight-gradient (X^T.dY) products need no transposes in HBM;
//   * epilogues: dense (bias/activation/accumulate), and the fused vocabulary
//     cross-entropy forward (online softmax partials + argmax + target logit) and
//     backward (softmax - onehot), which never materialise fp32 logits twice.
//
// Descriptor bit layouts follow the PTX ISA "tcgen05 shared memory descriptor" and
// "instruction descriptor" tables (as restated in CUTLASS cute/arch/mma_sm100_desc.hpp).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "gemm_tc.h"
#include "tc_ptx.cuh"

namespace nm {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                      // fp32 elements = 128 bytes = one swizzle row
constexpr int TC_UMMA_K = 8;                   // tf32: 32 bytes per instruction
constexpr int TC_THREADS = 320;                // TMA warp, MMA warp, 8 epilogue warps
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;  // 16 KB
constexpr int TC_SMEM_BUDGET = 200 * 1024;

// CTAS = 2: a CTA pair works on one 256 x BN tile (cta_group::2); each CTA stages BN/2 rows of the B tile
template <int BN, int CTAS = 1>
struct TcCfg {
  static constexpr int B_BYTES = (BN / CTAS) * TC_BK * 4;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int STAGES = (TC_SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (TC_SMEM_BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = BN <= 64 ? 128 : (BN <= 128 ? 256 : 512);  // 2*BN rounded to a power of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ +
                                    TC_EPI_WARPS * 4096 /*per-warp store staging*/;
};

// ---------------------------------------------------------------------------
// epilogue for one 32-column chunk owned by one thread (= one output row).
// The epilogue runs once per output element, so it is written for instruction count:
// 32-bit column arithmetic, bias fetched as 8 x 16-byte uniform loads, exp as one FFMA +
// one MUFU.EX2, and the rare cases (the <unk> column, the row's target column, ragged
// right edge) handled in branches only the affected chunk takes.
// ---------------------------------------------------------------------------
struct RowStats {
  float mx, sum, tgt;
  int32_t arg;
};

constexpr float TC_LOG2E = 1.4426950408889634f;

// 2^x as ONE MUFU instruction.  exp2f() wraps the same MUFU.EX2 in a range fix for results below 2^-126
// (compare, halve, square: three more issue slots per element in epilogues that run once per logit); here such
// results flush to zero, which is what they contribute to a sum of probabilities anyway.
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void load_bias32(const float* __restrict__ bias, int col0, int ncols,
                                            float (&b)[32]) {
  if (bias == nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j) b[j] = 0.f;
    return;
  }
  if (ncols == 32 && ((reinterpret_cast<uintptr_t>(bias + col0) & 15) == 0)) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
      b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) b[j] = (j < ncols) ? __ldg(bias + col0 + j) : 0.f;
  }
}

__device__ __forceinline__ void store32(float* __restrict__ c, const float (&x)[32], int ncols,
                                        bool vec_ok) {
  if (vec_ok && ncols == 32) {
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<float4*>(c + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) c[j] = x[j];
  }
}

// Coalesced store of a warp's 32x32 chunk.  tcgen05.ld hands every thread one ROW (32 columns),
// so direct stores scatter 16-byte pieces over 32 different cache lines per instruction - the L2
// request rate, not DRAM, then bounds a kernel that writes a large C (the 1.6 GB dlogits).  Going
// through a 4 KB per-warp staging tile (XOR-swizzled float4 slots: conflict-free both ways) turns
// each store instruction into four full 128-byte row segments.
__device__ __forceinline__ void store32_coalesced(float* __restrict__ stage, float* __restrict__ C,
                                                  int64_t ldc, int64_t row_base, int col0, int64_t M,
                                                  const float (&x)[32], int lane) {
  float4* st4 = reinterpret_cast<float4*>(stage);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    st4[lane * 8 + (j ^ (lane & 7))] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
  __syncwarp();
  const int sub = lane >> 3, slot = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + sub;
    const float4 v = st4[r * 8 + (slot ^ (r & 7))];
    const int64_t grow = row_base + r;
    if (grow < M) *reinterpret_cast<float4*>(C + grow * ldc + col0 + 4 * slot) = v;
  }
  __syncwarp();
}

// mode is a compile-time constant so each kernel instance carries one epilogue only.
template <int MODE>
__device__ __forceinline__ void epilogue_chunk(const TcEpilogue& e, float (&x)[32], int64_t row,
                                               int col0, int64_t M, int N, RowStats& st,
                                               int target, float row_lse2, float row_w,
                                               bool vec_ok, float* __restrict__ stage, int lane) {
  if (col0 >= N) return;                                     // warp-uniform
  const int ncols = min(32, N - col0);
  // full, aligned chunks leave through the staging tile (all 32 lanes take part, rows >= M are
  // masked at the store); everything else keeps the per-row path
  const bool coalesced = vec_ok && ncols == 32 &&
                         (MODE == TC_EPI_XENT_BWD || (MODE == TC_EPI_DENSE && e.beta == 0.f));
  if (row >= M && !coalesced) return;
  float b[32];
  load_bias32(e.bias, col0, ncols, b);
  if (MODE == TC_EPI_DENSE) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = apply_act(x[j] + b[j], e.act);
    if (coalesced) {
      store32_coalesced(stage, e.C, e.ldc, row - lane, col0, M, x, lane);
      return;
    }
    float* c = e.C + row * e.ldc + col0;
    if (e.beta != 0.f) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) x[j] += c[j];
    }
    store32(c, x, ncols, vec_ok);
    return;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] += b[j];
  const int unk_rel = (int)e.unk_index - col0;  // rare: only the chunk holding <unk>
  if (unk_rel >= 0 && unk_rel < 32) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j == unk_rel) x[j] += -1e9f;
  }
  if (MODE == TC_EPI_XENT_FWD) {
    if (ncols < 32) {  // ragged right edge: padding columns must not win the max
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j >= ncols) x[j] = -INFINITY;
    }
    float cmx = x[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) cmx = fmaxf(cmx, x[j]);
    if (cmx > st.mx) {  // strict: earlier chunks (lower columns) keep ties
      int carg = 0;
#pragma unroll
      for (int j = 31; j >= 0; --j)
        if (x[j] == cmx) carg = j;  // lowest index among equals
      st.sum *= fast_ex2((st.mx - cmx) * TC_LOG2E);
      st.mx = cmx;
      st.arg = col0 + carg;
    }
    const float m2 = st.mx * TC_LOG2E;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // four chains of 8 dependent adds instead of one of 32
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      s0 += fast_ex2(fmaf(x[j], TC_LOG2E, -m2));
      s1 += fast_ex2(fmaf(x[j + 1], TC_LOG2E, -m2));
      s2 += fast_ex2(fmaf(x[j + 2], TC_LOG2E, -m2));
      s3 += fast_ex2(fmaf(x[j + 3], TC_LOG2E, -m2));
    }
    st.sum += (s0 + s1) + (s2 + s3);
    const int t_rel = target - col0;  // rare: the chunk holding this row's target
    if (t_rel >= 0 && t_rel < ncols) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j == t_rel) st.tgt = x[j];
    }
    if (e.C) store32(e.C + row * e.ldc + col0, x, ncols, vec_ok);
  } else {  // TC_EPI_XENT_BWD: (softmax - onehot) * row weight
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = fast_ex2(fmaf(x[j], TC_LOG2E, -row_lse2)) * row_w;
    const int t_rel = target - col0;
    if (t_rel >= 0 && t_rel < ncols) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j == t_rel) x[j] -= row_w;
    }
    if (coalesced)
      store32_coalesced(stage, e.C, e.ldc, row - lane, col0, M, x, lane);
    else
      store32(e.C + row * e.ldc + col0, x, ncols, vec_ok);
  }
}

// split-K partial tile: C += x (+ bias once, from split 0); no activation.
__device__ __forceinline__ void red_add_v4(float* addr, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__device__ __forceinline__ void epilogue_chunk_atomic(const TcEpilogue& e, const float (&x)[32],
                                                      int64_t row, int col0, int64_t M, int N,
                                                      bool add_bias, bool vec_ok,
                                                      float* __restrict__ stage, int lane) {
  if (col0 >= N) return;                                   // warp-uniform
  const int ncols = min(32, N - col0);
  if (vec_ok && ncols == 32) {
    // same transposition as store32_coalesced: one vector reduction covers 16 bytes of a row
    // and a warp instruction four full 128-byte row segments, instead of 32 scalar atomics
    // scattered over 32 rows
    float4* st4 = reinterpret_cast<float4*>(stage);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 v = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
      if (add_bias && e.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + col0 + 4 * j));
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      }
      st4[lane * 8 + (j ^ (lane & 7))] = v;
    }
    __syncwarp();
    const int sub = lane >> 3, slot = lane & 7;
    const int64_t row_base = row - lane;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + sub;
      const float4 v = st4[r * 8 + (slot ^ (r & 7))];
      const int64_t grow = row_base + r;
      if (grow < M) red_add_v4(e.C + grow * e.ldc + col0 + 4 * slot, v);
    }
    __syncwarp();
    return;
  }
  if (row >= M) return;
  float* c = e.C + row * e.ldc + col0;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (j < ncols) {
      float v = x[j];
      if (add_bias && e.bias) v += __ldg(e.bias + col0 + j);
      atomicAdd(c + j, v);
    }
}

// ---------------------------------------------------------------------------
// epilogues of the fp16-operand instances (ESZ == 2)
// ---------------------------------------------------------------------------
// Transposed store of a warp's 32x32 chunk: lane = row, so for one column the 32 lanes write 32
// consecutive elements of the transposed matrix - a full segment without any staging.
template <typename T>
__device__ __forceinline__ void store32_transposed(T* __restrict__ ct, int64_t ldt, int64_t row, int col0,
                                                   int64_t M, int ncols, const float (&x)[32], float beta) {
  if (row >= M) return;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < ncols) {
      T* dst = ct + (int64_t)(col0 + j) * ldt + row;
      if constexpr (sizeof(T) == 2) {
        *dst = __float2half_rn(x[j]);
      } else {
        *dst = (beta != 0.f) ? (x[j] + *dst) : x[j];
      }
    }
  }
}

// Row-major fp16 store of a warp's 32x32 chunk through the per-warp staging tile (32 rows x 64 B,
// 16-byte slots XOR-swizzled): every store instruction then writes eight full 64-byte row segments.
__device__ __forceinline__ void store32_half_coalesced(float* __restrict__ stage, __half* __restrict__ C,
                                                       int64_t ldc, int64_t row_base, int col0, int64_t M,
                                                       const float (&x)[32], int lane) {
  uint4* st4 = reinterpret_cast<uint4*>(stage);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __half2 h0 = __floats2half2_rn(x[8 * j], x[8 * j + 1]), h1 = __floats2half2_rn(x[8 * j + 2], x[8 * j + 3]);
    __half2 h2 = __floats2half2_rn(x[8 * j + 4], x[8 * j + 5]), h3 = __floats2half2_rn(x[8 * j + 6], x[8 * j + 7]);
    uint4 v;
    v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
    v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
    st4[lane * 4 + (j ^ ((lane >> 1) & 3))] = v;     // rows 2q, 2q+1 share a swizzle: conflict-free both ways
  }
  __syncwarp();
  const int sub = lane >> 2, slot = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + sub;
    const uint4 v = st4[r * 4 + (slot ^ ((r >> 1) & 3))];
    const int64_t grow = row_base + r;
    if (grow < M) *reinterpret_cast<uint4*>(C + grow * ldc + col0 + 8 * slot) = v;
  }
  __syncwarp();
}

// TC_EPI_XENT_BWD16: (softmax - onehot) * row weight, stored as fp16 row-major and transposed.
// The row weight is the 0/1 mask here (the caller applies the upstream scale in the consumers), so
// the stored values lie in [-1, 1] and fp16 keeps TF32's 10 mantissa bits for them.
__device__ __forceinline__ void epilogue_chunk_xent_bwd16(const TcEpilogue& e, const TcExt& ext, float (&x)[32],
                                                          int64_t row, int col0, int64_t M, int N, int target,
                                                          float row_lse2, float row_w,
                                                          float* __restrict__ stage, int lane) {
  if (col0 >= N) return;                                     // warp-uniform
  const int ncols = min(32, N - col0);
  float b[32];
  load_bias32(e.bias, col0, ncols, b);
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] += b[j];
  const int unk_rel = (int)e.unk_index - col0;
  if (unk_rel >= 0 && unk_rel < 32) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j == unk_rel) x[j] += -1e9f;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = fast_ex2(fmaf(x[j], TC_LOG2E, -row_lse2)) * row_w;
  const int t_rel = target - col0;
  if (t_rel >= 0 && t_rel < ncols) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j == t_rel) x[j] -= row_w;
  }
  __half* c16 = reinterpret_cast<__half*>(ext.C16);
  const bool vec_ok = ncols == 32 && ((ext.ldc16 & 7) == 0) && ((reinterpret_cast<uintptr_t>(c16) & 15) == 0);
  if (vec_ok) {
    store32_half_coalesced(stage, c16, ext.ldc16, row - lane, col0, M, x, lane);
  } else if (row < M) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) c16[row * ext.ldc16 + col0 + j] = __float2half_rn(x[j]);
  }
  if (ext.C16T)
    store32_transposed(reinterpret_cast<__half*>(ext.C16T), ext.ldc16t, row, col0, M, ncols, x, 0.f);
}

// TC_EPI_DENSE of the fp16 instances: C (or C^T) = alpha * row_scale[m] * acc + bias (+ C).
__device__ __forceinline__ void epilogue_chunk_dense16(const TcEpilogue& e, const TcExt& ext, float (&x)[32],
                                                       int64_t row, int col0, int64_t M, int N, float factor,
                                                       bool vec_ok, float* __restrict__ stage, int lane,
                                                       RowStats& st) {
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] *= factor;
  if (!ext.transposed) {
    epilogue_chunk<TC_EPI_DENSE>(e, x, row, col0, M, N, st, -1, 0.f, 0.f, vec_ok, stage, lane);
    return;
  }
  if (col0 >= N) return;
  const int ncols = min(32, N - col0);
  if (e.bias) {
    float b[32];
    load_bias32(e.bias, col0, ncols, b);
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] += b[j];
  }
  store32_transposed(e.C, e.ldc, row, col0, M, ncols, x, e.beta);
}

// ---------------------------------------------------------------------------
// TC_EPI_SOFTMAX / TC_EPI_DSOFTMAX: one thread owns one query row of one (sentence, head) problem; the whole
// row of energies sits in this thread's TMEM lane (N <= BN), so the row reductions need no exchange - the lane
// is simply read again for every pass (tensor memory is next to the SM).  Mask semantics of the reference
// (scaled_dot_product.py:160-206): causal positions are REPLACED by -1e9, padded keys get x*m + (1-m)*(-1e9),
// both before the softmax; dropout multiplies the softmax output (:208-214).
//
// Everything a row needs from memory - its slice of the dropout mask, of the saved softmax - and everything it
// writes goes through the warp's staging tile, 32 rows x 32 columns at a time (the inverse of
// store32_coalesced), so global memory sees full 128-byte row segments; the key mask of the sentence is copied
// once per tile into the idle partner warp's staging tile.  (The first version read and wrote per-thread rows
// element by element: 99 us per launch at the bench shape, six times the products themselves.)
__device__ __forceinline__ void load32_coalesced(float* __restrict__ stage, const float* __restrict__ src,
                                                 int64_t ld, int64_t row_base, int col0, int64_t M,
                                                 float (&x)[32], int lane) {
  float4* st4 = reinterpret_cast<float4*>(stage);
  const int sub = lane >> 3, slot = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + sub;
    const int64_t grow = row_base + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grow < M) v = *reinterpret_cast<const float4*>(src + grow * ld + col0 + 4 * slot);
    st4[r * 8 + (slot ^ (r & 7))] = v;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 v = st4[lane * 8 + (j ^ (lane & 7))];
    x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
  }
  __syncwarp();
}

template <int MODE>
__device__ __forceinline__ void attn_epilogue(const TcEpilogue& e, const TcBatch& bt, uint32_t t_row, int r,
                                              int M, int N, int o, int p, int64_t c_off,
                                              float* __restrict__ stage, int lane) {
  const int nchunks = bt.n_pad >> 5;                 // n_pad is a multiple of 32
  float* km_s = stage + 4 * 1024;                    // the partner warp's tile (idle in these modes)
  const bool has_km = bt.key_mask != nullptr;
  if (has_km) {
    const float* km = bt.key_mask + (int64_t)o * N;
    for (int c = lane; c < nchunks * 32; c += 32) km_s[c] = c < N ? km[c] : 1.f;
    __syncwarp();
  }
  const bool row_ok = r < M;
  const int64_t row_base = r - lane;
  float* cbase = e.C + c_off;
  const float* dbase = bt.drop ? bt.drop + (int64_t)p * M * N : nullptr;
  const bool drop_vec = dbase && (N & 3) == 0 && (reinterpret_cast<uintptr_t>(dbase) & 15) == 0;
  // this row's 32 entries of the dropout mask from column col0 (1 where there is no mask, 0 outside the matrix)
  auto load_drop = [&](int col0, float (&d)[32]) {
    if (!dbase) {
#pragma unroll
      for (int j = 0; j < 32; ++j) d[j] = 1.f;
    } else if (drop_vec && col0 + 32 <= N) {
      load32_coalesced(stage, dbase, N, row_base, col0, M, d, lane);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) d[j] = (row_ok && col0 + j < N) ? dbase[(int64_t)r * N + col0 + j] : 0.f;
    }
  };
  // the rows this tile will read from memory (dropout mask, saved softmax): on their way into L2 while the
  // tensor-memory passes run
  if (row_ok) {
    if (dbase) asm volatile("prefetch.global.L2 [%0];" ::"l"(dbase + (int64_t)r * N) : "memory");
    if (MODE == TC_EPI_DSOFTMAX) asm volatile("prefetch.global.L2 [%0];" ::"l"(bt.P + c_off + (int64_t)r * e.ldc) : "memory");
  }
  float v[32];
  if (MODE == TC_EPI_SOFTMAX) {
    // energies in units of log2: 2^(x*log2e - max) is ONE MUFU instruction (fast_ex2), the masks' -1e9 stays -1e9*log2e
    const float sc2 = bt.scale * TC_LOG2E;
    constexpr float MASKED2 = -1e9f * TC_LOG2E;
    auto energy = [&](float acc, int col) {
      float x = acc * sc2;
      if (bt.causal && col > r) x = MASKED2;
      if (has_km) {
        const float m = km_s[col];
        x = x * m + (1.f - m) * MASKED2;
      }
      return x;
    };
    float mx = -INFINITY;
    for (int c = 0; c < nchunks; ++c) {
      tmem_ld32(t_row + (uint32_t)(c * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < N) mx = fmaxf(mx, energy(v[j], c * 32 + j));
    }
    float sum = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      tmem_ld32(t_row + (uint32_t)(c * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < N) sum += fast_ex2(energy(v[j], c * 32 + j) - mx);
    }
    const float inv = 1.f / sum;
    for (int c = 0; c < nchunks; ++c) {
      tmem_ld32(t_row + (uint32_t)(c * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = c * 32 + j;
        v[j] = (row_ok && col < N) ? fast_ex2(energy(v[j], col) - mx) * inv : 0.f;
      }
      store32_coalesced(stage, cbase, e.ldc, row_base, c * 32, bt.m_pad, v, lane);   // padding rows: zeros
      if (bt.C2) {
        float d[32];
        load_drop(c * 32, d);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= d[j];
        store32_coalesced(stage, bt.C2 + c_off, e.ldc, row_base, c * 32, bt.m_pad, v, lane);
      }
    }
  } else {   // TC_EPI_DSOFTMAX: acc = d(dropped weights)
    const float* pbase = bt.P + c_off;
    float pr[32], d[32];
    float dot = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      tmem_ld32(t_row + (uint32_t)(c * 32), v);
      load32_coalesced(stage, pbase, e.ldc, row_base, c * 32, M, pr, lane);   // zero beyond the rows / columns
      load_drop(c * 32, d);
#pragma unroll
      for (int j = 0; j < 32; ++j) dot = fmaf(v[j] * d[j], pr[j], dot);
    }
    for (int c = 0; c < nchunks; ++c) {
      tmem_ld32(t_row + (uint32_t)(c * 32), v);
      load32_coalesced(stage, pbase, e.ldc, row_base, c * 32, M, pr, lane);
      load_drop(c * 32, d);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = c * 32 + j;
        float g = pr[j] * (v[j] * d[j] - dot);             // pr = 0 in the padding: g = 0 there
        if (has_km) g *= km_s[col];                         // d(x*m + c)/dx = m
        if (bt.causal && col > r) g = 0.f;                  // tf.where: no gradient into replaced entries
        v[j] = row_ok ? g * bt.scale : 0.f;
      }
      store32_coalesced(stage, cbase, e.ldc, row_base, c * 32, bt.m_pad, v, lane);
    }
  }
}

// smem descriptor fields of an MN-major operand tile (bytes); a kernel argument so a
// diagnostic run can probe them (NMB200_MN_* environment variables), fixed otherwise.
struct MnDesc {
  uint32_t layout, sbo, lbo, kadv;
};

// ESZ = operand element size: 4 = fp32 consumed as TF32 (kind::tf32), 2 = fp16 (kind::f16, K-major
// operands only).  A k-block is 128 bytes of K either way (32 or 64 elements) and one instruction
// consumes 32 of them, so the smem ring, the descriptors and the TMEM layout are the same.
template <int BN, bool A_MN, bool B_MN, int MODE, int ESZ = 4, int CTAS = 1>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               int64_t M, int64_t N, int64_t K, TcEpilogue epi, MnDesc mn, int splits, int kb_per_split,
               TcExt ext, TcBatch bt) {
  static_assert(ESZ == 4 || !(A_MN || B_MN) || (BN % 64 == 0), "MN-major fp16 B tiles come in 64-column boxes");
  static_assert(ESZ == 2 || MODE != TC_EPI_XENT_BWD16, "the fp16 epilogue belongs to the fp16 instances");
  static_assert(CTAS == 1 || CTAS == 2, "one CTA or a CTA pair per tile");
  static_assert(CTAS == 1 || (BN % 64 == 0), "a pair splits the B tile in two halves of whole swizzle atoms");
  constexpr int BK = 128 / ESZ;                // elements per 128-byte k-block
  constexpr int BM_T = TC_BM * CTAS;           // rows of one tile (pair: 256, 128 per CTA)
  constexpr int BN_LOAD = BN / CTAS;           // rows of the B tile this CTA stages
  using Cfg = TcCfg<BN, CTAS>;
  // pair mode: rank 0 leads (issues the MMAs, owns the `full` and `tmem_empty` barriers the pair reports to)
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int64_t worker = CTAS == 2 ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;
  const int64_t nworkers = CTAS == 2 ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SW128 tiles: 1 KB aligned
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  // barrier layout (8 B each): full[STAGES] | empty[STAGES] | tmem_full[2] | tmem_empty[2] | tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tiles_m = (M + BM_T - 1) / BM_T;
  const int64_t tiles_n = (N + BN - 1) / BN;
  // split-K: the reduction is cut into `splits` slices, each an independent work item whose
  // epilogue adds its partial tile into C with red.global.add (weight-gradient products
  // have tiny outputs and very long K: without this only a handful of SMs would work)
  // batched: bt.count independent problems of this M x N x K, each a window of the operand tensors (no split-K)
  const bool batched = bt.count > 0;
  const int64_t tiles_per = tiles_m * tiles_n;
  const int64_t num_tiles = batched ? tiles_per * bt.count : tiles_per * splits;
  const int num_kb_total = (int)((K + BK - 1) / BK);  // host guarantees no empty split

  if (threadIdx.x == 32) {   // descriptor fetch off the critical path of the first TMA load
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), TC_EPI_WARPS * CTAS);  // one arrive per epilogue warp (of both CTAs)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CTAS == 2) {   // the same warp of both CTAs allocates the pair's columns
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                   "r"((uint32_t)Cfg::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                   "r"((uint32_t)Cfg::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (CTAS == 2) cluster_sync_all();   // the peer's barriers exist before anything arrives on them
  tcgen05_fence_after();
  // Programmatic dependent launch (NMB200_TC_PDL): everything above - barriers, the TMEM allocation - touches no
  // global memory and may run while the previous kernel of the stream drains; from here on its results are
  // needed.  Without the launch attribute both instructions do nothing.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = worker; tile < num_tiles; tile += nworkers) {
        const int32_t m0 = (int32_t)((tile % tiles_m) * BM_T) + (int32_t)cta_rank * TC_BM;
        const int32_t n0 = (int32_t)(((tile / tiles_m) % tiles_n) * BN) + (int32_t)cta_rank * BN_LOAD;
        int kb_begin = (int)(tile / tiles_per) * kb_per_split;
        int kb_end = min(num_kb_total, kb_begin + kb_per_split);
        int32_t a_row = 0, a_col = 0, b_row = 0, b_col = 0;   // this problem's operand windows
        if (batched) {
          const int prob = (int)(tile / tiles_per), o = prob / bt.inner, i = prob - o * bt.inner;
          a_row = o * bt.a_row_outer + i * bt.a_row_inner;
          a_col = i * bt.a_col_inner;
          b_row = o * bt.b_row_outer + i * bt.b_row_inner;
          b_col = i * bt.b_col_inner;
          kb_begin = 0;
          kb_end = num_kb_total;
        }
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t b_dst = a_dst + TC_A_BYTES;
          // pair: the leader's barrier counts the bytes of both CTAs
          if (leader) mbar_expect_tx(full_bar(stage), (uint32_t)(Cfg::STAGE_BYTES * CTAS));
          const uint32_t full_dst = CTAS == 2 ? mapa_u32(full_bar(stage), 0u) : full_bar(stage);
          auto tma_load = [&](uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1) {
            if constexpr (CTAS == 2) tma_load_2d_pair(dst, map, full_dst, c0, c1);
            else tma_load_2d(dst, map, full_dst, c0, c1);
          };
          const int32_t k0 = kb * BK;
          // MN-major tiles arrive as boxes of one 128-byte swizzle row of MN elements (32 fp32 / 64 fp16)
          // by one k-block of rows: 4 KB (fp32) or 8 KB (fp16) each
          constexpr int MN_BOX = 128 / ESZ;
          constexpr int MN_BOX_BYTES = MN_BOX * BK * ESZ;
          if (!A_MN) {
            tma_load(a_dst, &map_a, k0 + a_col, m0 + a_row);  // box {one k-block, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < TC_BM / MN_BOX; ++j)
              tma_load(a_dst + j * MN_BOX_BYTES, &map_a, m0 + a_col + MN_BOX * j, k0 + a_row);
          }
          if (!B_MN) {
            tma_load(b_dst, &map_b, k0 + b_col, n0 + b_row);  // box {one k-block, BN (pair: BN/2) rows}
          } else {
#pragma unroll
            for (int j = 0; j < BN_LOAD / MN_BOX; ++j)
              tma_load(b_dst + j * MN_BOX_BYTES, &map_b, n0 + b_col + MN_BOX * j, k0 + b_row);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && leader) {
      // instruction descriptor: c=F32 [4,6)=1, a=TF32 [7,10)=2, b=TF32 [10,13)=2,
      // a_major bit15, b_major bit16, N>>3 [17,23), M>>4 [24,29)
      constexpr uint32_t FMT = ESZ == 4 ? 2u : 0u;   // TF32 = 2, F16 = 0
      const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((A_MN ? 1u : 0u) << 15) |
                             ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM_T >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int64_t it = 0;
      for (int64_t tile = worker; tile < num_tiles; tile += nworkers, ++it) {
        const int acc = (int)(it & 1);
        const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        const int kb_begin = (int)(tile / tiles_per) * kb_per_split;
        const int num_kb = batched ? num_kb_total : min(num_kb_total, kb_begin + kb_per_split) - kb_begin;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t b_addr = a_addr + TC_A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
            // K-major (SWIZZLE_128B): atoms of 8 rows x 128 B, next 8-row group at +1024 B
            //   (SBO), K step of one instruction (8 tf32) = +32 B inside the swizzled row.
            // MN-major: tf32 operands whose reduction dim is strided must use the
            //   128B-swizzle-with-32B-atoms layout (the only MN-major tf32 layout tcgen05
            //   accepts): atoms of 4 k-rows x 128 B (32 MN elements); next 4-k-row group at
            //   +512 B (SBO), next 32-element MN atom = next TMA box at +4096 B (LBO);
            //   K step (8 k-rows) = +1024 B.
            // MN-major fp16 (kind::f16 takes both majors in the plain 128-byte swizzle): atoms of
            //   8 k-rows x 128 B (64 MN elements); next 8-k-row group at +1024 B (SBO), next 64-element
            //   MN atom = next TMA box at +8192 B (LBO); K step of one instruction (16 k-rows) = +2048 B.
            const uint64_t da = A_MN ? smem_desc(a_addr + k * mn.kadv, mn.lbo, mn.sbo, mn.layout)
                                     : smem_desc(a_addr + k * 32, 16, 1024, SMEM_LAYOUT_SW128);
            const uint64_t db = B_MN ? smem_desc(b_addr + k * mn.kadv, mn.lbo, mn.sbo, mn.layout)
                                     : smem_desc(b_addr + k * 32, 16, 1024, SMEM_LAYOUT_SW128);
            const uint32_t accum = (kb > 0 || k > 0) ? 1u : 0u;
            if constexpr (CTAS == 2) {
              if (ESZ == 4) umma_tf32_pair(d_tmem, da, db, idesc, accum);
              else umma_f16_pair(d_tmem, da, db, idesc, accum);
            } else {
              if (ESZ == 4) umma_tf32(d_tmem, da, db, idesc, accum);
              else umma_f16(d_tmem, da, db, idesc, accum);
            }
          }
          // frees this smem stage (of both CTAs) when the MMAs retire
          if constexpr (CTAS == 2) umma_commit_pair(empty_bar(stage), 3);
          else umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (CTAS == 2) umma_commit_pair(tfull_bar(acc), 3);
        else umma_commit(tfull_bar(acc));
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    // Two warps per TMEM lane quadrant (quadrant = warp_idx % 4 is a hardware rule); the
    // pair splits the tile's 32-column chunks (even / odd), which also gives every SM
    // sub-partition two epilogue warps to overlap their latencies.
    const int quad = warp & 3;                // TMEM lanes [32*quad, 32*quad+32)
    const int half = (warp - 2) >> 2;         // 0: even chunks, 1: odd chunks
    const bool vec_ok = ((epi.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(epi.C) & 15) == 0);
    // per-warp staging tile behind the barriers (generic pointer of the 1 KB-aligned window)
    float* stage = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) +
                                            STAGES * Cfg::STAGE_BYTES + 256 + (warp - 2) * 4096);
    const int n32 = (int)N;
    int64_t it = 0;
    for (int64_t tile = worker; tile < num_tiles; tile += nworkers, ++it) {
      const int acc = (int)(it & 1);
      const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
      const int64_t tm = tile % tiles_m, tn = (tile / tiles_m) % tiles_n;
      const int split = batched ? 0 : (int)(tile / tiles_per);
      const int64_t row = tm * BM_T + (int64_t)cta_rank * TC_BM + quad * 32 + lane;
      // batched: this problem's output window
      int prob = 0, prob_o = 0;
      int64_t c_off = 0;
      if (batched) {
        prob = (int)(tile / tiles_per);
        prob_o = prob / bt.inner;
        c_off = (int64_t)prob_o * bt.c_outer + (int64_t)(prob - prob_o * bt.inner) * bt.c_inner;
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      RowStats st{-INFINITY, 0.f, -INFINITY, 0x7fffffff};
      int target = -1;
      float row_lse2 = 0.f, row_w = 0.f;
      if (MODE != TC_EPI_DENSE && row < M) {
        if (epi.targets) target = (int)epi.targets[row];
        if (MODE == TC_EPI_XENT_BWD) {
          row_lse2 = epi.lse[row] * TC_LOG2E;
          row_w = (epi.weights ? epi.weights[row] : 1.f) * epi.scale[0];
        }
        if (MODE == TC_EPI_XENT_BWD16) {
          row_lse2 = epi.lse[row] * TC_LOG2E;
          row_w = epi.weights ? epi.weights[row] : 1.f;
        }
      }
      float factor16 = 1.f;
      if (ESZ == 2 && MODE == TC_EPI_DENSE) {
        if (ext.alpha) factor16 = ext.alpha[0];
        if (ext.row_scale && row < M) factor16 *= ext.row_scale[row];
      }
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      if constexpr (MODE == TC_EPI_SOFTMAX || MODE == TC_EPI_DSOFTMAX) {
        // whole rows per thread: the first warp of each lane quadrant does the tile (a handful of columns)
        if (half == 0) attn_epilogue<MODE>(epi, bt, t_row, (int)row, (int)M, n32, prob_o, prob, c_off, stage, lane);
      } else {
      TcEpilogue epi_p = epi;               // this problem's window of C
      if (batched) epi_p.C += c_off;
      const TcEpilogue& epi = epi_p;
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        float v[32];
        tmem_ld32(t_row + (uint32_t)(c * 32), v);
        if constexpr (MODE == TC_EPI_XENT_BWD16) {
          epilogue_chunk_xent_bwd16(epi, ext, v, row, (int)(tn * BN) + c * 32, M, n32, target, row_lse2,
                                    row_w, stage, lane);
        } else if constexpr (ESZ == 2 && MODE == TC_EPI_DENSE) {
          epilogue_chunk_dense16(epi, ext, v, row, (int)(tn * BN) + c * 32, M, n32, factor16, vec_ok,
                                 stage, lane, st);
        } else {
        if (MODE == TC_EPI_DENSE && splits > 1)
          epilogue_chunk_atomic(epi, v, row, (int)(tn * BN) + c * 32, M, n32, split == 0,
                                vec_ok && ((reinterpret_cast<uintptr_t>(epi.bias) & 15) == 0), stage, lane);
        else
          epilogue_chunk<MODE>(epi, v, row, (int)(tn * BN) + c * 32, M, n32, st, target, row_lse2,
                               row_w, vec_ok, stage, lane);
        }
      }
      }
      if (MODE == TC_EPI_XENT_FWD && row < M)
        epi.part[(row * tiles_n + tn) * 2 + half] =
            make_float4(st.mx, st.sum, __int_as_float(st.arg), st.tgt);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTAS == 2) mbar_arrive_cluster(mapa_u32(tempty_bar(acc), 0u));   // the leader's barrier
        else mbar_arrive(tempty_bar(acc));
      }
    }
  }

  // teardown: everyone done with TMEM before dealloc
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (CTAS == 2) cluster_sync_all();   // the peer is done with this CTA's barriers and with the pair's TMEM
  if (warp == 1) {
    tcgen05_fence_after();
    if constexpr (CTAS == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"((uint32_t)Cfg::TMEM_COLS)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"((uint32_t)Cfg::TMEM_COLS)
                   : "memory");
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor [rows, cols] with row pitch ld (elements); box = {box_cols, box_rows}.
static uint32_t env_u32(const char* name, uint32_t dflt) {
  const char* e = getenv(name);
  return (e && *e) ? (uint32_t)strtoul(e, nullptr, 0) : dflt;
}

static MnDesc mn_desc_config() {
  static MnDesc d = {0, 0, 0, 0};
  static bool init = false;
  if (!init) {
    d.layout = env_u32("NMB200_MN_LAYOUT", SMEM_LAYOUT_SW128_32B);
    d.sbo = env_u32("NMB200_MN_SBO", 512);
    d.lbo = env_u32("NMB200_MN_LBO", 4096);
    d.kadv = env_u32("NMB200_MN_KADV", 1024);
    init = true;
  }
  return d;
}

static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld,
                    uint32_t box_cols, uint32_t box_rows, bool mn_major) {
  static int mn_swizzle = -1;  // CUtensorMapSwizzle for MN-major operand tiles
  if (mn_swizzle < 0) mn_swizzle = (int)env_u32("NMB200_MN_SWIZZLE", (uint32_t)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  EncodeTiledFn fn = get_encode_fn();
  NM_REQUIRE(fn != nullptr, NM_E_NO_DEVICE, "tc_gemm: cuTensorMapEncodeTiled not available");
  static int dtype_mode = -1;  // NMB200_TMA_DTYPE=fp32 keeps raw fp32 bits (MMA truncates)
  if (dtype_mode < 0) {
    const char* e = getenv("NMB200_TMA_DTYPE");
    dtype_mode = (e && strcmp(e, "fp32") == 0) ? 1 : 0;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, dtype_mode ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32,
                        2, const_cast<float*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        mn_major ? (CUtensorMapSwizzle)mn_swizzle : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NM_REQUIRE(r == CUDA_SUCCESS, NM_E_INVALID,
             "tc_gemm: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
             (long long)rows, (long long)cols, (long long)ld);
  return NM_OK;
}

bool tc_gemm_supported(int transA, int transB, int64_t M, int64_t N, int64_t K, int64_t lda,
                       int64_t ldb, int64_t ldc, const void* A, const void* B, const void* C) {
  (void)transA; (void)transB; (void)ldc; (void)C;
  if (M < 1 || N < 1 || K < 1) return false;
  if ((lda & 3) || (ldb & 3)) return false;  // TMA global strides are multiples of 16 bytes
  if (A && (reinterpret_cast<uintptr_t>(A) & 15)) return false;
  if (B && (reinterpret_cast<uintptr_t>(B) & 15)) return false;
  if (M > 0x7fffffffLL || N > 0x7fffffffLL || K > 0x7fffffffLL) return false;
  return true;
}

constexpr bool TC_PAIR_DEFAULT = false;   // until the pair instances are verified on the GPU: opt-in only
// NMB200_TC_PAIR: 0 = never, 1 = wherever the shape allows, unset = the policy of pair_wanted()
static int g_pair_override = -2;   // nm_gemm_set_pair_mode(); -2 = not set
static int pair_mode() {
  static const int mode = [] {
    const char* e = getenv("NMB200_TC_PAIR");
    return (e && *e) ? atoi(e) : -1;
  }();
  return g_pair_override != -2 ? g_pair_override : mode;
}
int tc_gemm_set_pair_mode(int mode) {
  const int before = pair_mode();
  g_pair_override = mode;
  return before;
}

// A CTA pair per 256 x bn tile (cta_group::2): worth it when the product is bound by the bytes the SMs pull
// from L2 (large M and N), not for the skinny / split-K shapes whose tiles would no longer fill the chip.
static bool pair_wanted(int64_t M, int64_t N, int64_t K, int bn, int splits) {
  (void)K;
  if (bn != 128 && bn != 256) return false;
  if (sm_count() < 2 || M < 2 * TC_BM) return false;
  const int mode = pair_mode();
  if (mode == 0) return false;
  if (mode == 1) return true;
  if (!TC_PAIR_DEFAULT) return false;
  const int64_t pair_tiles = ceil_div(M, 2 * TC_BM) * ceil_div(N, bn) * splits;
  return splits == 1 && pair_tiles >= (int64_t)(sm_count() / 2);
}

template <class Kern, class... Args>
static int launch_pair_kernel(Kern kern, int smem_bytes, int64_t pair_tiles, cudaStream_t s, Args... args) {
  const int64_t max_pairs = sm_count() / 2;
  const int64_t pairs = pair_tiles < max_pairs ? pair_tiles : max_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NM_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, args...));
  NM_LAUNCH_CHECK("tc_gemm_kernel(pair)");
  return NM_OK;
}

template <int BN, bool A_MN, bool B_MN, int MODE>
static int launch_pair(const CUtensorMap& ma, const CUtensorMap& mb, int64_t M, int64_t N, int64_t K,
                       const TcEpilogue& epi, cudaStream_t s, int splits = 1, int kb_per_split = 0) {
  using Cfg = TcCfg<BN, 2>;
  auto kern = tc_gemm_kernel<BN, A_MN, B_MN, MODE, 4, 2>;
  static bool attr_done = false;
  if (!attr_done) {
    NM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  if (kb_per_split <= 0) kb_per_split = (int)ceil_div(K, TC_BK);
  return launch_pair_kernel(kern, Cfg::SMEM_BYTES, ceil_div(M, 2 * TC_BM) * ceil_div(N, BN) * splits, s, ma, mb, M,
                            N, K, epi, mn_desc_config(), splits, kb_per_split, TcExt{}, TcBatch{});
}

template <int BN, int MODE, bool MN = false>
static int launch_pair16(const CUtensorMap& ma, const CUtensorMap& mb, int64_t M, int64_t N, int64_t K,
                         const TcEpilogue& epi, const TcExt& ext, cudaStream_t s) {
  using Cfg = TcCfg<BN, 2>;
  auto kern = tc_gemm_kernel<BN, MN, MN, MODE, 2, 2>;
  static bool attr_done = false;
  if (!attr_done) {
    NM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  return launch_pair_kernel(kern, Cfg::SMEM_BYTES, ceil_div(M, 2 * TC_BM) * ceil_div(N, BN), s, ma, mb, M, N, K, epi,
                            MN ? MnDesc{SMEM_LAYOUT_SW128, 1024, 8192, 2048} : MnDesc{0, 0, 0, 0}, 1,
                            (int)ceil_div(K, 64), ext, TcBatch{});
}

// NMB200_TC_PDL=1: launch with programmatic stream serialization, so that a kernel's prologue overlaps the tail of
// its predecessor (the kernel waits with griddepcontrol.wait before it touches global memory)
static bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("NMB200_TC_PDL");
    return e && atoi(e) != 0;
  }();
  return on;
}

template <class Kern, class... Args>
static cudaError_t launch_one(Kern kern, unsigned grid, int smem_bytes, cudaStream_t s, Args... args) {
  if (!pdl_enabled()) {
    kern<<<grid, TC_THREADS, smem_bytes, s>>>(args...);
    return cudaSuccess;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

template <int BN, bool A_MN, bool B_MN, int MODE>
static int launch_cfg(const CUtensorMap& ma, const CUtensorMap& mb, int64_t M, int64_t N, int64_t K,
                      const TcEpilogue& epi, cudaStream_t s, int splits = 1, int kb_per_split = 0) {
  using Cfg = TcCfg<BN>;
  auto kern = tc_gemm_kernel<BN, A_MN, B_MN, MODE>;
  static bool attr_done = false;
  if (!attr_done) {
    NM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  const int64_t tiles = ceil_div(M, TC_BM) * ceil_div(N, BN) * splits;
  const int64_t grid = tiles < sm_count() ? tiles : sm_count();
  if (kb_per_split <= 0) kb_per_split = (int)ceil_div(K, TC_BK);
  NM_CUDA_TRY(launch_one(kern, (unsigned)grid, Cfg::SMEM_BYTES, s, ma, mb, M, N, K, epi, mn_desc_config(), splits,
                         kb_per_split, TcExt{}, TcBatch{}));
  NM_LAUNCH_CHECK("tc_gemm_kernel");
  return NM_OK;
}

// fp16 operands: no split-K, one wave structure as above.  MN = both operands MN-major (the
// reduction dimension strided: X^T . dY products), otherwise both K-major.
template <int BN, int MODE, bool MN = false>
static int launch_cfg16(const CUtensorMap& ma, const CUtensorMap& mb, int64_t M, int64_t N, int64_t K,
                        const TcEpilogue& epi, const TcExt& ext, cudaStream_t s) {
  using Cfg = TcCfg<BN>;
  auto kern = tc_gemm_kernel<BN, MN, MN, MODE, 2>;
  static bool attr_done = false;
  if (!attr_done) {
    NM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  const int64_t tiles = ceil_div(M, TC_BM) * ceil_div(N, BN);
  const int64_t grid = tiles < sm_count() ? tiles : sm_count();
  NM_CUDA_TRY(launch_one(kern, (unsigned)grid, Cfg::SMEM_BYTES, s, ma, mb, M, N, K, epi,
                         MN ? MnDesc{SMEM_LAYOUT_SW128, 1024, 8192, 2048} : MnDesc{0, 0, 0, 0}, 1,
                         (int)ceil_div(K, 64), ext, TcBatch{}));
  NM_LAUNCH_CHECK("tc_gemm_kernel(fp16)");
  return NM_OK;
}

// 2-D fp16 tensor [rows, cols] with row pitch ld (elements); box = {64 inner elements, box_rows},
// 128-byte swizzle.  K-major operand: inner = K, rows = M or N; MN-major: inner = M or N, rows = K
// (box_rows = 64: one k-block).
static int make_map16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld,
                      uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  NM_REQUIRE(fn != nullptr, NM_E_NO_DEVICE, "tc_gemm16: cuTensorMapEncodeTiled not available");
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NM_REQUIRE(r == CUDA_SUCCESS, NM_E_INVALID,
             "tc_gemm16: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
             (long long)rows, (long long)cols, (long long)ld);
  return NM_OK;
}

template <int BN, bool A_MN, bool B_MN, int MODE>
static int launch_batched(const CUtensorMap& ma, const CUtensorMap& mb, int64_t M, int64_t N, int64_t K,
                          const TcEpilogue& epi, const TcBatch& bt, cudaStream_t s) {
  using Cfg = TcCfg<BN>;
  auto kern = tc_gemm_kernel<BN, A_MN, B_MN, MODE>;
  static bool attr_done = false;
  if (!attr_done) {
    NM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  const int64_t tiles = ceil_div(M, TC_BM) * ceil_div(N, BN) * bt.count;
  const int64_t grid = tiles < sm_count() ? tiles : sm_count();
  NM_CUDA_TRY(launch_one(kern, (unsigned)grid, Cfg::SMEM_BYTES, s, ma, mb, M, N, K, epi, mn_desc_config(), 1,
                         (int)ceil_div(K, TC_BK), TcExt{}, bt));
  NM_LAUNCH_CHECK("tc_gemm_kernel(batched)");
  return NM_OK;
}

int tc_gemm_batched_launch(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t a_rows,
                           int64_t a_cols, int64_t lda, const float* B, int64_t b_rows, int64_t b_cols,
                           int64_t ldb, const TcEpilogue& epi, const TcBatch& bt, cudaStream_t s) {
  const bool a_mn = transA != 0, b_mn = transB == 0;
  NM_REQUIRE(bt.count >= 1 && bt.inner >= 1 && bt.count % bt.inner == 0, NM_E_INVALID,
             "tc_gemm_batched: %d problems do not split into groups of %d", bt.count, bt.inner);
  NM_REQUIRE(M >= 1 && N >= 1 && K >= 1 && (lda & 3) == 0 && (ldb & 3) == 0 &&
                 (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
             NM_E_INVALID, "tc_gemm_batched: operands must be 16-byte aligned with 16-byte row pitches");
  const bool attn = epi.mode == TC_EPI_SOFTMAX || epi.mode == TC_EPI_DSOFTMAX;
  NM_REQUIRE(attn || epi.mode == TC_EPI_DENSE, NM_E_INVALID, "tc_gemm_batched: unsupported epilogue %d", epi.mode);
  NM_REQUIRE(((bt.c_outer | bt.c_inner | epi.ldc) & 3) == 0 && (reinterpret_cast<uintptr_t>(epi.C) & 15) == 0,
             NM_E_INVALID, "tc_gemm_batched: output windows must be 16-byte aligned");
  const int bn = attn ? 128 : (N <= 64 ? 64 : 128);
  NM_REQUIRE(!attn || (N <= 128 && bt.n_pad >= N && bt.n_pad <= 128 && bt.m_pad >= M && !a_mn && !b_mn),
             NM_E_UNSUPPORTED, "tc_gemm_batched: attention epilogues take K-major operands and N <= 128 (N=%lld)",
             (long long)N);
  CUtensorMap ma, mb;
  int rc;
  if (!a_mn) rc = make_map(&ma, A, a_rows, a_cols, lda, TC_BK, TC_BM, false);
  else       rc = make_map(&ma, A, a_rows, a_cols, lda, 32, TC_BK, true);
  if (rc) return rc;
  if (!b_mn) rc = make_map(&mb, B, b_rows, b_cols, ldb, TC_BK, (uint32_t)bn, false);
  else       rc = make_map(&mb, B, b_rows, b_cols, ldb, 32, TC_BK, true);
  if (rc) return rc;
  if (epi.mode == TC_EPI_SOFTMAX) return launch_batched<128, false, false, TC_EPI_SOFTMAX>(ma, mb, M, N, K, epi, bt, s);
  if (epi.mode == TC_EPI_DSOFTMAX) return launch_batched<128, false, false, TC_EPI_DSOFTMAX>(ma, mb, M, N, K, epi, bt, s);
  NM_REQUIRE(b_mn, NM_E_UNSUPPORTED, "tc_gemm_batched: dense products take an MN-major B operand");
  if (bn == 64) {
    if (a_mn) return launch_batched<64, true, true, TC_EPI_DENSE>(ma, mb, M, N, K, epi, bt, s);
    return launch_batched<64, false, true, TC_EPI_DENSE>(ma, mb, M, N, K, epi, bt, s);
  }
  if (a_mn) return launch_batched<128, true, true, TC_EPI_DENSE>(ma, mb, M, N, K, epi, bt, s);
  return launch_batched<128, false, true, TC_EPI_DENSE>(ma, mb, M, N, K, epi, bt, s);
}

static int pick_bn(int64_t M, int64_t N, int64_t K) {
  static const int forced = [] {   // NMB200_TC_BN={64,128,160,256}: tile-width experiments (tools/gemm_sweep.py)
    const char* e = getenv("NMB200_TC_BN");
    const int v = e ? atoi(e) : 0;
    return (v == 64 || v == 128 || v == 160 || v == 256) ? v : 0;
  }();
  if (forced) return forced;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  // skinny output, very long reduction (dX of the vocabulary projection: 12800 x 300 x 32000):
  // the A stream dominates and is re-read once per N tile, so cover N with the fewest equal tiles
  if (K >= 8192 && N > 256 && N <= 320) return 160;
  const int64_t pad128 = ceil_div(N, 128) * 128, pad256 = ceil_div(N, 256) * 256;
  const int64_t tiles256 = ceil_div(M, TC_BM) * ceil_div(N, 256);
  if (pad256 == pad128 && tiles256 >= sm_count()) return 256;
  if (pad256 * 10 <= pad128 * 11 && tiles256 >= 2 * (int64_t)sm_count()) return 256;
  return 128;
}

int tc_gemm_launch(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A,
                   int64_t lda, const float* B, int64_t ldb, const TcEpilogue& epi, cudaStream_t s) {
  // op(A) is [M,K]: transA=0 -> A stored [M,K], K contiguous (K-major);
  //                 transA=1 -> A stored [K,M], M contiguous (MN-major).
  // op(B) is [K,N]: transB=0 -> B stored [K,N], N contiguous (MN-major);
  //                 transB=1 -> B stored [N,K], K contiguous (K-major).
  const bool a_mn = transA != 0, b_mn = transB == 0;
  const int bn = (epi.mode == TC_EPI_DENSE) ? pick_bn(M, N, K) : TC_XENT_BN;
  // split-K when the output alone cannot occupy the chip (and nothing forbids partial sums)
  int splits = 1;
  int kb_per = (int)ceil_div(K, TC_BK);
  if (epi.mode == TC_EPI_DENSE) {
    const int64_t tiles = ceil_div(M, TC_BM) * ceil_div(N, bn);
    const int64_t num_kb = ceil_div(K, TC_BK);
    if (epi.act == NM_ACT_NONE && tiles * 2 <= sm_count() && num_kb >= 16) {
      int64_t want = ceil_div(sm_count(), tiles);
      if (want > num_kb / 8) want = num_kb / 8;
      if (want > 1) {
        kb_per = (int)ceil_div(num_kb, want);
        splits = (int)ceil_div(num_kb, kb_per);  // every split owns >= 1 k-block
      }
    } else if (epi.act == NM_ACT_NONE && num_kb >= 256 && tiles < 4 * (int64_t)sm_count()) {
      // a few waves of very long tiles: the last, partly filled wave costs a whole tile time.
      // Cut K so that the work items fill the waves (static round-robin: ceil(items/SMs) rounds).
      int best = 1;
      double best_cost = (double)ceil_div(tiles, sm_count());
      for (int sp = 2; sp <= 8; ++sp) {
        if (num_kb / sp < 64) break;
        const double cost = (double)ceil_div(tiles * sp, sm_count()) / sp;
        if (cost < best_cost * 0.93) { best_cost = cost; best = sp; }
      }
      if (best > 1) {
        kb_per = (int)ceil_div(num_kb, best);
        splits = (int)ceil_div(num_kb, kb_per);
      }
    }
  }
  const bool pair = pair_wanted(M, N, K, bn, splits);
  CUtensorMap ma, mb;
  int rc;
  if (!a_mn) rc = make_map(&ma, A, M, K, lda, TC_BK, TC_BM, false);
  else       rc = make_map(&ma, A, K, M, lda, 32, TC_BK, true);
  if (rc) return rc;
  if (!b_mn) rc = make_map(&mb, B, N, K, ldb, TC_BK, (uint32_t)(pair ? bn / 2 : bn), false);
  else       rc = make_map(&mb, B, K, N, ldb, 32, TC_BK, true);
  if (rc) return rc;
  if (splits > 1 && epi.beta == 0.f)  // partial sums are added: start from zero
    NM_CUDA_TRY(cudaMemset2DAsync(epi.C, sizeof(float) * epi.ldc, 0, sizeof(float) * N, M, s));
#define NM_TC_DISPATCH_PAIR(BN_, MODE_)                                                                 \
  do {                                                                                                 \
    if (!a_mn && !b_mn) return launch_pair<BN_, false, false, MODE_>(ma, mb, M, N, K, epi, s, splits, kb_per); \
    if (!a_mn && b_mn) return launch_pair<BN_, false, true, MODE_>(ma, mb, M, N, K, epi, s, splits, kb_per);   \
    if (a_mn && !b_mn) return launch_pair<BN_, true, false, MODE_>(ma, mb, M, N, K, epi, s, splits, kb_per);   \
    return launch_pair<BN_, true, true, MODE_>(ma, mb, M, N, K, epi, s, splits, kb_per);                       \
  } while (0)
  if (pair) {
    if (epi.mode == TC_EPI_XENT_FWD) {
      if (b_mn) return launch_pair<256, false, true, TC_EPI_XENT_FWD>(ma, mb, M, N, K, epi, s);
      return launch_pair<256, false, false, TC_EPI_XENT_FWD>(ma, mb, M, N, K, epi, s);
    }
    if (epi.mode == TC_EPI_XENT_BWD) {
      if (b_mn) return launch_pair<256, false, true, TC_EPI_XENT_BWD>(ma, mb, M, N, K, epi, s);
      return launch_pair<256, false, false, TC_EPI_XENT_BWD>(ma, mb, M, N, K, epi, s);
    }
    if (bn == 128) NM_TC_DISPATCH_PAIR(128, TC_EPI_DENSE);
    NM_TC_DISPATCH_PAIR(256, TC_EPI_DENSE);
  }
#undef NM_TC_DISPATCH_PAIR
#define NM_TC_DISPATCH(BN_, MODE_)                                                                     \
  do {                                                                                                 \
    if (!a_mn && !b_mn) return launch_cfg<BN_, false, false, MODE_>(ma, mb, M, N, K, epi, s, splits, kb_per);  \
    if (!a_mn && b_mn) return launch_cfg<BN_, false, true, MODE_>(ma, mb, M, N, K, epi, s, splits, kb_per);    \
    if (a_mn && !b_mn) return launch_cfg<BN_, true, false, MODE_>(ma, mb, M, N, K, epi, s, splits, kb_per);    \
    return launch_cfg<BN_, true, true, MODE_>(ma, mb, M, N, K, epi, s, splits, kb_per);                        \
  } while (0)
  if (epi.mode == TC_EPI_XENT_FWD) {  // A is always K-major for the vocabulary projection
    if (b_mn) return launch_cfg<256, false, true, TC_EPI_XENT_FWD>(ma, mb, M, N, K, epi, s);
    return launch_cfg<256, false, false, TC_EPI_XENT_FWD>(ma, mb, M, N, K, epi, s);
  }
  if (epi.mode == TC_EPI_XENT_BWD) {
    if (b_mn) return launch_cfg<256, false, true, TC_EPI_XENT_BWD>(ma, mb, M, N, K, epi, s);
    return launch_cfg<256, false, false, TC_EPI_XENT_BWD>(ma, mb, M, N, K, epi, s);
  }
  if (bn == 64) NM_TC_DISPATCH(64, TC_EPI_DENSE);
  if (bn == 128) NM_TC_DISPATCH(128, TC_EPI_DENSE);
  if (bn == 160) NM_TC_DISPATCH(160, TC_EPI_DENSE);
  NM_TC_DISPATCH(256, TC_EPI_DENSE);
#undef NM_TC_DISPATCH
}

int tc_gemm16_mn_launch(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
                        int64_t ldb, const TcEpilogue& epi, const TcExt& ext, cudaStream_t s) {
  // A stored [K, M] (row pitch lda), B stored [K, N] (row pitch ldb): both MN-major
  NM_REQUIRE(M >= 1 && N >= 1 && K >= 1 && M <= 0x7fffffffLL && N <= 0x7fffffffLL && K <= 0x7fffffffLL,
             NM_E_INVALID, "tc_gemm16_mn: bad shape %lld x %lld x %lld", (long long)M, (long long)N, (long long)K);
  NM_REQUIRE((lda & 7) == 0 && (ldb & 7) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(B) & 15) == 0,
             NM_E_INVALID, "tc_gemm16_mn: fp16 operands need 16-byte aligned bases and row pitches");
  NM_REQUIRE(epi.mode == TC_EPI_DENSE, NM_E_INVALID, "tc_gemm16_mn: dense epilogue only");
  CUtensorMap ma, mb;
  int rc = make_map16(&ma, A, K, M, lda, 64);
  if (rc) return rc;
  rc = make_map16(&mb, B, K, N, ldb, 64);
  if (rc) return rc;
  // 64-column boxes: BN in {64, 128, 256}
  int bn = 256;
  if (N <= 64) bn = 64;
  else if (N <= 128) bn = 128;
  else if (ceil_div(M, TC_BM) * ceil_div(N, 256) < sm_count()) bn = 128;
  if (pair_wanted(M, N, K, bn, 1)) {
    if (bn == 128) return launch_pair16<128, TC_EPI_DENSE, true>(ma, mb, M, N, K, epi, ext, s);
    return launch_pair16<256, TC_EPI_DENSE, true>(ma, mb, M, N, K, epi, ext, s);
  }
  if (bn == 64) return launch_cfg16<64, TC_EPI_DENSE, true>(ma, mb, M, N, K, epi, ext, s);
  if (bn == 128) return launch_cfg16<128, TC_EPI_DENSE, true>(ma, mb, M, N, K, epi, ext, s);
  return launch_cfg16<256, TC_EPI_DENSE, true>(ma, mb, M, N, K, epi, ext, s);
}

int tc_gemm16_launch(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
                     int64_t ldb, const TcEpilogue& epi, const TcExt& ext, cudaStream_t s) {
  NM_REQUIRE(M >= 1 && N >= 1 && K >= 1 && M <= 0x7fffffffLL && N <= 0x7fffffffLL && K <= 0x7fffffffLL,
             NM_E_INVALID, "tc_gemm16: bad shape %lld x %lld x %lld", (long long)M, (long long)N, (long long)K);
  NM_REQUIRE((lda & 7) == 0 && (ldb & 7) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(B) & 15) == 0,
             NM_E_INVALID, "tc_gemm16: fp16 operands need 16-byte aligned bases and row pitches");
  const int bn = (epi.mode == TC_EPI_DENSE) ? pick_bn(M, N, K) : TC_XENT_BN;
  const bool pair = pair_wanted(M, N, K, bn, 1);
  CUtensorMap ma, mb;
  int rc = make_map16(&ma, A, M, K, lda, TC_BM);
  if (rc) return rc;
  rc = make_map16(&mb, B, N, K, ldb, (uint32_t)(pair ? bn / 2 : bn));
  if (rc) return rc;
  if (pair) {
    if (epi.mode == TC_EPI_XENT_FWD) return launch_pair16<256, TC_EPI_XENT_FWD>(ma, mb, M, N, K, epi, ext, s);
    if (epi.mode == TC_EPI_XENT_BWD16) return launch_pair16<256, TC_EPI_XENT_BWD16>(ma, mb, M, N, K, epi, ext, s);
    NM_REQUIRE(epi.mode == TC_EPI_DENSE, NM_E_INVALID, "tc_gemm16: unsupported epilogue %d", epi.mode);
    if (bn == 128) return launch_pair16<128, TC_EPI_DENSE>(ma, mb, M, N, K, epi, ext, s);
    return launch_pair16<256, TC_EPI_DENSE>(ma, mb, M, N, K, epi, ext, s);
  }
  if (epi.mode == TC_EPI_XENT_FWD) return launch_cfg16<256, TC_EPI_XENT_FWD>(ma, mb, M, N, K, epi, ext, s);
  if (epi.mode == TC_EPI_XENT_BWD16) return launch_cfg16<256, TC_EPI_XENT_BWD16>(ma, mb, M, N, K, epi, ext, s);
  NM_REQUIRE(epi.mode == TC_EPI_DENSE, NM_E_INVALID, "tc_gemm16: unsupported epilogue %d", epi.mode);
  if (bn == 64) return launch_cfg16<64, TC_EPI_DENSE>(ma, mb, M, N, K, epi, ext, s);
  if (bn == 128) return launch_cfg16<128, TC_EPI_DENSE>(ma, mb, M, N, K, epi, ext, s);
  if (bn == 160) return launch_cfg16<160, TC_EPI_DENSE>(ma, mb, M, N, K, epi, ext, s);
  return launch_cfg16<256, TC_EPI_DENSE>(ma, mb, M, N, K, epi, ext, s);
}

}  // namespace nm
