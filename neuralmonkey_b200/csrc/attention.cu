// K4: Bahdanau (MLP) attention, attention/feed_forward.py:120-166 of the reference,
// for NQ query steps at once.
//
//   e[b,q,t] = sum_a v[a] * tanh(keys[b,t,a] + qproj[b,q,a]) + bias
//   p        = softmax_t(e)                       (over ALL Tx, padding included)
//   w        = p*mask / (sum_t p*mask + 1e-8)     (reference :139-144)
//   ctx[b,q] = sum_t w[b,q,t] * values[b,t,:]
//
// The kernel is bound by the tanh count B*NQ*Tx*A (one ex2 + one rcp on the SFU
// each) and by reads of the encoder tensors; a CTA therefore handles QCH queries
// of one sentence so every keys/values element loaded from L2/HBM is reused QCH
// times from registers, lanes walk the contiguous A (or C) axis for coalesced
// 128-byte requests, and the softmax over Tx is a warp-shuffle reduction.
#include "common.cuh"

namespace nm {

constexpr int ATT_QCH = 8;       // queries per CTA (forward / energy-gradient kernels)
constexpr int ATT_THREADS = 256;

// |abs err| <= ~2e-7: 1 - 2/(1+exp(2x)) with SFU ex2 and rcp; saturates correctly.
__device__ __forceinline__ float fast_tanh(float x) {
  return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x));
}

// dynamic smem: vs[A] | qs[QCH][A] | es[QCH][Tx]
__global__ void __launch_bounds__(ATT_THREADS)
bahdanau_fwd_kernel(const float* __restrict__ keys, const float* __restrict__ values,
                    const float* __restrict__ mask, const float* __restrict__ qproj,
                    const float* __restrict__ v, const float* __restrict__ bias,
                    float* __restrict__ energies, float* __restrict__ weights,
                    float* __restrict__ ctx, int Tx, int NQ, int A, int C) {
  extern __shared__ float smem[];
  float* vs = smem;
  float* qs = vs + A;
  float* es = qs + ATT_QCH * A;
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * ATT_QCH;
  const int nq = min(ATT_QCH, NQ - q0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = ATT_THREADS / 32;

  for (int a = threadIdx.x; a < A; a += ATT_THREADS) vs[a] = v[a];
  for (int i = threadIdx.x; i < ATT_QCH * A; i += ATT_THREADS) {
    const int j = i / A, a = i - j * A;
    qs[i] = (j < nq) ? qproj[((int64_t)b * NQ + q0 + j) * A + a] : 0.f;
  }
  __syncthreads();

  const float bs = bias[0];
  for (int t = warp; t < Tx; t += nwarps) {
    const float* kr = keys + ((int64_t)b * Tx + t) * A;
    float acc[ATT_QCH];
#pragma unroll
    for (int j = 0; j < ATT_QCH; ++j) acc[j] = 0.f;
    for (int a = lane; a < A; a += 32) {
      const float k = kr[a], vv = vs[a];
#pragma unroll
      for (int j = 0; j < ATT_QCH; ++j) acc[j] = fmaf(vv, fast_tanh(k + qs[j * A + a]), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < ATT_QCH; ++j) {
      const float e = warp_sum(acc[j]) + bs;
      if (lane == 0) es[j * Tx + t] = e;
    }
  }
  __syncthreads();

  // softmax over Tx, then mask + renormalise: one warp per query
  for (int j = warp; j < nq; j += nwarps) {
    float* er = es + j * Tx;
    const int64_t orow = ((int64_t)b * NQ + q0 + j) * Tx;
    float mx = -INFINITY;
    for (int t = lane; t < Tx; t += 32) mx = fmaxf(mx, er[t]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int t = lane; t < Tx; t += 32) s += expf(er[t] - mx);
    s = warp_sum(s);
    float ws = 0.f;
    for (int t = lane; t < Tx; t += 32) {
      const float e = er[t];
      if (energies) energies[orow + t] = e;
      float p = expf(e - mx) / s;
      if (mask) p *= mask[(int64_t)b * Tx + t];
      er[t] = p;
      ws += p;
    }
    if (mask) {
      const float norm = warp_sum(ws) + 1e-8f;
      for (int t = lane; t < Tx; t += 32) er[t] = er[t] / norm;
    }
    __syncwarp();
    for (int t = lane; t < Tx; t += 32) weights[orow + t] = er[t];
  }
  __syncthreads();

  // ctx[b,q,c] = sum_t w[q,t] * values[b,t,c]
  for (int c = threadIdx.x; c < C; c += ATT_THREADS) {
    float acc[ATT_QCH];
#pragma unroll
    for (int j = 0; j < ATT_QCH; ++j) acc[j] = 0.f;
    const float* vc = values + (int64_t)b * Tx * C + c;
    for (int t = 0; t < Tx; ++t) {
      const float val = vc[(int64_t)t * C];
#pragma unroll
      for (int j = 0; j < ATT_QCH; ++j) acc[j] = fmaf(es[j * Tx + t], val, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < ATT_QCH; ++j)
      if (j < nq) ctx[((int64_t)b * NQ + q0 + j) * C + c] = acc[j];
  }
}

// Backward A: de[b,q,t] from dctx.  dynamic smem: dws[QCH][Tx]
__global__ void __launch_bounds__(ATT_THREADS)
bahdanau_bwd_energy_kernel(const float* __restrict__ values, const float* __restrict__ mask,
                           const float* __restrict__ energies, const float* __restrict__ weights,
                           const float* __restrict__ dctx, float* __restrict__ de,
                           float* __restrict__ dbias, int Tx, int NQ, int C) {
  extern __shared__ float smem[];
  float* dws = smem;  // [QCH][Tx]
  __shared__ float red[32];
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * ATT_QCH;
  const int nq = min(ATT_QCH, NQ - q0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = ATT_THREADS / 32;

  // dw[q,t] = sum_c dctx[b,q,c] * values[b,t,c]: one warp per t, lanes over C
  for (int t = warp; t < Tx; t += nwarps) {
    const float* vr = values + ((int64_t)b * Tx + t) * C;
    float acc[ATT_QCH];
#pragma unroll
    for (int j = 0; j < ATT_QCH; ++j) acc[j] = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float val = vr[c];
#pragma unroll
      for (int j = 0; j < ATT_QCH; ++j)
        if (j < nq) acc[j] = fmaf(val, dctx[((int64_t)b * NQ + q0 + j) * C + c], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < ATT_QCH; ++j) {
      const float s = warp_sum(acc[j]);
      if (lane == 0) dws[j * Tx + t] = s;
    }
  }
  __syncthreads();

  float dbias_local = 0.f;
  for (int j = warp; j < nq; j += nwarps) {
    const int64_t row = ((int64_t)b * NQ + q0 + j) * Tx;
    const float* er = energies + row;
    const float* wr = weights + row;
    float* dwr = dws + j * Tx;
    // recompute p and the renormaliser N
    float mx = -INFINITY;
    for (int t = lane; t < Tx; t += 32) mx = fmaxf(mx, er[t]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int t = lane; t < Tx; t += 32) s += expf(er[t] - mx);
    s = warp_sum(s);
    float norm = 1.f;
    float dot_dw_w = 0.f;
    if (mask) {
      float ws = 0.f;
      for (int t = lane; t < Tx; t += 32) {
        ws += expf(er[t] - mx) / s * mask[(int64_t)b * Tx + t];
        dot_dw_w += dwr[t] * wr[t];
      }
      norm = warp_sum(ws) + 1e-8f;
      dot_dw_w = warp_sum(dot_dw_w);
    }
    // dp_t = mask_t * (dw_t - sum dw.w) / N   (no mask: dp = dw)
    float pdp = 0.f;
    for (int t = lane; t < Tx; t += 32) {
      const float p = expf(er[t] - mx) / s;
      float dp = dwr[t];
      if (mask) dp = mask[(int64_t)b * Tx + t] * (dp - dot_dw_w) / norm;
      dwr[t] = dp;
      pdp += p * dp;
    }
    pdp = warp_sum(pdp);
    for (int t = lane; t < Tx; t += 32) {
      const float p = expf(er[t] - mx) / s;
      const float g = p * (dwr[t] - pdp);
      de[row + t] = g;
      dbias_local += g;
    }
  }
  dbias_local = block_sum(dbias_local, red);
  if (threadIdx.x == 0 && dbias) atomicAdd(dbias, dbias_local);
}

// Backward B: dkeys, dqproj, dv.  grid (ceil(A/128), B), block 128: one a-column per
// thread.  dynamic smem: des[NQ][Tx] | qps[NQ][128] | dqs[NQ][128]
constexpr int ATT_ACH = 128;
__global__ void __launch_bounds__(ATT_ACH)
bahdanau_bwd_keys_kernel(const float* __restrict__ keys, const float* __restrict__ qproj,
                         const float* __restrict__ v, const float* __restrict__ de,
                         float* __restrict__ dkeys, float* __restrict__ dqproj,
                         float* __restrict__ dv, int Tx, int NQ, int A) {
  extern __shared__ float smem[];
  float* des = smem;                 // [NQ][Tx]
  float* qps = des + NQ * Tx;        // [NQ][ACH]
  float* dqs = qps + NQ * ATT_ACH;   // [NQ][ACH]
  const int b = blockIdx.y;
  const int a = blockIdx.x * ATT_ACH + threadIdx.x;
  const bool ok = a < A;
  for (int i = threadIdx.x; i < NQ * Tx; i += ATT_ACH) des[i] = de[(int64_t)b * NQ * Tx + i];
  for (int q = 0; q < NQ; ++q) {
    qps[q * ATT_ACH + threadIdx.x] = ok ? qproj[((int64_t)b * NQ + q) * A + a] : 0.f;
    dqs[q * ATT_ACH + threadIdx.x] = 0.f;
  }
  __syncthreads();
  const float va = ok ? v[a] : 0.f;
  float dv_acc = 0.f;
  if (ok) {
    for (int t = 0; t < Tx; ++t) {
      const float k = keys[((int64_t)b * Tx + t) * A + a];
      float dk = 0.f;
      for (int q = 0; q < NQ; ++q) {
        const float th = fast_tanh(k + qps[q * ATT_ACH + threadIdx.x]);
        const float d = des[q * Tx + t];
        const float g = d * va * (1.f - th * th);
        dk += g;
        dqs[q * ATT_ACH + threadIdx.x] += g;
        dv_acc = fmaf(d, th, dv_acc);
      }
      dkeys[((int64_t)b * Tx + t) * A + a] = dk;
    }
    for (int q = 0; q < NQ; ++q)
      dqproj[((int64_t)b * NQ + q) * A + a] = dqs[q * ATT_ACH + threadIdx.x];
    atomicAdd(dv + a, dv_acc);
  }
}

// Register variant of the keys kernel for NQ <= NQMAX: the NQ query projections and the NQ
// dqproj accumulators of this thread's column live in registers, so the inner loop is one
// broadcast shared-memory load (de), one MUFU.TANH and four FMAs per (t, q) - no read-modify-write
// of shared memory.  dynamic smem: des[NQ][Tx] only.
template <int NQMAX>
__global__ void __launch_bounds__(ATT_ACH)
bahdanau_bwd_keys_reg_kernel(const float* __restrict__ keys, const float* __restrict__ qproj,
                             const float* __restrict__ v, const float* __restrict__ de,
                             float* __restrict__ dkeys, float* __restrict__ dqproj,
                             float* __restrict__ dv, int Tx, int NQ, int A) {
  extern __shared__ float smem[];
  float* des = smem;                 // [Tx][NQMAX]: q contiguous for vector broadcast loads
  const int b = blockIdx.y;
  const int a = blockIdx.x * ATT_ACH + threadIdx.x;
  const bool ok = a < A;
  for (int i = threadIdx.x; i < Tx * NQMAX; i += ATT_ACH) {
    const int t = i / NQMAX, q = i - t * NQMAX;
    des[i] = q < NQ ? de[((int64_t)b * NQ + q) * Tx + t] : 0.f;
  }
  float qp[NQMAX], dq[NQMAX];
#pragma unroll
  for (int q = 0; q < NQMAX; ++q) {
    qp[q] = (ok && q < NQ) ? qproj[((int64_t)b * NQ + q) * A + a] : 0.f;
    dq[q] = 0.f;
  }
  __syncthreads();
  if (!ok) return;
  const float va = v[a];
  float dv_acc = 0.f;
  const float* kp = keys + (int64_t)b * Tx * A + a;
  float* dkp = dkeys + (int64_t)b * Tx * A + a;
  float k_next = kp[0];
  for (int t = 0; t < Tx; ++t) {
    const float k = k_next;
    if (t + 1 < Tx) k_next = kp[(int64_t)(t + 1) * A];
    const float4* d4 = reinterpret_cast<const float4*>(des + t * NQMAX);
    float dk = 0.f;
#pragma unroll
    for (int q4 = 0; q4 < NQMAX / 4; ++q4) {
      const float4 d = d4[q4];
      const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = 4 * q4 + j;
        const float th = fast_tanh(k + qp[q]);
        const float g = dd[j] * (va - va * th * th);   // zero for the padded q (de = 0)
        dk += g;
        dq[q] += g;
        dv_acc = fmaf(dd[j], th, dv_acc);
      }
    }
    dkp[(int64_t)t * A] = dk;
  }
#pragma unroll
  for (int q = 0; q < NQMAX; ++q)
    if (q < NQ) dqproj[((int64_t)b * NQ + q) * A + a] = dq[q];
  atomicAdd(dv + a, dv_acc);
}

// Backward C: dvalues[b,t,c] = sum_q w[b,q,t] * dctx[b,q,c].  grid (ceil(C/128), B).
// dynamic smem: ws[NQ][Tx] | dcs[NQ][128]
__global__ void __launch_bounds__(ATT_ACH)
bahdanau_bwd_values_kernel(const float* __restrict__ weights, const float* __restrict__ dctx,
                           float* __restrict__ dvalues, int Tx, int NQ, int C) {
  extern __shared__ float smem[];
  float* ws = smem;              // [NQ][Tx]
  float* dcs = ws + NQ * Tx;     // [NQ][ACH]
  const int b = blockIdx.y;
  const int c = blockIdx.x * ATT_ACH + threadIdx.x;
  const bool ok = c < C;
  for (int i = threadIdx.x; i < NQ * Tx; i += ATT_ACH) ws[i] = weights[(int64_t)b * NQ * Tx + i];
  for (int q = 0; q < NQ; ++q)
    dcs[q * ATT_ACH + threadIdx.x] = ok ? dctx[((int64_t)b * NQ + q) * C + c] : 0.f;
  __syncthreads();
  if (!ok) return;
  for (int t = 0; t < Tx; ++t) {
    float acc = 0.f;
    for (int q = 0; q < NQ; ++q) acc = fmaf(ws[q * Tx + t], dcs[q * ATT_ACH + threadIdx.x], acc);
    dvalues[((int64_t)b * Tx + t) * C + c] = acc;
  }
}

constexpr size_t ATT_SMEM_LIMIT = 200 * 1024;

}  // namespace nm

using namespace nm;

extern "C" {

int nm_bahdanau_fwd(const float* keys, const float* values, const float* mask, const float* qproj,
                    const float* v, const float* bias, float* energies, float* weights, float* ctx,
                    int64_t B, int64_t Tx, int64_t NQ, int64_t A, int64_t C, void* stream) {
  NM_REQUIRE(keys && values && qproj && v && bias && weights && ctx, NM_E_INVALID,
             "nm_bahdanau_fwd: null pointer");
  NM_REQUIRE(B > 0 && Tx > 0 && NQ > 0 && A > 0 && C > 0, NM_E_INVALID, "nm_bahdanau_fwd: bad sizes");
  NM_REQUIRE(B <= 65535, NM_E_UNSUPPORTED, "nm_bahdanau_fwd: B > 65535");
  const size_t smem = sizeof(float) * (size_t)(A + ATT_QCH * A + ATT_QCH * Tx);
  NM_REQUIRE(smem <= ATT_SMEM_LIMIT, NM_E_UNSUPPORTED,
             "nm_bahdanau_fwd: A=%lld Tx=%lld need %zu B of shared memory", (long long)A,
             (long long)Tx, smem);
  static bool attr_set = false;
  if (!attr_set) {
    NM_CUDA_TRY(cudaFuncSetAttribute(bahdanau_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)ATT_SMEM_LIMIT));
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div(NQ, ATT_QCH), (unsigned)B);
  bahdanau_fwd_kernel<<<grid, ATT_THREADS, smem, (cudaStream_t)stream>>>(
      keys, values, mask, qproj, v, bias, energies, weights, ctx, (int)Tx, (int)NQ, (int)A, (int)C);
  NM_LAUNCH_CHECK("nm_bahdanau_fwd");
  return NM_OK;
}

int nm_bahdanau_bwd(const float* keys, const float* values, const float* mask, const float* qproj,
                    const float* v, const float* energies, const float* weights, const float* dctx,
                    float* dkeys, float* dvalues, float* dqproj, float* dv, float* dbias, float* de_work,
                    int64_t B, int64_t Tx, int64_t NQ, int64_t A, int64_t C, void* stream) {
  NM_REQUIRE(keys && values && qproj && v && energies && weights && dctx && dkeys && dvalues &&
                 dqproj && dv && dbias && de_work,
             NM_E_INVALID, "nm_bahdanau_bwd: null pointer");
  NM_REQUIRE(B > 0 && Tx > 0 && NQ > 0 && A > 0 && C > 0, NM_E_INVALID, "nm_bahdanau_bwd: bad sizes");
  NM_REQUIRE(B <= 65535, NM_E_UNSUPPORTED, "nm_bahdanau_bwd: B > 65535");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem_a = sizeof(float) * (size_t)(ATT_QCH * Tx);
  const size_t smem_b = sizeof(float) * (size_t)(NQ * Tx + 2 * NQ * ATT_ACH);
  const size_t smem_c = sizeof(float) * (size_t)(NQ * Tx + NQ * ATT_ACH);
  NM_REQUIRE(smem_a <= ATT_SMEM_LIMIT && smem_b <= ATT_SMEM_LIMIT, NM_E_UNSUPPORTED,
             "nm_bahdanau_bwd: NQ=%lld Tx=%lld need %zu B of shared memory", (long long)NQ,
             (long long)Tx, smem_b);
  static bool attr_set = false;
  if (!attr_set) {
    NM_CUDA_TRY(cudaFuncSetAttribute(bahdanau_bwd_energy_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_LIMIT));
    NM_CUDA_TRY(cudaFuncSetAttribute(bahdanau_bwd_keys_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_LIMIT));
    NM_CUDA_TRY(cudaFuncSetAttribute(bahdanau_bwd_values_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_LIMIT));
    NM_CUDA_TRY(cudaFuncSetAttribute(bahdanau_bwd_keys_reg_kernel<32>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    NM_CUDA_TRY(cudaFuncSetAttribute(bahdanau_bwd_keys_reg_kernel<64>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  dim3 grid_a((unsigned)ceil_div(NQ, ATT_QCH), (unsigned)B);
  bahdanau_bwd_energy_kernel<<<grid_a, ATT_THREADS, smem_a, s>>>(values, mask, energies, weights, dctx,
                                                                de_work, dbias, (int)Tx, (int)NQ, (int)C);
  NM_LAUNCH_CHECK("nm_bahdanau_bwd(energy)");
  dim3 grid_b((unsigned)ceil_div(A, ATT_ACH), (unsigned)B);
  if (NQ <= 64 && sizeof(float) * Tx * 64 <= 96 * 1024) {
    if (NQ <= 32)
      bahdanau_bwd_keys_reg_kernel<32><<<grid_b, ATT_ACH, sizeof(float) * Tx * 32, s>>>(
          keys, qproj, v, de_work, dkeys, dqproj, dv, (int)Tx, (int)NQ, (int)A);
    else
      bahdanau_bwd_keys_reg_kernel<64><<<grid_b, ATT_ACH, sizeof(float) * Tx * 64, s>>>(
          keys, qproj, v, de_work, dkeys, dqproj, dv, (int)Tx, (int)NQ, (int)A);
  } else {
    bahdanau_bwd_keys_kernel<<<grid_b, ATT_ACH, smem_b, s>>>(keys, qproj, v, de_work, dkeys, dqproj, dv,
                                                            (int)Tx, (int)NQ, (int)A);
  }
  NM_LAUNCH_CHECK("nm_bahdanau_bwd(keys)");
  dim3 grid_c((unsigned)ceil_div(C, ATT_ACH), (unsigned)B);
  bahdanau_bwd_values_kernel<<<grid_c, ATT_ACH, smem_c, s>>>(weights, dctx, dvalues, (int)Tx, (int)NQ,
                                                            (int)C);
  NM_LAUNCH_CHECK("nm_bahdanau_bwd(values)");
  return NM_OK;
}

}  // extern "C"
