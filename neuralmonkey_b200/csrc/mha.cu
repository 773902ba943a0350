// K8: multi-head scaled dot-product attention core
// (attention/scaled_dot_product.py:160-214 of the reference) between the q/k/v
// projections and the output projection.  Reproduces the reference's masking
// semantics exactly: causal positions are REPLACED by -1e9 (tf.where), padded keys
// get E*m + (1-m)*(-1e9), both before the softmax; -1e9, not -inf.
//
// Round-1 implementation: one CTA per (query position, head, sentence); K/V rows of the
// head are re-read from L2 by each query CTA.  Correctness first; the tiled
// tensor-core version is later work ("what comes next").
#include "common.cuh"

namespace nm {

constexpr int MHA_THREADS = 128;
constexpr float MHA_MASK = -1e9f;

// dynamic smem: qs[dh] | e[Tk]
__global__ void __launch_bounds__(MHA_THREADS)
mha_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
               const float* __restrict__ key_mask, int causal, float* __restrict__ out,
               float* __restrict__ probs, int Tq, int Tk, int heads, int dh) {
  extern __shared__ float smem[];
  __shared__ float red[32];
  float* qs = smem;
  float* e = smem + dh;
  const int tq = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int D = heads * dh;
  const float scale = sqrtf((float)dh);
  for (int d = threadIdx.x; d < dh; d += MHA_THREADS)
    qs[d] = q[((int64_t)b * Tq + tq) * D + h * dh + d] / scale;
  __syncthreads();
  float lmax = -INFINITY;
  for (int tk = threadIdx.x; tk < Tk; tk += MHA_THREADS) {
    const float* kr = k + ((int64_t)b * Tk + tk) * D + h * dh;
    float acc = 0.f;
    for (int d = 0; d < dh; ++d) acc = fmaf(qs[d], kr[d], acc);
    if (causal && tk > tq) acc = MHA_MASK;
    if (key_mask) {
      const float m = key_mask[(int64_t)b * Tk + tk];
      acc = acc * m + (1.f - m) * MHA_MASK;
    }
    e[tk] = acc;
    lmax = fmaxf(lmax, acc);
  }
  const float mx = block_max(lmax, red);
  float lsum = 0.f;
  for (int tk = threadIdx.x; tk < Tk; tk += MHA_THREADS) {
    const float p = expf(e[tk] - mx);
    e[tk] = p;
    lsum += p;
  }
  const float s = block_sum(lsum, red);
  float* pr = probs + (((int64_t)b * heads + h) * Tq + tq) * Tk;
  for (int tk = threadIdx.x; tk < Tk; tk += MHA_THREADS) {
    const float p = e[tk] / s;
    e[tk] = p;
    pr[tk] = p;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < dh; d += MHA_THREADS) {
    const float* vc = v + (int64_t)b * Tk * D + h * dh + d;
    float acc = 0.f;
    for (int tk = 0; tk < Tk; ++tk) acc = fmaf(e[tk], vc[(int64_t)tk * D], acc);
    out[((int64_t)b * Tq + tq) * D + h * dh + d] = acc;
  }
}

// backward A: per query row: dE (pre-mask gradient) and dq.  smem: dos[dh] | de[Tk]
__global__ void __launch_bounds__(MHA_THREADS)
mha_bwd_q_kernel(const float* __restrict__ k, const float* __restrict__ v,
                 const float* __restrict__ key_mask, int causal, const float* __restrict__ probs,
                 const float* __restrict__ dout, float* __restrict__ dq, float* __restrict__ de_out,
                 int Tq, int Tk, int heads, int dh) {
  extern __shared__ float smem[];
  __shared__ float red[32];
  float* dos = smem;
  float* de = smem + dh;
  const int tq = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int D = heads * dh;
  for (int d = threadIdx.x; d < dh; d += MHA_THREADS)
    dos[d] = dout[((int64_t)b * Tq + tq) * D + h * dh + d];
  __syncthreads();
  const float* pr = probs + (((int64_t)b * heads + h) * Tq + tq) * Tk;
  float lsum = 0.f;
  for (int tk = threadIdx.x; tk < Tk; tk += MHA_THREADS) {
    const float* vr = v + ((int64_t)b * Tk + tk) * D + h * dh;
    float dp = 0.f;
    for (int d = 0; d < dh; ++d) dp = fmaf(dos[d], vr[d], dp);
    de[tk] = dp;
    lsum += dp * pr[tk];
  }
  const float pdp = block_sum(lsum, red);
  float* der = de_out + (((int64_t)b * heads + h) * Tq + tq) * Tk;
  for (int tk = threadIdx.x; tk < Tk; tk += MHA_THREADS) {
    float g = pr[tk] * (de[tk] - pdp);
    if (key_mask) g *= key_mask[(int64_t)b * Tk + tk];  // d(E*m + c)/dE = m
    if (causal && tk > tq) g = 0.f;                     // tf.where: no gradient to replaced entries
    de[tk] = g;
    der[tk] = g;
  }
  __syncthreads();
  const float scale = sqrtf((float)dh);
  for (int d = threadIdx.x; d < dh; d += MHA_THREADS) {
    const float* kc = k + (int64_t)b * Tk * D + h * dh + d;
    float acc = 0.f;
    for (int tk = 0; tk < Tk; ++tk) acc = fmaf(de[tk], kc[(int64_t)tk * D], acc);
    dq[((int64_t)b * Tq + tq) * D + h * dh + d] = acc / scale;
  }
}

// backward B: per key row: dk and dv.
__global__ void __launch_bounds__(MHA_THREADS)
mha_bwd_kv_kernel(const float* __restrict__ q, const float* __restrict__ probs,
                  const float* __restrict__ de, const float* __restrict__ dout,
                  float* __restrict__ dk, float* __restrict__ dv, int Tq, int Tk, int heads, int dh) {
  const int tk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int D = heads * dh;
  const float scale = sqrtf((float)dh);
  const float* pcol = probs + ((int64_t)b * heads + h) * Tq * Tk + tk;
  const float* dcol = de + ((int64_t)b * heads + h) * Tq * Tk + tk;
  for (int d = threadIdx.x; d < dh; d += MHA_THREADS) {
    float ak = 0.f, av = 0.f;
    for (int tq = 0; tq < Tq; ++tq) {
      const int64_t o = ((int64_t)b * Tq + tq) * D + h * dh + d;
      ak = fmaf(dcol[(int64_t)tq * Tk], q[o] / scale, ak);
      av = fmaf(pcol[(int64_t)tq * Tk], dout[o], av);
    }
    const int64_t o = ((int64_t)b * Tk + tk) * D + h * dh + d;
    dk[o] = ak;
    dv[o] = av;
  }
}

}  // namespace nm

using namespace nm;

extern "C" {

int nm_mha_fwd(const float* q, const float* k, const float* v, const float* key_mask, int causal,
               float* out, float* probs, int64_t B, int64_t Tq, int64_t Tk, int64_t heads,
               int64_t dh, void* stream) {
  NM_REQUIRE(q && k && v && out && probs, NM_E_INVALID, "nm_mha_fwd: null pointer");
  NM_REQUIRE(B > 0 && Tq > 0 && Tk > 0 && heads > 0 && dh > 0, NM_E_INVALID, "nm_mha_fwd: bad sizes");
  NM_REQUIRE(B <= 65535 && heads <= 65535, NM_E_UNSUPPORTED, "nm_mha_fwd: grid too large");
  const size_t smem = sizeof(float) * (size_t)(dh + Tk);
  NM_REQUIRE(smem <= 48 * 1024, NM_E_UNSUPPORTED, "nm_mha_fwd: Tk+dh too large for this kernel");
  dim3 grid((unsigned)Tq, (unsigned)heads, (unsigned)B);
  mha_fwd_kernel<<<grid, MHA_THREADS, smem, (cudaStream_t)stream>>>(q, k, v, key_mask, causal, out,
                                                                   probs, (int)Tq, (int)Tk,
                                                                   (int)heads, (int)dh);
  NM_LAUNCH_CHECK("nm_mha_fwd");
  return NM_OK;
}

int nm_mha_bwd(const float* q, const float* k, const float* v, const float* key_mask, int causal,
               const float* probs, const float* dout, float* dq, float* dk, float* dv,
               float* de_work, int64_t B, int64_t Tq, int64_t Tk, int64_t heads, int64_t dh,
               void* stream) {
  NM_REQUIRE(q && k && v && probs && dout && dq && dk && dv && de_work, NM_E_INVALID,
             "nm_mha_bwd: null pointer");
  NM_REQUIRE(B > 0 && Tq > 0 && Tk > 0 && heads > 0 && dh > 0, NM_E_INVALID, "nm_mha_bwd: bad sizes");
  NM_REQUIRE(B <= 65535 && heads <= 65535, NM_E_UNSUPPORTED, "nm_mha_bwd: grid too large");
  const size_t smem = sizeof(float) * (size_t)(dh + Tk);
  NM_REQUIRE(smem <= 48 * 1024, NM_E_UNSUPPORTED, "nm_mha_bwd: Tk+dh too large for this kernel");
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid_q((unsigned)Tq, (unsigned)heads, (unsigned)B);
  mha_bwd_q_kernel<<<grid_q, MHA_THREADS, smem, s>>>(k, v, key_mask, causal, probs, dout, dq, de_work,
                                                     (int)Tq, (int)Tk, (int)heads, (int)dh);
  NM_LAUNCH_CHECK("nm_mha_bwd(q)");
  dim3 grid_k((unsigned)Tk, (unsigned)heads, (unsigned)B);
  mha_bwd_kv_kernel<<<grid_k, MHA_THREADS, 0, s>>>(q, probs, de_work, dout, dk, dv, (int)Tq, (int)Tk,
                                                   (int)heads, (int)dh);
  NM_LAUNCH_CHECK("nm_mha_bwd(kv)");
  return NM_OK;
}

}  // extern "C"
