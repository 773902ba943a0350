// Persistent GRU sequence kernels for thread-block clusters (sm_100a).
//
// One launch runs the whole time loop.  A cluster of 8 CTAs owns a slice of the batch
// (Bc sentences); inside it, CTA r owns hidden units [r*UW, (r+1)*UW), UW = ceil(H/8).
// The recurrent weights never leave the SM: each thread keeps the three weight vectors
// of ONE hidden unit restricted to ONE of 8 reduction slices in REGISTERS (3*4*CH
// floats), so the inner product streams only the state vector from shared memory
// (one 16-byte broadcast load per 8 or 4 FMAs) and is FMA-issue bound, not
// shared-memory bound.  The 8 reduction slices of a unit sit in 8 adjacent lanes
// and are combined with three shuffle steps.
//
// The two matmuls of a TF GRUCell step are dependent (the candidate needs r*h for ALL
// units), so a step has two phases separated by hardware cluster barriers; the vectors
// exchanged between the phases (h, r*h; backward: dz_c, dz_r, dz_u) are exactly the
// tensors the backward pass / the weight-gradient GEMMs need in HBM anyway, so the
// exchange costs only an L2 read of Bc*H floats per CTA per phase.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace nm {

constexpr int GC_CLUSTER = 8;          // CTAs per cluster = reduction... = unit slices
constexpr int GC_KS = 8;               // reduction slices per unit (lanes)
constexpr int GC_WARPS = 10;           // 4 units per warp -> up to 40 units per CTA
constexpr int GC_THREADS = GC_WARPS * 32;
constexpr int GC_MAX_UNITS = GC_WARPS * 4;

__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                   : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ float reduce8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// Geometry shared by forward and backward.
struct GcGeom {
  int H, SL, SLP, ROW, Bc;
  // SL  = ceil(H/8): units per CTA and reduction-slice length
  // SLP = padded slice stride in the smem vector buffer (multiple of 4, odd number of
  //       16-byte chunks -> the 8 slice lanes hit 8 different bank groups)
  // ROW = 8*SLP floats per batch row
};

__host__ __device__ inline int gc_slice_pad(int SL) {
  int chunks = (SL + 3) / 4;
  if ((chunks & 1) == 0) chunks += 1;
  return chunks * 4;
}

// Stage `width` contiguous floats per row (rows b0.., row pitch `pitch`, column offset
// applied by the caller) into the sliced smem layout; rows >= nrows are zero-filled.
__device__ __forceinline__ void gc_load_vec(float* __restrict__ vec, const float* __restrict__ src,
                                            int64_t pitch, int nrows, const GcGeom& g) {
  const int total = g.Bc * g.H;
  for (int idx = threadIdx.x; idx < total; idx += GC_THREADS) {
    const int b = idx / g.H, k = idx - b * g.H;
    const float v = (b < nrows) ? __ldcg(src + (int64_t)b * pitch + k) : 0.f;
    vec[b * g.ROW + (k / g.SL) * g.SLP + (k % g.SL)] = v;
  }
}

// acc[i] for 4 consecutive rows: sum over this lane's slice of vec[row][k] * w[k].
template <int CH>
__device__ __forceinline__ void gc_dot4(const float* __restrict__ vec, int row0, int ks,
                                        const GcGeom& g, const float (&w)[4 * CH], float (&acc)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = 0.f;
  const float* base = vec + row0 * g.ROW + ks * g.SLP;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = *reinterpret_cast<const float4*>(base + i * g.ROW + 4 * c);
      acc[i] = fmaf(v.x, w[4 * c], acc[i]);
      acc[i] = fmaf(v.y, w[4 * c + 1], acc[i]);
      acc[i] = fmaf(v.z, w[4 * c + 2], acc[i]);
      acc[i] = fmaf(v.w, w[4 * c + 3], acc[i]);
    }
  }
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
struct GcFwdArgs {
  const float* xproj;   // [B,T,3H]
  const float* Wgh;     // [H,2H]
  const float* Wch;     // [H,H]
  const float* h0;      // [B,H] or null
  const int32_t* lengths;
  const float* drop_mask;
  float* states;        // [B,T,H]
  float* raw_states;    // or null
  float* final_state;   // [B,H]
  float* gates;         // [B,T,3H]
  float* hprev;         // [B,T,H]
  float* rh;            // [B,T,H]
  int B, T, H, Bc, reverse;
};

template <int CH>
__global__ void __launch_bounds__(GC_THREADS, 1) gru_seq_fwd_cluster_kernel(GcFwdArgs a) {
  extern __shared__ __align__(16) float gc_smem[];
  float* vec = gc_smem;  // [Bc][ROW]
  GcGeom g;
  g.H = a.H;
  g.SL = (a.H + GC_KS - 1) / GC_KS;
  g.SLP = gc_slice_pad(g.SL);
  g.ROW = GC_KS * g.SLP;
  g.Bc = a.Bc;
  const int rank = (int)cluster_rank();
  const int cluster_id = blockIdx.x / GC_CLUSTER;
  const int b0 = cluster_id * a.Bc;
  const int nrows = min(a.Bc, a.B - b0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ks = lane & 7;
  const int unit_local = warp * 4 + (lane >> 3);
  const int j = rank * g.SL + unit_local;               // hidden unit of this thread
  const bool unit_ok = (unit_local < g.SL) && (j < a.H);
  const int H = a.H, T = a.T;

  // this thread's weights: columns (r_j, u_j, c_j), rows k in slice ks
  float wr[4 * CH], wu[4 * CH], wc[4 * CH];
#pragma unroll
  for (int i = 0; i < 4 * CH; ++i) {
    const int k = ks * g.SL + i;
    const bool ok = unit_ok && i < g.SL && k < H;
    wr[i] = ok ? a.Wgh[(int64_t)k * 2 * H + j] : 0.f;
    wu[i] = ok ? a.Wgh[(int64_t)k * 2 * H + H + j] : 0.f;
    wc[i] = ok ? a.Wch[(int64_t)k * H + j] : 0.f;
  }

  // seed the state history slot of the first step with h0 (own units, own rows)
  const int t_first = a.reverse ? T - 1 : 0;
  for (int idx = threadIdx.x; idx < a.Bc * g.SL; idx += GC_THREADS) {
    const int b = idx / g.SL, u = idx - b * g.SL;
    const int jj = rank * g.SL + u;
    if (b < nrows && jj < H)
      a.hprev[((int64_t)(b0 + b) * T + t_first) * H + jj] =
          a.h0 ? a.h0[(int64_t)(b0 + b) * H + jj] : 0.f;
  }
  cluster_barrier();

  for (int step = 0; step < T; ++step) {
    const int t = a.reverse ? T - 1 - step : step;
    const bool last = (step == T - 1);
    const int t_next = a.reverse ? t - 1 : t + 1;
    // ---- phase 1: [r,u] = sigmoid(xg + h.Wgh), rh = r*h ----
    gc_load_vec(vec, a.hprev + ((int64_t)b0 * T + t) * H, (int64_t)T * H, nrows, g);
    __syncthreads();
    for (int r0 = 0; r0 < a.Bc; r0 += 4) {
      float ar[4], au[4];
      gc_dot4<CH>(vec, r0, ks, g, wr, ar);
      gc_dot4<CH>(vec, r0, ks, g, wu, au);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ar[i] = reduce8(ar[i]);
        au[i] = reduce8(au[i]);
      }
      // lanes 0..3 of each 8-lane group finish one row each
      if (ks < 4 && unit_ok) {
        const int b = r0 + ks;
        if (b < nrows) {
          const float sr = (ks == 0) ? ar[0] : (ks == 1) ? ar[1] : (ks == 2) ? ar[2] : ar[3];
          const float su = (ks == 0) ? au[0] : (ks == 1) ? au[1] : (ks == 2) ? au[2] : au[3];
          const int64_t row = (int64_t)(b0 + b) * T + t;
          const float rr = sigmoidf_(sr + a.xproj[row * 3 * H + j]);
          const float uu = sigmoidf_(su + a.xproj[row * 3 * H + H + j]);
          a.gates[row * 3 * H + j] = rr;
          a.gates[row * 3 * H + H + j] = uu;
          const float hv = vec[b * g.ROW + (j / g.SL) * g.SLP + (j % g.SL)];
          a.rh[row * H + j] = rr * hv;
        }
      }
    }
    cluster_barrier();
    // ---- phase 2: c = tanh(xc + rh.Wch), h' = u*h + (1-u)*c ----
    gc_load_vec(vec, a.rh + ((int64_t)b0 * T + t) * H, (int64_t)T * H, nrows, g);
    __syncthreads();
    for (int r0 = 0; r0 < a.Bc; r0 += 4) {
      float ac[4];
      gc_dot4<CH>(vec, r0, ks, g, wc, ac);
#pragma unroll
      for (int i = 0; i < 4; ++i) ac[i] = reduce8(ac[i]);
      if (ks < 4 && unit_ok) {
        const int b = r0 + ks;
        if (b < nrows) {
          const float sc = (ks == 0) ? ac[0] : (ks == 1) ? ac[1] : (ks == 2) ? ac[2] : ac[3];
          const int64_t row = (int64_t)(b0 + b) * T + t;
          const float c = tanhf(sc + a.xproj[row * 3 * H + 2 * H + j]);
          a.gates[row * 3 * H + 2 * H + j] = c;
          const float uu = a.gates[row * 3 * H + H + j];
          const float hv = a.hprev[row * H + j];
          const bool live = (a.lengths == nullptr) || (t < a.lengths[b0 + b]);
          float hn = live ? (uu * hv + (1.f - uu) * c) : hv;
          if (a.raw_states) a.raw_states[row * H + j] = live ? hn : 0.f;
          if (a.drop_mask && live) hn *= a.drop_mask[row * H + j];
          a.states[row * H + j] = live ? hn : 0.f;
          if (last) a.final_state[(int64_t)(b0 + b) * H + j] = hn;
          else a.hprev[((int64_t)(b0 + b) * T + t_next) * H + j] = hn;
        }
      }
    }
    cluster_barrier();
  }
}

// ---------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------
struct GcBwdArgs {
  const float* Wgh;
  const float* Wch;
  const int32_t* lengths;
  const float* drop_mask;
  const float* gates;
  const float* hprev;
  const float* dstates;  // or null
  const float* draw;     // or null
  const float* dfinal;   // or null
  float* dxproj;         // [B,T,3H]
  float* dh0;            // or null
  int B, T, H, Bc, reverse;
};

template <int CH>
__global__ void __launch_bounds__(GC_THREADS, 1) gru_seq_bwd_cluster_kernel(GcBwdArgs a) {
  extern __shared__ __align__(16) float gc_smem[];
  GcGeom g;
  g.H = a.H;
  g.SL = (a.H + GC_KS - 1) / GC_KS;
  g.SLP = gc_slice_pad(g.SL);
  g.ROW = GC_KS * g.SLP;
  g.Bc = a.Bc;
  float* vec = gc_smem;                       // [Bc][ROW]  dz_c, then dz_r
  float* vec2 = vec + a.Bc * g.ROW;           // [Bc][ROW]  dz_u
  float* dcarry = vec2 + a.Bc * g.ROW;        // [Bc][GC_MAX_UNITS] grad of h'_t (own units)
  float* dhp = dcarry + a.Bc * GC_MAX_UNITS;  // [Bc][GC_MAX_UNITS]
  const int rank = (int)cluster_rank();
  const int cluster_id = blockIdx.x / GC_CLUSTER;
  const int b0 = cluster_id * a.Bc;
  const int nrows = min(a.Bc, a.B - b0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ks = lane & 7;
  const int unit_local = warp * 4 + (lane >> 3);
  const int i_unit = rank * g.SL + unit_local;  // OUTPUT unit i of this thread
  const bool unit_ok = (unit_local < g.SL) && (i_unit < a.H);
  const int H = a.H, T = a.T;

  // weights: row i of Wch / Wgh restricted to the reduction slice ks (over j)
  float w1[4 * CH], w2r[4 * CH], w2u[4 * CH];
#pragma unroll
  for (int q = 0; q < 4 * CH; ++q) {
    const int jj = ks * g.SL + q;
    const bool ok = unit_ok && q < g.SL && jj < H;
    w1[q] = ok ? a.Wch[(int64_t)i_unit * H + jj] : 0.f;
    w2r[q] = ok ? a.Wgh[(int64_t)i_unit * 2 * H + jj] : 0.f;
    w2u[q] = ok ? a.Wgh[(int64_t)i_unit * 2 * H + H + jj] : 0.f;
  }
  for (int idx = threadIdx.x; idx < a.Bc * GC_MAX_UNITS; idx += GC_THREADS) {
    const int b = idx / GC_MAX_UNITS, u = idx - b * GC_MAX_UNITS;
    const int jj = rank * g.SL + u;
    dcarry[idx] = (a.dfinal && b < nrows && u < g.SL && jj < H) ? a.dfinal[(int64_t)(b0 + b) * H + jj] : 0.f;
  }
  __syncthreads();

  for (int step = T - 1; step >= 0; --step) {
    const int t = a.reverse ? T - 1 - step : step;
    // ---- E1: gate gradients that need no matmul (own units) ----
    for (int idx = threadIdx.x; idx < a.Bc * g.SL; idx += GC_THREADS) {
      const int b = idx / g.SL, u = idx - b * g.SL;
      const int jj = rank * g.SL + u;
      if (b >= nrows || jj >= H) continue;
      const int64_t row = (int64_t)(b0 + b) * T + t;
      const bool live = (a.lengths == nullptr) || (t < a.lengths[b0 + b]);
      float dh = dcarry[b * GC_MAX_UNITS + u];
      if (!live) {
        a.dxproj[row * 3 * H + jj] = 0.f;
        a.dxproj[row * 3 * H + H + jj] = 0.f;
        a.dxproj[row * 3 * H + 2 * H + jj] = 0.f;
        dhp[b * GC_MAX_UNITS + u] = dh;
        continue;
      }
      const float uu = a.gates[row * 3 * H + H + jj], c = a.gates[row * 3 * H + 2 * H + jj];
      const float hv = a.hprev[row * H + jj];
      if (a.dstates) dh += a.dstates[row * H + jj];
      if (a.drop_mask) dh *= a.drop_mask[row * H + jj];
      if (a.draw) dh += a.draw[row * H + jj];
      const float du = dh * (hv - c);
      const float dc = dh * (1.f - uu);
      a.dxproj[row * 3 * H + 2 * H + jj] = dc * (1.f - c * c);
      a.dxproj[row * 3 * H + H + jj] = du * uu * (1.f - uu);
      dhp[b * GC_MAX_UNITS + u] = dh * uu;
    }
    cluster_barrier();
    // ---- G1: drh = dz_c . Wch^T ; dz_r = drh*h*r*(1-r) ; dhp += drh*r ----
    gc_load_vec(vec, a.dxproj + ((int64_t)b0 * T + t) * 3 * H + 2 * H, (int64_t)T * 3 * H, nrows, g);
    gc_load_vec(vec2, a.dxproj + ((int64_t)b0 * T + t) * 3 * H + H, (int64_t)T * 3 * H, nrows, g);
    __syncthreads();
    for (int r0 = 0; r0 < a.Bc; r0 += 4) {
      float acc[4];
      gc_dot4<CH>(vec, r0, ks, g, w1, acc);
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = reduce8(acc[q]);
      if (ks < 4 && unit_ok) {
        const int b = r0 + ks;
        if (b < nrows) {
          const bool live = (a.lengths == nullptr) || (t < a.lengths[b0 + b]);
          if (live) {
            const float drh = (ks == 0) ? acc[0] : (ks == 1) ? acc[1] : (ks == 2) ? acc[2] : acc[3];
            const int64_t row = (int64_t)(b0 + b) * T + t;
            const float rr = a.gates[row * 3 * H + i_unit];
            const float hv = a.hprev[row * H + i_unit];
            a.dxproj[row * 3 * H + i_unit] = drh * hv * rr * (1.f - rr);
            dhp[b * GC_MAX_UNITS + unit_local] += drh * rr;
          }
        }
      }
    }
    cluster_barrier();
    // ---- G2: dcarry = dhp + [dz_r, dz_u] . Wgh^T ----
    gc_load_vec(vec, a.dxproj + ((int64_t)b0 * T + t) * 3 * H, (int64_t)T * 3 * H, nrows, g);
    __syncthreads();
    for (int r0 = 0; r0 < a.Bc; r0 += 4) {
      float accr[4], accu[4];
      gc_dot4<CH>(vec, r0, ks, g, w2r, accr);
      gc_dot4<CH>(vec2, r0, ks, g, w2u, accu);
#pragma unroll
      for (int q = 0; q < 4; ++q) accr[q] = reduce8(accr[q] + accu[q]);
      if (ks < 4 && unit_ok) {
        const int b = r0 + ks;
        if (b < nrows) {
          const float s = (ks == 0) ? accr[0] : (ks == 1) ? accr[1] : (ks == 2) ? accr[2] : accr[3];
          dcarry[b * GC_MAX_UNITS + unit_local] = dhp[b * GC_MAX_UNITS + unit_local] + s;
        }
      }
    }
    __syncthreads();  // dcarry/dhp/vec are CTA-private: no cluster barrier needed here
  }
  if (a.dh0) {
    for (int idx = threadIdx.x; idx < a.Bc * g.SL; idx += GC_THREADS) {
      const int b = idx / g.SL, u = idx - b * g.SL;
      const int jj = rank * g.SL + u;
      if (b < nrows && jj < H) a.dh0[(int64_t)(b0 + b) * H + jj] = dcarry[b * GC_MAX_UNITS + u];
    }
  }
}

}  // namespace nm
