// Persistent GRU sequence kernels on the 5th-generation tensor cores (sm_100a).
//
// One launch runs the whole time loop.  A cluster of 8 CTAs owns up to 32 sentences; CTA j
// owns hidden units [j*UPC, (j+1)*UPC), UPC = ceil(H/8) rounded to 4.  The recurrent
// weights of those units stay ON CHIP for the whole sequence as the A operand of
// tcgen05.mma, *in tensor memory*: TMEM lane m = (gate, unit) row, TMEM column k = input
// unit, written once with tcgen05.st (TF32, round-to-nearest).  The B operand is the state
// matrix h [32 sentences x KP], K-major / no swizzle in shared memory, and the fp32
// accumulators D[128 x 32] live in TMEM next to A.  A step is two dependent products
// (TF-1.12 GRUCell: the candidate needs r*h of ALL units):
//
//     phase 1:  D1 = A . h^T        -> r, u for the CTA's units;  r*h slice -> all 8 CTAs
//     phase 2:  D2 = A . (r*h)^T    -> c, h' for the CTA's units; h' slice  -> all 8 CTAs
//
// Slices travel as 16-byte asynchronous stores (st.async ... mbarrier::complete_tx) straight into
// the peers' B tiles: the data carries its own completion signal, so a CTA starts a phase when
// its own transaction barrier has counted the bytes of all 8 slices - no fences, no arrive
// instructions and no cluster-wide barrier in the loop.  The two B tiles (h and r*h) alternate,
// and a peer can only overwrite a tile after it has received this CTA's slice of the *other*
// tile, which this CTA sends only after its MMA on the first one has completed.
//
// Measured on B200: one tcgen05.mma costs ~100 cycles of ISSUE latency from one
// thread whatever its N, so the K loop is split over several issuing warps with separate
// accumulators, the forward pass uses FP16 operands (same 11-bit significand as TF32, half the
// K steps and half the exchange bytes; fp32 accumulation), and all element-wise work runs in
// an "item space" that spreads (sentence, unit) pairs evenly over the 256 threads, with the
// recurrent state of an item held in a register of its thread for the whole sequence.
#pragma once
#include "common.cuh"
#include "gru_cluster.cuh"
#include "tc_ptx.cuh"

namespace nm {

constexpr int GT_CLUSTER = 8;
constexpr int GT_THREADS = 256;     // 2 threads per TMEM lane: sentence columns [0,16) and [16,32)
constexpr int GT_NB = 32;           // sentences per cluster = MMA N
constexpr int GT_MAX_UPC = 40;      // 3 * UPC <= 128 rows
constexpr int GT_NACC = 4;          // accumulators = MMA-issuing warps (the issue path is per-warp
                                    // latency bound: ~100 cycles per tcgen05.mma from one thread)
constexpr int GT_TMEM_COLS = 512;
constexpr int GT_HS = GT_NB + 1;    // pitch of the [unit][sentence] fp32 arrays
constexpr int GT_MIN_SMEM = 120 * 1024;  // > half an SM: one CTA per SM (each takes all of TMEM)

// ESZ = operand element size: 4 = TF32 (backward: gradients need the fp32 exponent range),
// 2 = FP16 (forward: h, r*h in (-1,1) and the weights keep the same 11-bit significand as TF32
// in half the bytes, halving both the MMA count and the exchange traffic).
template <int ESZ>
struct GtGeom {
  static constexpr int CK = 16 / ESZ;          // K elements per 16-byte core-matrix row
  int UPC, KP, SBO, tile_bytes, nk;
  __host__ __device__ explicit GtGeom(int H) {
    UPC = ((H + GT_CLUSTER - 1) / GT_CLUSTER + 7) / 8 * 8;
    KP = UPC * GT_CLUSTER;                     // a multiple of 64
    SBO = KP * ESZ * 8;                        // 8 sentences x KP elements
    tile_bytes = (GT_NB / 8) * SBO;
    nk = KP * ESZ / 32;                        // 32 bytes of K per instruction; a multiple of 4
  }
};

// dynamic shared memory carve-up (bytes from the 128-byte aligned base)
struct GtSmem {
  int tile0, tile1, tile2, stage, hs, us, lens, bars, tmem_slot, total;
  __host__ __device__ GtSmem(int tile_bytes, int ntiles) {
    tile0 = 0;
    tile1 = tile_bytes;
    tile2 = 2 * tile_bytes;
    stage = ntiles * tile_bytes;
    hs = stage + GT_NB * GT_MAX_UPC * 4;
    us = hs + 3 * GT_MAX_UPC * GT_HS * 4;    // hs: up to 3 [unit][sentence] arrays
    lens = us + GT_MAX_UPC * GT_HS * 4;
    bars = lens + GT_NB * 4;
    tmem_slot = bars + 8 * 8;
    total = tmem_slot + 16 + 128;
    if (total < GT_MIN_SMEM) total = GT_MIN_SMEM;
  }
};

// Sentence s of the cluster's slice sits in tile row s.
__device__ __forceinline__ int gt_row_of(int s) { return s; }

template <int ESZ>
__device__ __forceinline__ void gt_stage_put(void* stage, int n, int i, float v) {
  if (ESZ == 4)
    reinterpret_cast<uint32_t*>(stage)[n * GT_MAX_UPC + i] = to_tf32(v);
  else
    reinterpret_cast<__half*>(stage)[n * GT_MAX_UPC + i] = __float2half_rn(v);
}

// Send the staged [tile row][UPC] slice of sentences [0, nb) to the B tile at `tile_off` of all 8
// CTAs with asynchronous 16-byte stores that count their bytes on the destination CTA's barrier
// at `bar_off` (no fences, no arrive: the data carries its own completion signal).
template <int ESZ>
__device__ __forceinline__ void gt_send_slice(const void* stage, int nb, int UPC, int u0, int SBO,
                                              uint32_t smem_base, uint32_t tile_off, uint32_t bar_off,
                                              int tid) {
  constexpr int CK = 16 / ESZ;
  const int Q = UPC / CK;
  const int pairs = nb * Q * GT_CLUSTER;      // (16-byte chunk, destination CTA): spread over all threads
  for (int p = tid; p < pairs; p += GT_THREADS) {
    const int idx = p >> 3;
    const uint32_t dst = (uint32_t)(p & 7);
    const int s = idx / Q, q = idx - s * Q;
    const int n = gt_row_of(s);
    const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(stage) +
                                                    (n * GT_MAX_UPC + CK * q) * ESZ);
    const uint32_t off = tile_off + (uint32_t)((n >> 3) * SBO + ((u0 / CK + q) << 7) + ((n & 7) << 4));
    const uint32_t remote = mapa_u32(smem_base, dst);
    st_async_v4(remote + off, remote + bar_off, v.x, v.y, v.z, v.w);
  }
}

// Warps 0..NACC-1 each issue a share of the K loop of D = A[tmem] . B[tile]^T (N sentence rows)
// into their own accumulator (N columns each; the epilogue adds them) and commit to `bar`
// (count NACC).  Called by lane 0 of those warps.
template <int ESZ, int N, int NACC>
__device__ __forceinline__ void gt_issue_mma(int warp, uint32_t tmem_d, uint32_t tmem_a,
                                             uint32_t tile_addr, int nk, int SBO, uint32_t bar) {
  // instruction descriptor: c=F32 [4,6)=1, a/b format [7,10)/[10,13) (TF32 = 2, F16 = 0), K-major
  // both, N>>3 at [17,23), M>>4 at [24,29)
  constexpr uint32_t FMT = ESZ == 4 ? 2u : 0u;
  const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(N >> 3) << 17) |
                         ((uint32_t)(128 >> 4) << 24);
  tcgen05_fence_after();
  // K-major, no swizzle: core matrix = 8 sentences x 16 bytes; LBO = 128 B between the two K
  // cores of one instruction, SBO between groups of 8 sentences; +256 B and +8 TMEM columns per step
  const int per = (nk + NACC - 1) / NACC;
  const int k0 = warp * per, k1 = min(nk, k0 + per);
  const uint64_t db = smem_desc(tile_addr, 128, (uint32_t)SBO, 0);
  const uint32_t d = tmem_d + warp * N;
#pragma unroll 5
  for (int k = k0; k < k1; ++k) {
    if (ESZ == 4)
      umma_tf32_ts(d, tmem_a + k * 8, db + (uint64_t)(k * 16), idesc, k > k0 ? 1u : 0u);
    else
      umma_f16_ts(d, tmem_a + k * 8, db + (uint64_t)(k * 16), idesc, k > k0 ? 1u : 0u);
  }
  umma_commit(bar);
}

// acc[j] = sum over accumulators of D[lane][16*half + j]
__device__ __forceinline__ void gt_load_acc(uint32_t tmem_d, uint32_t lane_base, int half, float (&acc)[16]) {
  float part[GT_NACC][16];
#pragma unroll
  for (int k = 0; k < GT_NACC; ++k) tmem_ld16_nowait(tmem_d + lane_base + k * GT_NB + 16 * half, part[k]);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float v = part[0][j];
#pragma unroll
    for (int k = 1; k < GT_NACC; ++k) v += part[k][j];
    acc[j] = v;
  }
}

// Load the CTA's weight rows into TMEM columns [0, KP*ESZ/4): lane m = gate*UPC + i.
//   forward : row (g,i), K index k = W[k][g*H + unit]  (Wgh for g<2, Wch for g=2): h . W
//   backward: row (g,i), K index k = W[unit][...k]      : dz . W^T  (see gru_seq_bwd_tc_kernel)
template <int ESZ, bool BACKWARD>
__device__ __forceinline__ void gt_load_weights(uint32_t tmem_base, const float* __restrict__ Wgh,
                                                const float* __restrict__ Wch, int H, int UPC, int KP,
                                                int u0, int tid) {
  const int m = tid & 127, half = tid >> 7;
  const int g = m / UPC, i = m - g * UPC, unit = u0 + i;
  const bool valid = g < 3 && unit < H;
  const uint32_t lane_base = (uint32_t)(m & ~31) << 16;
  constexpr int EPC = 4 / ESZ;                 // K elements per 32-bit TMEM column
  const int ncols = KP / EPC;                  // a multiple of 32
  for (int cc = half; cc < ncols / 32; cc += 2) {
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float w[EPC];
#pragma unroll
      for (int e = 0; e < EPC; ++e) {
        const int k = (cc * 32 + j) * EPC + e;
        w[e] = 0.f;
        if (valid && k < H) {
          if (!BACKWARD) {
            w[e] = g < 2 ? Wgh[(int64_t)k * 2 * H + g * H + unit] : Wch[(int64_t)k * H + unit];
          } else {
            // g = 0: d(rh) = dz_c . Wch^T -> Wch[unit][k]; g = 1: dz_u . Wgh[:, H:]^T; g = 2: dz_r . Wgh[:, :H]^T
            w[e] = g == 0 ? Wch[(int64_t)unit * H + k] : Wgh[(int64_t)unit * 2 * H + (g == 1 ? H : 0) + k];
          }
        }
      }
      if (ESZ == 4) {
        v[j] = to_tf32(w[0]);
      } else {
        const __half2 h2 = __floats2half2_rn(w[0], w[EPC - 1]);   // low half = lower K index
        v[j] = *reinterpret_cast<const uint32_t*>(&h2);
      }
    }
    tmem_st32(tmem_base + lane_base + cc * 32, v);
  }
  tmem_st_wait();
}

struct GtFwdArgs {
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
  long long* prof;      // 8 cycle counters (CTA 0, thread 0) or null
};

constexpr int GT_ITEMS = GT_NB * GT_MAX_UPC / GT_THREADS;   // (sentence, unit) pairs per thread: 5

// Dump this thread's 16 accumulator columns into the [row][sentence] transposition buffer.
// tcgen05.ld is warp-collective (.sync.aligned): every thread loads, `store` only gates the writes.
__device__ __forceinline__ void gt_dump_acc(float* __restrict__ P, uint32_t tmem_d, uint32_t lane_base,
                                            int m, int half, bool store) {
  float acc[16];
  gt_load_acc(tmem_d, lane_base, half, acc);
  if (store) {
#pragma unroll
    for (int j = 0; j < 16; ++j) P[m * GT_HS + 16 * half + j] = acc[j];
  }
}

// Two independent sequences can share one launch (the fw and bw directions of a bidirectional layer:
// clusters [0, split) work on a0, the rest on a1), so that both recurrences are co-resident instead
// of running back to back on half-empty SMs.  A single sequence passes split = number of clusters.
__global__ void __launch_bounds__(GT_THREADS, 1) gru_seq_fwd_tc_kernel(const GtFwdArgs a0, const GtFwdArgs a1,
                                                                      const int split) {
  const bool second_seq = (int)(blockIdx.x / GT_CLUSTER) >= split;
  const GtFwdArgs a = second_seq ? a1 : a0;
  constexpr int ESZ = 2;   // fp16 operands
  extern __shared__ uint8_t gt_smem_raw[];
  const GtGeom<ESZ> geo(a.H);
  const GtSmem lay(geo.tile_bytes, 2);
  const uint32_t raw_addr = smem_u32(gt_smem_raw);
  const uint32_t pad = (128u - (raw_addr & 127u)) & 127u;   // identical in every CTA of the launch
  uint8_t* smem = gt_smem_raw + pad;
  const uint32_t smem_base = raw_addr + pad;
  void* stage = smem + lay.stage;
  float* __restrict__ P = reinterpret_cast<float*>(smem + lay.hs);   // [128 rows][GT_HS]
  int* lens = reinterpret_cast<int*>(smem + lay.lens);
  const uint32_t bar_h = smem_base + lay.bars, bar_rh = bar_h + 8, bar_mma = bar_h + 16;
  const uint32_t tmem_slot = smem_base + lay.tmem_slot;

  const int tid = threadIdx.x;
  const int m = tid & 127, half = tid >> 7;
  const int rank = (int)cluster_rank();
  const int cluster_id = (int)(blockIdx.x / GT_CLUSTER) - (second_seq ? split : 0);
  const int b0 = cluster_id * a.Bc;
  const int nb = min(a.Bc, a.B - b0);
  const int H = a.H, T = a.T, UPC = geo.UPC, KP = geo.KP, SBO = geo.SBO;
  const int u0 = rank * UPC;
  // outputs / inputs of this cluster's sentences, addressed with 32-bit offsets
  float* __restrict__ gates_out = a.gates + (int64_t)b0 * T * 3 * H;
  float* __restrict__ rh_out = a.rh + (int64_t)b0 * T * H;
  float* __restrict__ states_out = a.states + (int64_t)b0 * T * H;
  float* __restrict__ raw_out = a.raw_states ? a.raw_states + (int64_t)b0 * T * H : nullptr;
  float* __restrict__ hprev_out = a.hprev + (int64_t)b0 * T * H;
  float* __restrict__ final_out = a.final_state + (int64_t)b0 * H;
  const float* __restrict__ xproj = a.xproj + (int64_t)b0 * T * 3 * H;
  const float* __restrict__ dropm = a.drop_mask ? a.drop_mask + (int64_t)b0 * T * H : nullptr;

  if (tid == 0) {
    mbar_init(bar_h, 1);    // one local arrive.expect_tx per phase; the peers' st.async supply the bytes
    mbar_init(bar_rh, 1);
    mbar_init(bar_mma, GT_NACC);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < GT_NB) lens[tid] = (a.lengths && tid < nb) ? a.lengths[b0 + tid] : T;
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(GT_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + lay.tmem_slot);
  const uint32_t tmem_a = tmem_base;
  const uint32_t tmem_d = tmem_base + KP * ESZ / 4;
  const int warp = tid >> 5, nk = geo.nk;
  const bool issuer = warp < GT_NACC && (tid & 31) == 0;
  const uint32_t lane_base = (uint32_t)(m & ~31) << 16;
  const bool dump_row = m < 3 * UPC;

  gt_load_weights<ESZ, false>(tmem_base, a.Wgh, a.Wch, H, UPC, KP, u0, tid);

  // Element-wise work runs in "item space": item = (sentence s, own unit i), idx = s*UPC + i,
  // thread tid owns idx = tid + it*256.  The mapping is fixed over time, so the state h and the
  // update gate u of an item live in registers of its thread for the whole sequence.
  const int n_items = nb * UPC;
  int it_s[GT_ITEMS], it_i[GT_ITEMS];
  float h_reg[GT_ITEMS], u_reg[GT_ITEMS], xr[GT_ITEMS], xu[GT_ITEMS], xc[GT_ITEMS];
  const int t_first = a.reverse ? T - 1 : 0;
#pragma unroll
  for (int it = 0; it < GT_ITEMS; ++it) {
    const int idx = tid + it * GT_THREADS;
    const int s = idx / UPC;
    it_s[it] = idx < n_items ? s : -1;
    it_i[it] = idx - s * UPC;
    h_reg[it] = 0.f;
    u_reg[it] = 0.f;
    xr[it] = xu[it] = xc[it] = 0.f;
    if (it_s[it] >= 0) {
      const int unit = u0 + it_i[it];
      if (unit < H) {
        h_reg[it] = a.h0 ? a.h0[(int64_t)(b0 + s) * H + unit] : 0.f;
        hprev_out[(s * T + t_first) * H + unit] = h_reg[it];
        const float* xp = xproj + (s * T + t_first) * 3 * H + unit;
        xr[it] = xp[0];
        xu[it] = xp[H];
        xc[it] = xp[2 * H];
      }
      gt_stage_put<ESZ>(stage, gt_row_of(s), it_i[it], h_reg[it]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  // barriers of every CTA are initialised before anyone signals them
  cluster_barrier();
  const uint32_t phase_bytes = (uint32_t)(nb * KP * ESZ);   // 8 senders x nb sentences x UPC elements
  if (tid == 0) {
    mbar_expect_tx(bar_h, phase_bytes);
    mbar_expect_tx(bar_rh, phase_bytes);
  }
  gt_send_slice<ESZ>(stage, nb, UPC, u0, SBO, smem_base, lay.tile0, (uint32_t)lay.bars, tid);

  uint32_t ph_h = 0, ph_rh = 0, ph_mma = 0;
  const bool prof = a.prof != nullptr && blockIdx.x == 0 && tid == 0;
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tp = prof ? clock64() : 0;
#define GT_PROF(slot)                 \
  if (prof) {                         \
    const long long now = clock64();  \
    pc[slot] += now - tp;             \
    tp = now;                         \
  }
  for (int step = 0; step < T; ++step) {
    const int t = a.reverse ? T - 1 - step : step;
    const bool last = step == T - 1;
    const int t_next = a.reverse ? t - 1 : t + 1;
    // ---------------- phase 1: r, u ----------------
    mbar_wait_cluster(bar_h, ph_h);
    ph_h ^= 1;
    if (tid == 0 && !last) mbar_expect_tx(bar_h, phase_bytes);   // arm the next phase
    if (issuer) gt_issue_mma<ESZ, GT_NB, GT_NACC>(warp, tmem_d, tmem_a, smem_base + lay.tile0, nk, SBO, bar_mma);
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tcgen05_fence_after();
    gt_dump_acc(P, tmem_d, lane_base, m, half, dump_row);
    tcgen05_fence_before();
    __syncthreads();
    GT_PROF(0)
#pragma unroll
    for (int it = 0; it < GT_ITEMS; ++it) {
      const int s = it_s[it], i = it_i[it];
      if (s < 0) continue;
      const int n = gt_row_of(s), unit = u0 + i;
      float rhv = 0.f;
      if (unit < H) {
        const float r = fast_sigmoid(P[i * GT_HS + n] + xr[it]);
        const float u = fast_sigmoid(P[(UPC + i) * GT_HS + n] + xu[it]);
        u_reg[it] = u;
        rhv = r * h_reg[it];
        const int row = s * T + t;
        gates_out[row * 3 * H + unit] = r;
        gates_out[row * 3 * H + H + unit] = u;
        rh_out[row * H + unit] = rhv;
      }
      gt_stage_put<ESZ>(stage, n, i, rhv);
    }
    GT_PROF(1)
    __syncthreads();
    GT_PROF(2)
    gt_send_slice<ESZ>(stage, nb, UPC, u0, SBO, smem_base, lay.tile1, (uint32_t)lay.bars + 8, tid);
    GT_PROF(3)

    // ---------------- phase 2: c, h' ----------------
    mbar_wait_cluster(bar_rh, ph_rh);
    ph_rh ^= 1;
    if (tid == 0 && !last) mbar_expect_tx(bar_rh, phase_bytes);
    if (issuer) gt_issue_mma<ESZ, GT_NB, GT_NACC>(warp, tmem_d, tmem_a, smem_base + lay.tile1, nk, SBO, bar_mma);
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tcgen05_fence_after();
    gt_dump_acc(P, tmem_d, lane_base, m, half, dump_row && m >= 2 * UPC);
    tcgen05_fence_before();
    __syncthreads();
    GT_PROF(4)
#pragma unroll
    for (int it = 0; it < GT_ITEMS; ++it) {
      const int s = it_s[it], i = it_i[it];
      if (s < 0) continue;
      const int n = gt_row_of(s), unit = u0 + i;
      float hn = 0.f;
      if (unit < H) {
        const float c = fast_tanh_exp(P[(2 * UPC + i) * GT_HS + n] + xc[it]);
        const float u = u_reg[it], h = h_reg[it];
        const bool live = t < lens[s];
        const int row = s * T + t;
        hn = live ? (u * h + (1.f - u) * c) : h;
        gates_out[row * 3 * H + 2 * H + unit] = c;
        if (raw_out) raw_out[row * H + unit] = live ? hn : 0.f;
        if (dropm != nullptr && live) hn *= dropm[row * H + unit];
        states_out[row * H + unit] = live ? hn : 0.f;
        if (last)
          final_out[s * H + unit] = hn;
        else
          hprev_out[(s * T + t_next) * H + unit] = hn;
        h_reg[it] = hn;
      }
      gt_stage_put<ESZ>(stage, n, i, hn);
    }
    // next step's x-projection: unconditional loads from always-valid addresses (a conditional
    // assignment would turn into a select that waits for the load)
#pragma unroll
    for (int it = 0; it < GT_ITEMS; ++it) {
      const int s = it_s[it] < 0 ? 0 : it_s[it];
      const int unit = min(u0 + it_i[it], H - 1);
      const float* xp = xproj + (s * T + (last ? t : t_next)) * 3 * H + unit;
      xr[it] = xp[0];
      xu[it] = xp[H];
      xc[it] = xp[2 * H];
    }
    GT_PROF(5)
    __syncthreads();
    GT_PROF(6)
    if (!last) gt_send_slice<ESZ>(stage, nb, UPC, u0, SBO, smem_base, lay.tile0, (uint32_t)lay.bars, tid);
    GT_PROF(7)
  }
  if (prof)
    for (int k = 0; k < 8; ++k) a.prof[k] = pc[k];
#undef GT_PROF

  // nobody leaves while a peer may still write into its shared memory
  tcgen05_fence_before();
  cluster_barrier();
  if (tid < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(GT_TMEM_COLS)
                 : "memory");
}

// ---------------------------------------------------------------------------------------
// Backward through time.  TF32 operands: the B operand is a gradient and needs the fp32 exponent
// range.  TMEM lane rows of A (K index = the unit the gradient comes FROM):
//   rows [0,UPC)      Wch[unit][k]      : d(rh)   = dz_c . Wch^T
//   rows [UPC,2UPC)   Wgh[unit][H + k]  : dh_prev += dz_u . Wgh[:, H:]^T
//   rows [2UPC,3UPC)  Wgh[unit][k]      : dh_prev += dz_r . Wgh[:, :H]^T
// Phase A multiplies A by a 64-row tile [dz_c sentences | dz_u sentences] in ONE K loop (the cost of
// an MMA here is its issue, not its N), phase B by the dz_r tile.
// ---------------------------------------------------------------------------------------
constexpr int GT_BWD_NACC = 3;     // 3 x 64 accumulator columns + 320 weight columns = 512

struct GtBwdArgs {
  const float* Wgh;
  const float* Wch;
  const int32_t* lengths;
  const float* drop_mask;
  const float* gates;    // [B,T,3H] (r,u,c)
  const float* hprev;    // [B,T,H]
  const float* dstates;  // or null
  const float* draw;     // or null
  const float* dfinal;   // or null
  float* dxproj;         // [B,T,3H] = (dz_r, dz_u, dz_c)
  float* dh0;            // or null
  int B, T, H, Bc, reverse;
};

__global__ void __launch_bounds__(GT_THREADS, 1) gru_seq_bwd_tc_kernel(const GtBwdArgs a0, const GtBwdArgs a1,
                                                                      const int split) {
  const bool second_seq = (int)(blockIdx.x / GT_CLUSTER) >= split;
  const GtBwdArgs a = second_seq ? a1 : a0;
  constexpr int ESZ = 4;
  extern __shared__ uint8_t gt_smem_raw[];
  const GtGeom<ESZ> geo(a.H);
  const GtSmem lay(geo.tile_bytes, 3);     // tiles 0,1 = the 64-row phase-A tile, tile 2 = dz_r
  const uint32_t raw_addr = smem_u32(gt_smem_raw);
  const uint32_t pad = (128u - (raw_addr & 127u)) & 127u;
  uint8_t* smem = gt_smem_raw + pad;
  const uint32_t smem_base = raw_addr + pad;
  uint8_t* stage = smem + lay.stage;                                   // dz_c / dz_r slices
  uint8_t* stage2 = smem + lay.us;                                     // dz_u slice
  float* __restrict__ P = reinterpret_cast<float*>(smem + lay.hs);     // [128 rows][GT_HS]
  int* lens = reinterpret_cast<int*>(smem + lay.lens);
  const uint32_t bar_a = smem_base + lay.bars, bar_b = bar_a + 8, bar_mma = bar_a + 16;
  const uint32_t tmem_slot = smem_base + lay.tmem_slot;

  const int tid = threadIdx.x;
  const int m = tid & 127, half = tid >> 7;
  const int rank = (int)cluster_rank();
  const int cluster_id = (int)(blockIdx.x / GT_CLUSTER) - (second_seq ? split : 0);
  const int b0 = cluster_id * a.Bc;
  const int nb = min(a.Bc, a.B - b0);
  const int H = a.H, T = a.T, UPC = geo.UPC, KP = geo.KP, SBO = geo.SBO;
  const int u0 = rank * UPC;
  const float* __restrict__ gates = a.gates + (int64_t)b0 * T * 3 * H;
  const float* __restrict__ hprev = a.hprev + (int64_t)b0 * T * H;
  const float* __restrict__ dstates = a.dstates ? a.dstates + (int64_t)b0 * T * H : nullptr;
  const float* __restrict__ draw = a.draw ? a.draw + (int64_t)b0 * T * H : nullptr;
  const float* __restrict__ dropm = a.drop_mask ? a.drop_mask + (int64_t)b0 * T * H : nullptr;
  float* __restrict__ dxproj = a.dxproj + (int64_t)b0 * T * 3 * H;

  if (tid == 0) {
    mbar_init(bar_a, 1);
    mbar_init(bar_b, 1);
    mbar_init(bar_mma, GT_BWD_NACC);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < GT_NB) lens[tid] = (a.lengths && tid < nb) ? a.lengths[b0 + tid] : T;
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(GT_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + lay.tmem_slot);
  const uint32_t tmem_a = tmem_base;
  const uint32_t tmem_d = tmem_base + KP;       // 3 accumulators x 64 columns
  const int warp = tid >> 5, nk = geo.nk;
  const bool issuer = warp < GT_BWD_NACC && (tid & 31) == 0;
  const uint32_t lane_base = (uint32_t)(m & ~31) << 16;

  gt_load_weights<ESZ, true>(tmem_base, a.Wgh, a.Wch, H, UPC, KP, u0, tid);

  // item space as in the forward kernel; the carried gradient of an item stays in a register
  const int n_items = nb * UPC;
  int it_s[GT_ITEMS], it_i[GT_ITEMS];
  float dh[GT_ITEMS], dhp[GT_ITEMS];
  float g_r[GT_ITEMS], g_u[GT_ITEMS], g_c[GT_ITEMS], h_p[GT_ITEMS], d_s[GT_ITEMS], d_w[GT_ITEMS], d_m[GT_ITEMS];
  const int t_last = a.reverse ? 0 : T - 1;      // the step processed first
#pragma unroll
  for (int it = 0; it < GT_ITEMS; ++it) {
    const int idx = tid + it * GT_THREADS;
    const int s = idx / UPC;
    it_s[it] = idx < n_items ? s : -1;
    it_i[it] = idx - s * UPC;
    const int unit = u0 + it_i[it];
    dh[it] = (it_s[it] >= 0 && unit < H && a.dfinal) ? a.dfinal[(int64_t)(b0 + s) * H + unit] : 0.f;
    dhp[it] = 0.f;
  }
  auto prefetch = [&](int t) {
#pragma unroll
    for (int it = 0; it < GT_ITEMS; ++it) {
      const int s = it_s[it] < 0 ? 0 : it_s[it];
      const int unit = min(u0 + it_i[it], H - 1);
      const int row = s * T + t;
      const float* gp = gates + row * 3 * H + unit;
      g_r[it] = gp[0];
      g_u[it] = gp[H];
      g_c[it] = gp[2 * H];
      h_p[it] = hprev[row * H + unit];
      d_s[it] = dstates ? dstates[row * H + unit] : 0.f;
      d_w[it] = draw ? draw[row * H + unit] : 0.f;
      d_m[it] = dropm ? dropm[row * H + unit] : 1.f;
    }
  };
  prefetch(t_last);
  tcgen05_fence_before();
  __syncthreads();
  cluster_barrier();
  const uint32_t bytes_a = (uint32_t)(2 * nb * KP * ESZ), bytes_b = (uint32_t)(nb * KP * ESZ);
  if (tid == 0) {
    mbar_expect_tx(bar_a, bytes_a);
    mbar_expect_tx(bar_b, bytes_b);
  }

  uint32_t ph_a = 0, ph_b = 0, ph_mma = 0;
  for (int step = T - 1; step >= 0; --step) {
    const int t = a.reverse ? T - 1 - step : step;
    const bool last = step == 0;
    // ---------------- element-wise part 1: dz_c, dz_u ----------------
#pragma unroll
    for (int it = 0; it < GT_ITEMS; ++it) {
      const int s = it_s[it], i = it_i[it];
      if (s < 0) continue;
      const int unit = u0 + i;
      float dzc = 0.f, dzu = 0.f;
      if (unit < H) {
        const bool live = t < lens[s];
        const int row = s * T + t;
        if (live) {
          float g = dh[it] + d_s[it];
          g = g * d_m[it] + d_w[it];        // through the dropout mask, plus the raw-output gradient
          const float u = g_u[it], c = g_c[it];
          dzc = g * (1.f - u) * (1.f - c * c);
          dzu = g * (h_p[it] - c) * u * (1.f - u);
          dhp[it] = g * u;
        } else {
          dhp[it] = dh[it];
          dxproj[row * 3 * H + unit] = 0.f;   // dz_r of a finished sentence
        }
        dxproj[row * 3 * H + H + unit] = dzu;
        dxproj[row * 3 * H + 2 * H + unit] = dzc;
      }
      gt_stage_put<ESZ>(stage, s, i, dzc);
      gt_stage_put<ESZ>(stage2, s, i, dzu);
    }
    __syncthreads();
    gt_send_slice<ESZ>(stage, nb, UPC, u0, SBO, smem_base, lay.tile0, (uint32_t)lay.bars, tid);
    gt_send_slice<ESZ>(stage2, nb, UPC, u0, SBO, smem_base, lay.tile1, (uint32_t)lay.bars, tid);

    // ---------------- phase A: d(rh) and the u part of dh_prev ----------------
    mbar_wait_cluster(bar_a, ph_a);
    ph_a ^= 1;
    if (tid == 0 && !last) mbar_expect_tx(bar_a, bytes_a);
    if (issuer) gt_issue_mma<ESZ, 2 * GT_NB, GT_BWD_NACC>(warp, tmem_d, tmem_a, smem_base + lay.tile0, nk, SBO, bar_mma);
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tcgen05_fence_after();
    {
      // rows [0,UPC) need columns [0,32) (dz_c sentences), rows [UPC,2UPC) columns [32,64)
      float lo[16], hi[16], part[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) lo[j] = hi[j] = 0.f;
#pragma unroll
      for (int k = 0; k < GT_BWD_NACC; ++k) {
        tmem_ld16_nowait(tmem_d + lane_base + k * 2 * GT_NB + 16 * half, part);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) lo[j] += part[j];
        tmem_ld16_nowait(tmem_d + lane_base + k * 2 * GT_NB + GT_NB + 16 * half, part);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) hi[j] += part[j];
      }
      if (m < 2 * UPC) {
#pragma unroll
        for (int j = 0; j < 16; ++j) P[m * GT_HS + 16 * half + j] = m < UPC ? lo[j] : hi[j];
      }
    }
    tcgen05_fence_before();
    __syncthreads();
#pragma unroll
    for (int it = 0; it < GT_ITEMS; ++it) {
      const int s = it_s[it], i = it_i[it];
      if (s < 0) continue;
      const int unit = u0 + i;
      float dzr = 0.f;
      if (unit < H) {
        const float drh = P[i * GT_HS + s];
        const float r = g_r[it];
        dzr = drh * h_p[it] * r * (1.f - r);          // zero for finished sentences: their dz_c is zero
        dhp[it] += drh * r + P[(UPC + i) * GT_HS + s];
        if (t < lens[s]) dxproj[(s * T + t) * 3 * H + unit] = dzr;
      }
      gt_stage_put<ESZ>(stage, s, i, dzr);
    }
    __syncthreads();
    gt_send_slice<ESZ>(stage, nb, UPC, u0, SBO, smem_base, lay.tile2, (uint32_t)lay.bars + 8, tid);
    if (!last) prefetch(a.reverse ? t + 1 : t - 1);

    // ---------------- phase B: the r part of dh_prev ----------------
    mbar_wait_cluster(bar_b, ph_b);
    ph_b ^= 1;
    if (tid == 0 && !last) mbar_expect_tx(bar_b, bytes_b);
    if (issuer) gt_issue_mma<ESZ, GT_NB, GT_BWD_NACC>(warp, tmem_d, tmem_a, smem_base + lay.tile2, nk, SBO, bar_mma);
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tcgen05_fence_after();
    {
      float acc[16], part[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
      for (int k = 0; k < GT_BWD_NACC; ++k) {
        tmem_ld16_nowait(tmem_d + lane_base + k * GT_NB + 16 * half, part);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += part[j];
      }
      if (m >= 2 * UPC && m < 3 * UPC) {
#pragma unroll
        for (int j = 0; j < 16; ++j) P[m * GT_HS + 16 * half + j] = acc[j];
      }
    }
    tcgen05_fence_before();
    __syncthreads();
#pragma unroll
    for (int it = 0; it < GT_ITEMS; ++it) {
      const int s = it_s[it], i = it_i[it];
      if (s < 0 || u0 + i >= H) continue;
      dh[it] = dhp[it] + P[(2 * UPC + i) * GT_HS + s];
    }
    // (the next element-wise pass overwrites `stage` only after this __syncthreads-separated read)
  }
  if (a.dh0) {
#pragma unroll
    for (int it = 0; it < GT_ITEMS; ++it) {
      const int s = it_s[it], unit = u0 + it_i[it];
      if (s >= 0 && unit < H) a.dh0[(int64_t)(b0 + s) * H + unit] = dh[it];
    }
  }
  tcgen05_fence_before();
  cluster_barrier();
  if (tid < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(GT_TMEM_COLS)
                 : "memory");
}

}  // namespace nm
