// Microbenchmark: cycles for a chain of small tcgen05.mma (M=128) with A in TMEM or smem,
// tf32 (K=8) or f16 (K=16), N in {16,32,64,128}.  Data is garbage; only timing matters.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../neuralmonkey_b200/csrc/tc_ptx.cuh"
using namespace nm;

__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__global__ void probe(long long* out, int nmma, int N, int mode, int nwarps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bars[4];
  __shared__ uint32_t slot;
  const uint32_t sb = smem_u32(smem);
  const uint32_t barA = smem_u32(&bar);
  if (threadIdx.x == 0) { mbar_init(barA, 1); for (int w = 0; w < 4; ++w) mbar_init(smem_u32(&bars[w]), 1); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tm = slot;
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && warp < nwarps) {
    const bool f16 = mode & 1, a_smem = mode & 2;
    const uint32_t fmt = f16 ? 0u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t SBO = 320 * 32;
    const uint64_t db = smem_desc(sb, 128, SBO, 0);
    const uint64_t da = smem_desc(sb + 64 * 1024, 128, SBO, 0);
    const uint32_t mybar = smem_u32(&bars[warp]);
    uint32_t ph = 0;
    const int per = nmma / nwarps;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      const uint32_t d = tm + 256 + warp * 64;
#pragma unroll 4
      for (int k = warp * per; k < (warp + 1) * per; ++k) {
        const uint32_t acc = k > warp * per;
        if (!f16) {
          if (a_smem) umma_tf32(d, da + k * 16, db + k * 16, idesc, acc);
          else umma_tf32_ts(d, tm + k * 8, db + k * 16, idesc, acc);
        } else {
          if (a_smem) umma_f16_ss(d, da + k * 16, db + k * 16, idesc, acc);
          else umma_f16_ts(d, tm + k * 8, db + k * 16, idesc, acc);
        }
      }
      const long long t1 = clock64();
      umma_commit(mybar);
      mbar_wait(mybar, ph); ph ^= 1;
      const long long t2 = clock64();
      if (warp == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  tcgen05_fence_before(); __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[4] = {"tf32 A=tmem", "f16  A=tmem", "tf32 A=smem", "f16  A=smem"};
  for (int mode = 0; mode < 4; ++mode)
    for (int N : {32, 64})
      for (int nw : {1, 2, 4}) {
        const int nm = 40;
        probe<<<1, 128, 200 * 1024>>>(d, nm, N, mode, nw);
        long long h[6]; cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
        printf("%s N=%3d nmma=%2d warps=%d  issue %lld total %lld cycles (%.1f/mma)  %s\n", names[mode], N, nm, nw, h[4], h[5], (double)h[5] / nm, e ? cudaGetErrorString(e) : "");
      }
  return 0;
}
