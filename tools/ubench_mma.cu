// ubench_mma.cu -- issue / completion cost of tcgen05.mma kind::tf32 with shared-memory operands as a function of N, of the
// A operand (128 / 64 rows in shared memory, or tensor memory), and of HOW the single issuing thread is selected
// (inside `if (threadIdx.x == 0)` versus a warp-uniform loop with elect.sync around the instruction).  Sized the MMA loops of
// seq_r3.cuh / seq_r2.cuh.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I vmlmf_b200/csrc tools/ubench_mma.cu -o build/ubench_mma
#include <cstdio>
#include <cuda_runtime.h>
#include "gemm_tc.cuh"
using namespace vmlmf::tc;

__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

// OM: 0 = SS M=128, 1 = SS M=64, 2 = TS M=128 (A in tensor memory), 3 = SS M=128 with the same A descriptor every time
// EXTRA: 0 = MMAs only, 1 = + tcgen05.commit to a second barrier after every 4 MMAs, 2 = + a wait on an already completed barrier and a
// tcgen05.fence::after_thread_sync before every 4 MMAs (the per-K-tile bookkeeping of a TMA-fed pipeline), 3 = both
template <int OM, bool UNIFORM, int EXTRA = 0>
__global__ void __launch_bounds__(128, 1) k(int n_tiles, int N, int nacc_mask, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_done, bar_misc, bar_ready;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(&bar_done, 1); mbar_init(&bar_misc, 1); mbar_init(&bar_ready, 1); fence_barrier_init(); mbar_arrive(&bar_ready); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (UNIFORM ? threadIdx.x < 32 : threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(OM == 1 ? 64 : 128, N);
    const uint64_t da = make_desc(smem_u32(base)), db = make_desc(smem_u32(base + 16384));
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < n_tiles; ++i) {
        const uint32_t acc = tmem + (i & nacc_mask) * N, fl = i > nacc_mask ? 1u : 0u;
        if (EXTRA & 2) { mbar_wait(&bar_ready, 0); tc_fence_after(); }
        if (!UNIFORM || elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (OM == 2) mma_tf32_ts(acc, tmem + 256 + 8 * ks, db + 2 * ks, idesc, (fl | ks) ? 1u : 0u);
            else if (OM == 3) mma_tf32_ss(acc, da, db + 2 * ks, idesc, (fl | ks) ? 1u : 0u);
            else mma_tf32_ss(acc, da + 2 * ks, db + 2 * ks, idesc, (fl | ks) ? 1u : 0u);
          }
          if (EXTRA & 1) mma_commit(&bar_misc);
        }
      }
      if (!UNIFORM || elect_one()) mma_commit(&bar_done);
      const long long t1 = clock64();
      mbar_wait(&bar_done, rep & 1);
      const long long t2 = clock64();
      if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  }
}

template <int OM, bool UNIFORM, int EXTRA = 0>
int run(const char* name, long long* d) {
  cudaFuncSetAttribute(k<OM, UNIFORM, EXTRA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  const int tiles = 16;
  printf("-- %s\n", name);
  for (int N : {32, 64, 128, 256})
    for (int nacc : {1}) {
      if (nacc * N > 256) continue;
      k<OM, UNIFORM, EXTRA><<<1, 128, 66 * 1024>>>(tiles, N, nacc - 1, d);
      long long h[2];
      cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("  N %3d  accumulators %d | issue %7.1f cycles/MMA   issue+completion %7.1f cycles/MMA\n", N, nacc, (double)h[0] / (4 * tiles),
             (double)h[1] / (4 * tiles));
    }
  return 0;
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  if (run<0, false>("SS M=128, issued inside if (threadIdx.x == 0)", d)) return 1;
  if (run<0, true>("SS M=128, warp-uniform loop, elect.sync around the MMAs", d)) return 1;
  if (run<1, true>("SS M=64, elect", d)) return 1;
  if (run<2, true>("TS M=128 (A in tensor memory), elect", d)) return 1;
  if (run<3, true>("SS M=128, same A descriptor for every MMA, elect", d)) return 1;
  if (run<2, false>("TS M=128, if (threadIdx.x == 0)", d)) return 1;
  if (run<0, false, 1>("SS M=128, if (threadIdx.x == 0), + commit per 4 MMAs", d)) return 1;
  if (run<0, false, 2>("SS M=128, if (threadIdx.x == 0), + barrier wait and fence per 4 MMAs", d)) return 1;
  if (run<0, false, 3>("SS M=128, if (threadIdx.x == 0), + wait, fence and commit per 4 MMAs", d)) return 1;
  if (run<0, true, 3>("SS M=128, elect, + wait, fence and commit per 4 MMAs", d)) return 1;
  return 0;
}
