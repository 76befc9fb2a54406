// common.cuh -- small device helpers shared by the VMLMF kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vmlmf {

constexpr int kWarp = 32;

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ constexpr int round_up(int a, int b) { return ceil_div(a, b) * b; }
__host__ __device__ constexpr int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}
__host__ __device__ constexpr int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// Gate non-linearities.  Parity budget: 1e-5 relative against the reference's fp32 eager ops.
// Default: ex2.approx / rcp.approx forms (4 and 7 instructions; each MUFU op is good to ~2^-22
// relative, so a gate value carries ~2e-7 relative error -- same order as the fp32 noise of the
// reference itself, measured end to end in tests/test_gpu_parity.py).  -DVMLMF_ACCURATE_MATH
// switches to expf/tanhf/IEEE reciprocal (~18-25 instructions each).
__device__ __forceinline__ float ex2_approx(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float rcp_approx(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
#ifdef VMLMF_ACCURATE_MATH
__device__ __forceinline__ float sigmoidf_acc(float v) { return __frcp_rn(1.0f + expf(-v)); }
__device__ __forceinline__ float tanhf_acc(float v) { return tanhf(v); }
#else
__device__ __forceinline__ float sigmoidf_acc(float v) {
  // 1 / (1 + 2^(-v log2 e)); v -> -inf gives rcp(inf) = 0, v -> +inf gives rcp(1) = 1
  return rcp_approx(1.0f + ex2_approx(v * -1.4426950408889634f));
}
__device__ __forceinline__ float tanhf_acc(float v) {
  // (1 - e) / (1 + e), e = exp(-2|v|) in (0,1]: no overflow; sign restored at the end
  const float e = ex2_approx(fabsf(v) * -2.8853900817779268f);
  return copysignf((1.0f - e) * rcp_approx(1.0f + e), v);
}
#endif

// Round an fp32 value to the nearest tf32 (10 explicit mantissa bits, ties away from zero, like cvt.rna.tf32.f32),
// returned as an fp32 bit pattern with the low 13 bits cleared.  Two integer instructions; ptxas expands
// cvt.rna.tf32.f32 into five (it also canonicalises NaN, which never reaches these kernels).
__device__ __forceinline__ float tf32_rna(float v) {
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
}

// Sum N (power of two, <= 32) per-lane values across the 32 lanes of a warp with
// N-1 + log2(32/N) shuffles instead of 5N ("transposing" butterfly: every stage halves the
// number of live values, each lane keeps the half selected by one of its lane-id bits).
// On return, v[0] in lane l is the warp-wide sum of value index  bitrev_{log2 N}(l mod N).
template <int N>
__device__ __forceinline__ float warp_multi_reduce(float (&v)[N], int lane) {
  static_assert(N >= 1 && N <= 32 && (N & (N - 1)) == 0, "N must be a power of two <= 32");
  int mask = 1;
#pragma unroll
  for (int cur = N; cur > 1; cur >>= 1) {
    const int half = cur >> 1;
    const bool upper = (lane & mask) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
    mask <<= 1;
  }
#pragma unroll
  for (; mask < 32; mask <<= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], mask);
  return v[0];
}

// value index held by `lane` after warp_multi_reduce<N>
template <int N>
__device__ __forceinline__ int warp_multi_reduce_index(int lane) {
  constexpr int L = ilog2(N);
  int idx = 0;
#pragma unroll
  for (int b = 0; b < L; ++b) idx |= ((lane >> b) & 1) << (L - 1 - b);
  return idx;
}

__device__ __forceinline__ float ldg_f(const float* p) { return __ldg(p); }

// ---- host side: per-DEVICE caches.  Function attributes (cudaFuncSetAttribute), occupancy and the SM count belong
// to a device, not to the process: a process that uses several GPUs (torch.cuda.device_of) must set / query them once
// per device.  Benign races throughout: every writer stores the same value.
constexpr int kMaxDevices = 64;
inline int device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) { (void)cudaGetLastError(); return kMaxDevices; }   // no device: shared slot
  return (d >= 0 && d < kMaxDevices) ? d : kMaxDevices;
}
struct PerDevice {
  int v[kMaxDevices + 1] = {0};
  int& cur() { return v[device_slot()]; }
};
inline int num_sms() {
  static PerDevice cache;
  int& n = cache.cur();
  if (n == 0) {
    int d = 0, q = 0;
    if (cudaGetDevice(&d) != cudaSuccess || cudaDeviceGetAttribute(&q, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || q <= 0) {
      (void)cudaGetLastError();
      q = 148;                                         // B200; also what vmlmf_seq_plan assumes when no device is visible
    }
    n = q;
  }
  return n;
}

}  // namespace vmlmf
