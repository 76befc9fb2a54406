// ubench.cu -- issue-rate microbenchmarks that size the SIMT recurrence kernels on B200 (sm_100a).
// Measures warp-instructions per cycle per SM sub-partition for the instruction mix of seq_r1.cuh:
// FFMA (3 distinct registers), packed FFMA2 (fma.rn.f32x2), MUFU ex2 / rcp, SHFL, and LDS.128.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu
// Run  : build/ubench   (prints one line per test; used only to choose kernel structure, never a bench value)
#include <cuda_runtime.h>
#include <stdio.h>

#define ITERS 2000

template <int MODE>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[16], w[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed * (i + 1) + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = seed * 0.001f * (i + 3);
  __shared__ float4 sm[256];
  sm[threadIdx.x & 255] = make_float4(seed, seed, seed, seed);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {            // 64 FFMA, acc += w*z pattern (3 distinct regs)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(w[(i + r) & 7], w[(r * 3 + 1) & 7], a[i]);
    } else if (MODE == 1) {     // 32 FFMA2 (= 64 FMA)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          unsigned long long d, x, y;
          asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a[i]), "f"(a[i + 1]));
          asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(w[(i + r) & 7]), "f"(w[(i + r + 1) & 7]));
          asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(w[(r * 3 + 1) & 7]), "f"(w[(r * 3 + 1) & 7]));
          asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(x), "l"(y));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(d));
        }
    } else if (MODE == 2) {     // 16 ex2 + 16 rcp
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      }
    } else if (MODE == 3) {     // 32 shuffles
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + r);
    } else if (MODE == 4) {     // 16 LDS.128 broadcast + 16 FADD to consume
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float4 v = sm[(it + i) & 255];
        a[i] += v.x + v.w;
      }
    } else if (MODE == 5) {     // 64 FFMA with an immediate-like / 2-distinct-reg pattern: a = a*w + a
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], w[r], a[i]);
    } else if (MODE == 7 || MODE == 8) {   // 16 mma.m16n8k8 tf32: 8 independent accumulators (7) / 1 dependent chain (8)
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int c = (MODE == 7) ? (r & 3) * 4 : 0;
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(a[c]), "+f"(a[c + 1]), "+f"(a[c + 2]), "+f"(a[c + 3])
                     : "r"(__float_as_uint(w[0])), "r"(__float_as_uint(w[1])), "r"(__float_as_uint(w[2])),
                       "r"(__float_as_uint(w[3])), "r"(__float_as_uint(w[4])), "r"(__float_as_uint(w[5])));
      }
    } else if (MODE == 9) {   // 16 mma.m16n8k16 bf16, 4 independent accumulators
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int c = (r & 3) * 4;
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(a[c]), "+f"(a[c + 1]), "+f"(a[c + 2]), "+f"(a[c + 3])
                     : "r"(__float_as_uint(w[0])), "r"(__float_as_uint(w[1])), "r"(__float_as_uint(w[2])),
                       "r"(__float_as_uint(w[3])), "r"(__float_as_uint(w[4])), "r"(__float_as_uint(w[5])));
      }
    } else if (MODE == 10) {  // 32 cvt.rna.tf32
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          unsigned u;
          asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a[i]));
          a[i] = __uint_as_float(u) + w[r];
        }
    } else if (MODE == 6) {     // mix: 56 FFMA + 5 ex2 + 5 rcp (one unit-step of the forward)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 14; ++i) a[i] = fmaf(w[(i + r) & 7], w[(r * 3 + 1) & 7], a[i]);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter, float* out, long long* cyc) {
  for (int nw = 4; nw <= 32; nw *= 2) {
    k<MODE><<<148, nw * 32>>>(out, cyc, 1.0001f);
    k<MODE><<<148, nw * 32>>>(out, cyc, 1.0001f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double wi = (double)instr_per_iter * ITERS * nw / 4.0;    // warp-instructions per SMSP
    printf("%-28s warps/SMSP=%d  cycles=%.0f  cyc/warp-instr/SMSP=%.3f\n", name, nw / 4, avg, avg / wi);
  }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  run<0>("FFMA 3-reg (64/iter)", 64, out, cyc);
  run<5>("FFMA 2-reg (64/iter)", 64, out, cyc);
  run<1>("FFMA2 f32x2 (32/iter)", 32, out, cyc);
  run<2>("MUFU ex2+rcp (32/iter)", 32, out, cyc);
  run<3>("SHFL (32/iter)", 32, out, cyc);
  run<4>("LDS.128 bcast (16/iter)", 16, out, cyc);
  run<6>("mix 56 FFMA+10 MUFU (66)", 66, out, cyc);
  run<7>("HMMA tf32 m16n8k8 x4acc (16)", 16, out, cyc);
  run<8>("HMMA tf32 m16n8k8 chain (16)", 16, out, cyc);
  run<9>("HMMA bf16 m16n8k16 x4acc(16)", 16, out, cyc);
  run<10>("cvt.rna.tf32+FADD (64)", 64, out, cyc);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
