// test_gemm_tc.cu -- standalone check of csrc/gemm_tc.cuh against a double-precision CPU product (development aid).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o build/test_gemm_tc tools/test_gemm_tc.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../vmlmf_b200/csrc/gemm_tc.cuh"
using namespace vmlmf;

static int run(int M, int N, int K, int lda, int ldb) {
  std::vector<float> A((size_t)M * lda), B((size_t)N * ldb), C((size_t)M * N, -7.f);
  srand(M * 31 + N * 7 + K);
  for (auto& v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (auto& v : B) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  float *dA, *dB, *dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, C.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dC, C.data(), C.size() * 4, cudaMemcpyHostToDevice);
  int rc = tc::gemm_tc(dA, lda, dB, ldb, M, N, K, tc::EpiStoreTC{dC, N, 0}, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("M=%d N=%d K=%d lda=%d ldb=%d rc=%d sync=%s\n", M, N, K, lda, ldb, rc, cudaGetErrorString(e));
  if (rc || e != cudaSuccess) return 1;
  cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
  double num = 0, den = 0, worst = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[(size_t)m * lda + k] * B[(size_t)n * ldb + k];
      const double d = C[(size_t)m * N + n] - s;
      num += d * d; den += s * s;
      if (fabs(d) > worst) worst = fabs(d);
    }
  printf("   rel-l2 %.3e  max abs err %.3e\n", sqrt(num / den), worst);
  // timing
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) tc::gemm_tc(dA, lda, dB, ldb, M, N, K, tc::EpiStoreTC{dC, N, 0}, 0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("   %.3f ms per call, %.1f TFLOP/s (fp32-accurate flops)\n", ms / 10, 2.0 * M * N * K / (ms / 10 * 1e-3) / 1e12);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return sqrt(num / den) < 2e-6 ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run(128, 128, 32, 32, 32);
  bad += run(128, 128, 64, 64, 64);
  bad += run(256, 384, 96, 96, 100);
  bad += run(700, 2600, 300, 300, 300);       // LM XP at B=20: ragged M, N, K
  bad += run(100, 300, 652, 652, 652);
  bad += run(17920, 2600, 300, 300, 300);     // LM XP at B=512
  bad += run(8192, 8192, 1024, 1024, 1024);
  printf(bad ? "FAILED %d\n" : "ALL OK\n", bad);
  return bad;
}
