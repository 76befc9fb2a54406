// tail.cuh -- the callers either side of the recurrence (SURVEY.md 8 "next" rows f1, f2 and the Net head of a4):
//   * softmax-NLL loss, forward and backward  (F.cross_entropy, V/train_test/train.py:63; nll_loss, lm_test.py:140-153)
//   * the small classifier head of Net          (nn.Linear(H,18) on the last step, V/models/vmlmf.py:345-347,:354-355)
//   * flat-bucket optimizer steps               (Adam, train.py:47,65; clip-norm + SGD, lm_test.py:203-209)
// All of them are HBM/launch-bound elementwise or skinny work: one pass over the data, fixed-order reductions
// (bit-reproducible run to run), no atomics, no library calls.
#pragma once
#include <float.h>

#include "common.cuh"

namespace vmlmf {

constexpr int kTailThreads = 256;
constexpr int kTailMaxBlocks = 148 * 8;

// ---- fixed-order block sum: every thread passes its value, thread 0 gets the total ----
__device__ __forceinline__ float block_sum_256(float v, float* red /* [8] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();                                   // red may still be read from a previous call
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < kTailThreads / 32; ++w) t += red[w];
  }
  return t;
}

// ============================================================================================================
// softmax-NLL.  row_loss = logsumexp(row) - row[label];  loss = scale * sum_rows row_loss.
// One read of the scores (online max/sum), lse[rows] kept for backward.
// ============================================================================================================
struct MaxSum { float m, s; };
__device__ __forceinline__ MaxSum ms_push(MaxSum a, float v) {
  if (v > a.m) { a.s = a.s * expf(a.m - v) + 1.f; a.m = v; }
  else a.s += expf(v - a.m);
  return a;
}
__device__ __forceinline__ MaxSum ms_merge(MaxSum a, MaxSum b) {
  const float m = fmaxf(a.m, b.m);
  return MaxSum{m, a.s * expf(a.m - m) + b.s * expf(b.m - m)};
}
__device__ __forceinline__ MaxSum ms_warp(MaxSum a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxSum b{__shfl_xor_sync(0xffffffffu, a.m, o), __shfl_xor_sync(0xffffffffu, a.s, o)};
    a = ms_merge(a, b);
  }
  return a;
}

// WARP_ROWS = true: one warp per row (C <= 2048), 8 rows per block; false: one block per row.
template <bool WARP_ROWS>
__global__ void __launch_bounds__(kTailThreads)
softmax_nll_fwd_kernel(const float* __restrict__ scores, long long ld, const long long* __restrict__ labels,
                       float* __restrict__ lse, float* __restrict__ partials, long long rows, int C) {
  __shared__ float red[8];
  __shared__ MaxSum wms[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float my_loss = 0.f;                                // valid in lane 0 of each warp (WARP_ROWS) / thread 0
  if (WARP_ROWS) {
    const long long r = (long long)blockIdx.x * 8 + warp;
    if (r < rows) {
      const float* row = scores + r * ld;
      MaxSum a{-FLT_MAX, 0.f};
      for (int c = lane; c < C; c += 32) a = ms_push(a, __ldg(row + c));
      a = ms_warp(a);
      if (lane == 0) {
        const float l = a.m + logf(a.s);
        lse[r] = l;
        const long long y = labels[r];
        my_loss = (y >= 0 && y < C) ? l - __ldg(row + y) : __int_as_float(0x7fc00000);   // out-of-range label: NaN loss, no stray read
      }
    }
    const float t = block_sum_256(lane == 0 ? my_loss : 0.f, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
  } else {
    const long long r = blockIdx.x;
    const float* row = scores + r * ld;
    MaxSum a{-FLT_MAX, 0.f};
    if ((ld & 3) == 0 && (reinterpret_cast<uintptr_t>(scores) & 15) == 0) {
      const float4* row4 = reinterpret_cast<const float4*>(row);
      const int c4 = C >> 2;
      for (int c = threadIdx.x; c < c4; c += kTailThreads) {
        const float4 v = __ldg(row4 + c);
        a = ms_push(ms_push(ms_push(ms_push(a, v.x), v.y), v.z), v.w);
      }
      for (int c = (c4 << 2) + threadIdx.x; c < C; c += kTailThreads) a = ms_push(a, __ldg(row + c));
    } else {
      for (int c = threadIdx.x; c < C; c += kTailThreads) a = ms_push(a, __ldg(row + c));
    }
    a = ms_warp(a);
    if (lane == 0) wms[warp] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      MaxSum t = wms[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) t = ms_merge(t, wms[w]);
      const float l = t.m + logf(t.s);
      lse[r] = l;
      const long long y = labels[r];
      partials[r] = (y >= 0 && y < C) ? l - __ldg(row + y) : __int_as_float(0x7fc00000);
    }
  }
}

// loss = scale * sum(partials[0..n)) in a fixed order (one block)
__global__ void __launch_bounds__(kTailThreads) sum_scale_kernel(const float* __restrict__ partials, long long n,
                                                                 float scale, float* __restrict__ out) {
  __shared__ float red[8];
  float v = 0.f;
  for (long long i = threadIdx.x; i < n; i += kTailThreads) v += partials[i];
  const float t = block_sum_256(v, red);
  if (threadIdx.x == 0) *out = t * scale;
}

// dscores[r,c] = (exp(scores[r,c] - lse[r]) - [c == label[r]]) * scale * (dloss ? *dloss : 1); may run in place.
// Same row mapping as the forward: a warp per row (8 rows per block) or a block per row.
template <bool WARP_ROWS>
__global__ void __launch_bounds__(kTailThreads)
softmax_nll_bwd_kernel(const float* __restrict__ scores, long long ld, const long long* __restrict__ labels,
                       const float* __restrict__ lse, const float* __restrict__ dloss, float scale,
                       float* __restrict__ dscores, long long ldd, long long rows, int C) {
  const float gs = scale * (dloss ? __ldg(dloss) : 1.f);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = WARP_ROWS ? (long long)blockIdx.x * 8 + warp : (long long)blockIdx.x;
  if (r >= rows) return;
  const int first = WARP_ROWS ? lane : threadIdx.x, step = WARP_ROWS ? 32 : kTailThreads;
  const float l = __ldg(lse + r);
  const int lab = (int)labels[r];
  const float* src = scores + r * ld;
  float* dst = dscores + r * ldd;
  if (!WARP_ROWS && ((ld | ldd) & 3) == 0 && ((reinterpret_cast<uintptr_t>(scores) | reinterpret_cast<uintptr_t>(dscores)) & 15) == 0) {
    const int c4 = C >> 2;
    for (int c = first; c < c4; c += step) {
      float4 v = reinterpret_cast<const float4*>(src)[c];
      const int b = c << 2;
      v.x = (expf(v.x - l) - (lab == b ? 1.f : 0.f)) * gs;
      v.y = (expf(v.y - l) - (lab == b + 1 ? 1.f : 0.f)) * gs;
      v.z = (expf(v.z - l) - (lab == b + 2 ? 1.f : 0.f)) * gs;
      v.w = (expf(v.w - l) - (lab == b + 3 ? 1.f : 0.f)) * gs;
      reinterpret_cast<float4*>(dst)[c] = v;
    }
    for (int c = (c4 << 2) + first; c < C; c += step) dst[c] = (expf(src[c] - l) - (lab == c ? 1.f : 0.f)) * gs;
  } else {
    for (int c = first; c < C; c += step) dst[c] = (expf(src[c] - l) - (lab == c ? 1.f : 0.f)) * gs;
  }
}

// ============================================================================================================
// Net head: out[b,n] = sum_k h[b,k] W[n,k] + bias[n], N <= NP <= 32 (18 in the reference), K = hidden size.
// forward: warp per 2 rows, lane owns k = lane + 32 i; W in shared memory ([NP][K], rows >= N zero).
// ============================================================================================================
template <int NP>
__global__ void __launch_bounds__(kTailThreads)
head_fwd_kernel(const float* __restrict__ h, long long ldh, const float* __restrict__ W, const float* __restrict__ bias,
                float* __restrict__ out, int B, int K, int N) {
  extern __shared__ float sw[];                       // [NP][K]
  for (int i = threadIdx.x; i < NP * K; i += kTailThreads) sw[i] = (i / K) < N ? __ldg(W + i) : 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (kTailThreads / 32);
  for (int pair = blockIdx.x * (kTailThreads / 32) + warp; pair * 2 < B; pair += nwarps) {
    const int b0 = pair * 2, b1 = min(b0 + 1, B - 1);
    float a0[NP], a1[NP];
#pragma unroll
    for (int n = 0; n < NP; ++n) a0[n] = a1[n] = 0.f;
    for (int kb = 0; kb < K; kb += 256) {               // 8 k's per lane per chunk, all loads issued before the math
      float h0[8], h1[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kb + lane + 32 * i;
        h0[i] = k < K ? __ldg(h + (size_t)b0 * ldh + k) : 0.f;
        h1[i] = k < K ? __ldg(h + (size_t)b1 * ldh + k) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = min(kb + lane + 32 * i, K - 1);   // out-of-range lanes multiply a zero h
#pragma unroll
        for (int n = 0; n < NP; ++n) {
          const float w = sw[n * K + k];
          a0[n] = fmaf(h0[i], w, a0[n]);
          a1[n] = fmaf(h1[i], w, a1[n]);
        }
      }
    }
    const float s0 = warp_multi_reduce<NP>(a0, lane), s1 = warp_multi_reduce<NP>(a1, lane);
    const int n = warp_multi_reduce_index<NP>(lane & (NP - 1));
    if (lane < NP && n < N) {
      const float bv = bias ? __ldg(bias + n) : 0.f;
      out[(size_t)b0 * N + n] = s0 + bv;
      if (b0 + 1 < B) out[(size_t)b0 * N + N + n] = s1 + bv;
    }
  }
}

// backward: thread = k (K <= 256 * KJ), block = a contiguous chunk of rows.  dh[b,k] = sum_n dout[b,n] W[n,k];
// per-block partial dW[n,k] = sum_b dout[b,n] h[b,k], db[n] = sum_b dout[b,n]  ->  partials[block][NP*K + NP].
template <int NP, int KJ>
__global__ void __launch_bounds__(kTailThreads)
head_bwd_kernel(const float* __restrict__ h, long long ldh, const float* __restrict__ W, const float* __restrict__ dout,
                float* __restrict__ dh, long long lddh, float* __restrict__ partials, int B, int K, int N,
                int rows_per_block) {
  extern __shared__ __align__(16) float sd[];         // [rows_per_block][NP] dout rows, zero padded
  const int r0 = blockIdx.x * rows_per_block;
  const int nr = max(0, min(rows_per_block, B - r0));
  for (int i = threadIdx.x; i < rows_per_block * NP; i += kTailThreads) {
    const int r = i / NP, n = i - r * NP;
    sd[i] = (r < nr && n < N) ? __ldg(dout + (size_t)(r0 + r) * N + n) : 0.f;
  }
  __syncthreads();
  float* part = partials + (size_t)blockIdx.x * (NP * K + NP);
#pragma unroll
  for (int j = 0; j < KJ; ++j) {
    const int k = threadIdx.x + j * kTailThreads;
    if (k >= K) break;
    float wn[NP], acc[NP];
#pragma unroll
    for (int n = 0; n < NP; ++n) { wn[n] = n < N ? __ldg(W + (size_t)n * K + k) : 0.f; acc[n] = 0.f; }
    for (int rb = 0; rb < nr; rb += 8) {               // 8 rows of h in flight per thread
      float hv8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) hv8[u] = rb + u < nr ? __ldg(h + (size_t)(r0 + rb + u) * ldh + k) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = rb + u;
        if (r >= nr) break;
        const float hv = hv8[u];
        const float4* d4 = reinterpret_cast<const float4*>(sd + r * NP);
        float g = 0.f;
#pragma unroll
        for (int q = 0; q < NP / 4; ++q) {
          const float4 d = d4[q];
          g = fmaf(d.x, wn[4 * q], g); g = fmaf(d.y, wn[4 * q + 1], g);
          g = fmaf(d.z, wn[4 * q + 2], g); g = fmaf(d.w, wn[4 * q + 3], g);
          acc[4 * q] = fmaf(d.x, hv, acc[4 * q]); acc[4 * q + 1] = fmaf(d.y, hv, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(d.z, hv, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(d.w, hv, acc[4 * q + 3]);
        }
        if (dh) dh[(size_t)(r0 + r) * lddh + k] = g;
      }
    }
#pragma unroll
    for (int n = 0; n < NP; ++n) part[n * K + k] = acc[n];
  }
  if (threadIdx.x < NP) {
    float s = 0.f;
    for (int r = 0; r < nr; ++r) s += sd[r * NP + threadIdx.x];
    part[NP * K + threadIdx.x] = s;
  }
}

// dW[n,k] / db[n] = sum over blocks of the partials in a fixed order: 16 outputs per block, each summed by 16 threads
// over interleaved slices of the block list (many independent loads in flight), then across slices in slice order.
__global__ void __launch_bounds__(kTailThreads)
head_reduce_kernel(const float* __restrict__ partials, int nblocks, int NP, int K, int N, float* __restrict__ dW,
                   float* __restrict__ db) {
  __shared__ float sl[16][17];
  const int e = threadIdx.x & 15, slice = threadIdx.x >> 4;
  const int i = blockIdx.x * 16 + e;
  const int stride = NP * K + NP;
  const bool is_w = i < N * K, is_b = !is_w && i < N * K + N;
  const int off = is_w ? i : NP * K + (i - N * K);
  float s = 0.f;
  if (is_w || is_b) {
#pragma unroll 4
    for (int b = slice; b < nblocks; b += 16) s += partials[(size_t)b * stride + off];
  }
  sl[slice][e] = s;
  __syncthreads();
  if (slice == 0 && (is_w || is_b)) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) t += sl[q][e];
    if (is_w) dW[i] = t;
    else if (db) db[i - N * K] = t;
  }
}

// ============================================================================================================
// Flat-bucket optimizer steps (parameters, gradients and moments are single contiguous fp32 buffers).
// ============================================================================================================
// torch.optim.Adam semantics (no amsgrad / weight decay): m += (g-m)(1-b1); v = b2 v + (1-b2) g^2;
// p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).  t comes from a device counter (CUDA-graph replay) or the host.
__global__ void __launch_bounds__(kTailThreads)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, float lr, float b1, float b2, float eps, const float* __restrict__ step_dev, int step_host) {
  __shared__ float sc[2];
  if (threadIdx.x == 0) {
    const double t = step_dev ? (double)*step_dev : (double)step_host;
    sc[0] = (float)((double)lr / (1.0 - pow((double)b1, t)));
    sc[1] = (float)(1.0 / sqrt(1.0 - pow((double)b2, t)));
  }
  __syncthreads();
  const float step_size = sc[0], inv_bc2_sqrt = sc[1];
  for (long long i = (long long)blockIdx.x * kTailThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kTailThreads) {
    const float gi = g[i];
    const float mi = fmaf(gi - m[i], 1.f - b1, m[i]);
    const float vi = fmaf(gi * gi, 1.f - b2, b2 * v[i]);
    m[i] = mi; v[i] = vi;
    p[i] -= step_size * (mi / (sqrtf(vi) * inv_bc2_sqrt + eps));
  }
}

// Data-parallel Adam: the gradient all-reduce fused into the optimizer step over NVLink peer memory (one-shot all-reduce).
// peer[r] is rank r's flat gradient bucket, mapped into this process (symmetric memory); every rank reads all `world`
// buckets and adds them in rank order, so the reduced gradient -- and with it the updated replica -- is bit-identical on
// every rank.  For the small factor-gradient buckets of the HAR nets (94 KB at cfg2) the step is latency-bound: this is one
// kernel between two cross-rank barriers instead of ncclAllReduce + scale + Adam.  Loads bypass L1 (peer data changes every step).
__device__ __forceinline__ float ld_peer(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__global__ void __launch_bounds__(kTailThreads)
p2p_adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* const* __restrict__ peer,
                int world, long long n, float scale, float lr, float b1, float b2, float eps,
                const float* __restrict__ step_dev, int step_host) {
  __shared__ float sc[2];
  __shared__ const float* sp[16];
  if (threadIdx.x == 0) {
    const double t = step_dev ? (double)*step_dev : (double)step_host;
    sc[0] = (float)((double)lr / (1.0 - pow((double)b1, t)));
    sc[1] = (float)(1.0 / sqrt(1.0 - pow((double)b2, t)));
  }
  if (threadIdx.x < world) sp[threadIdx.x] = peer[threadIdx.x];
  __syncthreads();
  const float step_size = sc[0], inv_bc2_sqrt = sc[1];
  for (long long i = (long long)blockIdx.x * kTailThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kTailThreads) {
    float gi = 0.f;
    for (int r = 0; r < world; ++r) gi += ld_peer(sp[r] + i);              // fixed rank order
    gi *= scale;
    const float mi = fmaf(gi - m[i], 1.f - b1, m[i]);
    const float vi = fmaf(gi * gi, 1.f - b2, b2 * v[i]);
    m[i] = mi; v[i] = vi;
    p[i] -= step_size * (mi / (sqrtf(vi) * inv_bc2_sqrt + eps));
  }
}

// LM input side: out[r, c] = W[tok[r], c] * (mask ? mask[r, c] * scale : 1) for c < E, 0 for E <= c < ldo.
// Embed (index -> row, V/models/vmlmf_lm.py:48) and the dropout that follows it (:436) in one pass, written into a
// pitch-padded buffer (ldo % 4 == 0) so that the x-projection GEMM can read it in place as its TMA A operand.
__global__ void __launch_bounds__(kTailThreads)
embed_dropout_kernel(const long long* __restrict__ tok, const float* __restrict__ W, const unsigned char* __restrict__ mask,
                     float scale, float* __restrict__ out, long long ldo, long long rows, int E, int V) {
  const long long n = rows * ldo;
  for (long long i = (long long)blockIdx.x * kTailThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kTailThreads) {
    const long long r = i / ldo;
    const int c = (int)(i - r * ldo);
    float v = 0.f;
    if (c < E) {
      const long long t = tok[r];
      v = (t >= 0 && t < V) ? __ldg(W + t * E + c) : __int_as_float(0x7fc00000);   // out-of-range token: NaN, no stray read
      if (mask) v = mask[r * E + c] ? v * scale : 0.f;
    }
    out[i] = v;
  }
}

// sum of squares of g, one partial per block (fixed grid-stride order)
__global__ void __launch_bounds__(kTailThreads) sumsq_kernel(const float* __restrict__ g, long long n,
                                                             float* __restrict__ partials) {
  __shared__ float red[8];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * kTailThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kTailThreads)
    s = fmaf(g[i], g[i], s);
  const float t = block_sum_256(s, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// clip_grad_norm_ (coef = min(1, max_norm / (norm + 1e-6))) followed by p -= lr * g  (lm_test.py:203-209).
// Every block re-derives the norm from the partials in the same order; scale_grads also writes g *= coef back.
__global__ void __launch_bounds__(kTailThreads)
sgd_clip_kernel(float* __restrict__ p, float* __restrict__ g, long long n, float lr, float max_norm,
                const float* __restrict__ partials, int nparts, int scale_grads, float* __restrict__ norm_out) {
  __shared__ float red[8];
  __shared__ float s_coef;
  float s = 0.f;
  for (int i = threadIdx.x; i < nparts; i += kTailThreads) s += partials[i];
  const float t = block_sum_256(s, red);
  if (threadIdx.x == 0) {
    const float norm = sqrtf(t);
    s_coef = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
    if (blockIdx.x == 0 && norm_out) *norm_out = norm;
  }
  __syncthreads();
  const float coef = s_coef;
  for (long long i = (long long)blockIdx.x * kTailThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kTailThreads) {
    const float gi = g[i] * coef;
    if (scale_grads) g[i] = gi;
    p[i] -= lr * gi;
  }
}

inline int tail_grid(long long n) {
  const long long b = (n + kTailThreads - 1) / kTailThreads;
  return (int)(b < 1 ? 1 : (b > kTailMaxBlocks ? kTailMaxBlocks : b));
}

}  // namespace vmlmf
