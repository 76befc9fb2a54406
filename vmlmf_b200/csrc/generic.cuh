// generic.cuh -- regime G: any (T,B,I,H,RX,RH) with I <= H.  Used when the factors do not fit the
// register-resident regime R1 (LM layer H=650 r=300, the H>=1024 sweep, wide group packings).
//
// Structure (fp32 SIMT throughout; fp32 parity, no TF32):
//   time-parallel   ZX = X Ux                    (vmlmf_xproj_fwd)            [T*B, RX]
//                   XP = ZX Vx^T + bias + x(.)Dx (epilogue-fused)             [T*B, 4H]
//   per timestep    z_t = h_{t-1} A              (GEMM, K=H)
//                   gates/c/h = f(XP_t + z_t Bm^T + h_{t-1}(.)Dh)  (GEMM K=RH + gate epilogue)
//   backward        per step: dPre_t (pointwise) ; dz_t = dPre_t Bm ; dh_{t-1} += dz_t A^T
//                   time-parallel, split-K, fixed-order: dBm = dPre^T Z, dA = Hprev^T dZ,
//                   dVx = dPre^T ZX, dZX = dPre Vx, dUx = X^T dZX, dX = dZX Ux^T + dPre(.)Dx,
//                   column reductions dDh, dDx, dbias.
// One 64x64x16 register-tiled GEMM template serves every contraction; operands are addressed
// through RowView so batch-first and time-major tensors need no copies.
#pragma once
#include "../../include/vmlmf_b200.h"
#include <stdlib.h>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace vmlmf {

// rows of a [rows, cols] matrix whose row r lives at p + (r / bdiv) * s_t + (r % bdiv) * s_b
struct RowView {
  float* p; long long s_t, s_b; int bdiv;
  __host__ __device__ float* row(long long r) const { return p + (r / bdiv) * s_t + (r % bdiv) * s_b; }
};
inline RowView plain_view(const float* p, long long ld) { return RowView{const_cast<float*>(p), 0, ld, 0x7fffffff}; }
inline RowView tb_view(const float* p, long long s_t, long long s_b, int B) { return RowView{const_cast<float*>(p), s_t, s_b, B}; }

constexpr int GBM = 64, GBN = 64, GBK = 16;

struct NIdent { __device__ long long operator()(int n) const { return n; } };
// gate-interleaved column order n' = j*4 + k  ->  stored row k*H + j of a [4H, R] factor
struct NGate { int H; __device__ long long operator()(int n) const { return (long long)(n & 3) * H + (n >> 2); } };

// ---- epilogues: called once per (row m, 4 consecutive columns n..n+3) ----
struct EpiStore {           // C = acc (+ C)
  RowView C; int beta;
  __device__ void operator()(int m, int n, int N, const float (&v)[4]) const {
    float* c = C.row(m) + n;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (n + q < N) c[q] = beta ? c[q] + v[q] : v[q];
  }
};
struct EpiPartial {         // raw partial tile for split-K: part[z][m][n]
  float* part; int M, N;
  __device__ void operator()(int m, int n, int, const float (&v)[4]) const {
    float* c = part + ((size_t)blockIdx.z * M + m) * N + n;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (n + q < N) c[q] = v[q];
  }
};
struct EpiXP {              // XP[m, kH+j] = acc + bias[kH+j] + [j<I] x[m,j] Dx[k,j]
  float* xp; const float* bias; RowView X; const float* Dx; int H, I;
  __device__ void operator()(int m, int n, int N, const float (&v)[4]) const {
    const float* xr = X.row(m);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int nn = n + q;
      if (nn < N) {
        const int k = nn / H, j = nn - k * H;
        float r = v[q] + bias[nn];
        if (j < I) r = fmaf(xr[j], Dx[k * I + j], r);
        xp[(size_t)m * N + nn] = r;
      }
    }
  }
};
struct EpiGate {            // columns are gate-interleaved: n = 4*j + k  (needs N = 4H, n % 4 == 0)
  const float* xp_t;        // XP rows of this step   [B, 4H]
  const float* hprev; long long hp_sb;    // h_{t-1}[b] = hprev + b*hp_sb (null = zeros)
  const float* cprev;                     // [B,H] or null
  const float* Dh;
  float* y_t; long long y_sb;             // h_t[b] = y_t + b*y_sb
  float* c_out;                           // [B,H]  (cs[t] when saving, else the running cT buffer)
  float* gates_t;                         // [B,4,H] or null
  float *hT, *cT;                         // written on the last step only (else null)
  int H;
  __device__ void operator()(int m, int n, int N, const float (&v)[4]) const {
    if (n >= N) return;
    const int j = n >> 2;
    const float hp = hprev ? hprev[(size_t)m * hp_sb + j] : 0.f;
    const float cp = cprev ? cprev[(size_t)m * H + j] : 0.f;
    const float* xr = xp_t + (size_t)m * 4 * H + j;
    float pre[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) pre[k] = v[k] + xr[(size_t)k * H] + hp * Dh[k * H + j];
    const float gi = sigmoidf_acc(pre[0]), gf = sigmoidf_acc(pre[1]);
    const float go = sigmoidf_acc(pre[2]), gn = tanhf_acc(pre[3]);
    const float c = fmaf(gf, cp, gi * gn);
    const float h = go * tanhf_acc(c);
    y_t[(size_t)m * y_sb + j] = h;
    c_out[(size_t)m * H + j] = c;
    if (gates_t) {
      float* g = gates_t + (size_t)m * 4 * H + j;
      g[0] = gi; g[H] = gf; g[2 * H] = go; g[3 * H] = gn;
    }
    if (hT) { hT[(size_t)m * H + j] = h; cT[(size_t)m * H + j] = c; }
  }
};
struct EpiDX {              // dX[m,j] = acc + sum_k dPre[m,kG+j] Dx[k,j]   (G = gate stride of dPre, row pitch 4G)
  RowView dX; const float* dpre; const float* Dx; int H, I;
  __device__ void operator()(int m, int n, int N, const float (&v)[4]) const {
    float* o = dX.row(m);
    const float* dp = dpre + (size_t)m * 4 * H;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = n + q;
      if (j < N) {
        float r = v[q];
#pragma unroll
        for (int k = 0; k < 4; ++k) r = fmaf(dp[k * H + j], Dx[k * I + j], r);
        o[j] = r;
      }
    }
  }
};

struct EpiDXTC {            // tensor-core epilogue of dX = dZX Ux^T + sum_k dPre_k Dx_k  (lane <-> column, see gemm_tc.cuh)
  float* dx; long long dxs_t, dxs_b; int Bsz; const float* dpre; const float* Dx; int H, I;   // H = gate stride of dPre
  static constexpr bool kGate = false;
  struct Col { float dx[4][4]; };                      // Dx[k][column i]
  struct In { float dp[4][4]; };                       // dPre[k][column i] of the row
  __device__ void cols(int n0, int N, int lane, Col& c) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = n0 + lane + 32 * i;
#pragma unroll
      for (int k = 0; k < 4; ++k) c.dx[k][i] = j < N ? __ldg(Dx + k * I + j) : 0.f;
    }
  }
  __device__ void load(int m, int n0, int N, int lane, const Col&, In& in) const {
    const float* dp = dpre + (size_t)m * 4 * H;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = n0 + lane + 32 * i;
#pragma unroll
      for (int k = 0; k < 4; ++k) in.dp[k][i] = j < N ? __ldg(dp + k * H + j) : 0.f;
    }
  }
  __device__ void finish(int m, int n0, int N, int lane, const float (&v)[4], const Col& c, const In& in) const {
    float* o = dx + (long long)(m / Bsz) * dxs_t + (long long)(m % Bsz) * dxs_b;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = n0 + lane + 32 * i;
      if (j < N) {
        float r = v[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) r = fmaf(in.dp[k][i], c.dx[k][i], r);
        o[j] = r;
      }
    }
  }
};

// C[m,n] = sum_k Aop(m,k) Bop(k,n).
//   A_T=false: A stored [M rows][K cols]   A_T=true: stored [K rows][M cols]
//   B_T=false: B stored [K rows][N cols]   B_T=true: stored [N rows][K cols], row = nmap(n)
// grid = (n tiles, m tiles, k splits); each split covers k_chunk consecutive k.
template <bool A_T, bool B_T, class NMap, class Epi>
__global__ void __launch_bounds__(256) gemm_kernel(RowView A, RowView Bv, int M, int N, long long K,
                                                   long long k_chunk, NMap nmap, Epi epi) {
  __shared__ __align__(16) float As[GBK][GBM + 4];
  __shared__ __align__(16) float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const long long kbeg = (long long)blockIdx.z * k_chunk;
  const long long kend = (kbeg + k_chunk < K) ? kbeg + k_chunk : K;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  for (long long k0 = kbeg; k0 < kend; k0 += GBK) {
#pragma unroll
    for (int e = 0; e < (GBM * GBK) / 256; ++e) {
      const int idx = e * 256 + tid;
      if constexpr (!A_T) {
        const int mm = idx / GBK, kk = idx % GBK;
        const int m = m0 + mm; const long long k = k0 + kk;
        As[kk][mm] = (m < M && k < kend) ? A.row(m)[k] : 0.f;
      } else {
        const int kk = idx / GBM, mm = idx % GBM;
        const int m = m0 + mm; const long long k = k0 + kk;
        As[kk][mm] = (m < M && k < kend) ? A.row(k)[m] : 0.f;
      }
    }
#pragma unroll
    for (int e = 0; e < (GBN * GBK) / 256; ++e) {
      const int idx = e * 256 + tid;
      if constexpr (B_T) {
        const int nn = idx / GBK, kk = idx % GBK;
        const int n = n0 + nn; const long long k = k0 + kk;
        Bs[kk][nn] = (n < N && k < kend) ? Bv.row(nmap(n))[k] : 0.f;
      } else {
        const int kk = idx / GBN, nn = idx % GBN;
        const int n = n0 + nn; const long long k = k0 + kk;
        Bs[kk][nn] = (n < N && k < kend) ? Bv.row(k)[n] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int m = m0 + ty * 4 + a;
    if (m < M) epi(m, n0 + tx * 4, N, acc[a]);
  }
}

template <bool A_T, bool B_T, class NMap, class Epi>
inline int gemm_launch(RowView A, RowView Bv, int M, int N, long long K, int splits, NMap nmap, Epi epi,
                       cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  long long chunk = (K + splits - 1) / splits;
  chunk = (chunk + GBK - 1) / GBK * GBK;
  dim3 grid(ceil_div(N, GBN), ceil_div(M, GBM), splits);
  gemm_kernel<A_T, B_T, NMap, Epi><<<grid, 256, 0, st>>>(A, Bv, M, N, K, chunk, nmap, epi);
  return (int)cudaGetLastError();
}

// out[i] = sum_z part[z][i]   (fixed order)
static __global__ void splitk_reduce_kernel(const float* __restrict__ part, int nsplit, long long n, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < nsplit; ++z) s += part[(size_t)z * n + i];
  out[i] = s;
}
// out[m*ldo + c] (+)= sum_z part[z][m*N + c]   (fixed order; rows of the output may be padded)
static __global__ void splitk_reduce_rows_kernel(const float* __restrict__ part, int nsplit, int M, int N,
                                                 float* __restrict__ out, long long ldo, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int m = (int)(i / N), c = (int)(i - (long long)m * N);
  float s = 0.f;
  for (int z = 0; z < nsplit; ++z) s += part[(size_t)z * M * N + i];
  float* o = out + (size_t)m * ldo + c;
  *o = accumulate ? *o + s : s;
}

// ---- backward, pointwise part of one timestep ----
struct DpreArgs {
  const float* gates_t;      // [B,4,H]
  const float* c_t;          // [B,H]
  const float* c_prev;       // [B,H] or null
  const float* dy_t; long long dy_sb;     // or null
  const float* Dh;
  float* dh;                 // [B,H] in: dh_next   out: sum_k dpre_k Dh_k   (seed of dh_{t-1})
  float* dc;                 // [B,H] in: dc_next   out: dc_next for t-1
  float* dpre_t;             // [B,4,H]
  int B, H;
};
static __global__ void dpre_step_kernel(const DpreArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)a.B * a.H) return;
  const int b = (int)(i / a.H), j = (int)(i % a.H);
  const float* g = a.gates_t + (size_t)b * 4 * a.H + j;
  const float gi = g[0], gf = g[a.H], go = g[2 * a.H], gn = g[3 * a.H];
  const float dh = a.dh[i] + (a.dy_t ? a.dy_t[(size_t)b * a.dy_sb + j] : 0.f);
  const float tc = tanhf_acc(a.c_t[i]);
  const float cp = a.c_prev ? a.c_prev[i] : 0.f;
  const float dc = fmaf(dh * go, 1.f - tc * tc, a.dc[i]);
  float d[4];
  d[0] = dc * gn * gi * (1.f - gi);
  d[1] = dc * cp * gf * (1.f - gf);
  d[2] = dh * tc * go * (1.f - go);
  d[3] = dc * gi * (1.f - gn * gn);
  float* o = a.dpre_t + (size_t)b * 4 * a.H + j;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    o[(size_t)k * a.H] = d[k];
    s = fmaf(d[k], a.Dh[k * a.H + j], s);
  }
  a.dc[i] = dc * gf;
  a.dh[i] = s;
}

// ---- column reductions over all T*B rows: dbias, dDh, dDx (split over rows, fixed order) ----
struct ColRedArgs {
  const float* dpre;         // [T*B, 4H]
  RowView Y; const float* h0; // h_{t-1}: t>0 -> Y.row((t-1)*B+b), t==0 -> h0[b] (null = 0)
  RowView X;
  float* part;               // [nsplit][4H + 4H + 4I]
  int T, B, H, I; long long rows_per_split;
  int G;                     // gate stride of dPre (row pitch 4G); G >= H
};
static __global__ void colreduce_kernel(const ColRedArgs a) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;          // column (gate k, unit j) of dPre
  if (n >= 4 * a.H) return;
  const int k = n / a.H, j = n - k * a.H;
  const long long rows = (long long)a.T * a.B;
  const long long r0 = (long long)blockIdx.y * a.rows_per_split;
  const long long r1 = r0 + a.rows_per_split < rows ? r0 + a.rows_per_split : rows;
  const bool has_x = j < a.I;
  float sb = 0.f, sh = 0.f, sx = 0.f;
  // (t, b) walked incrementally: no division per row; four rows of loads in flight
  int t = (int)(r0 / a.B), b = (int)(r0 - (long long)t * a.B);
  const float* dp = a.dpre + (size_t)r0 * 4 * a.G + (size_t)k * a.G + j;
  const size_t dstep = (size_t)4 * a.G;
  long long r = r0;
  while (r < r1) {
    float d[4], hp[4], xv[4];
    int nv = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      d[u] = hp[u] = xv[u] = 0.f;
      if (r + u < r1) {
        d[u] = dp[u * dstep];
        const float* hrow = t > 0 ? a.Y.p + (long long)(t - 1) * a.Y.s_t + (long long)b * a.Y.s_b
                                  : (a.h0 ? a.h0 + (size_t)b * a.H : nullptr);
        if (hrow) hp[u] = hrow[j];
        if (has_x) xv[u] = a.X.p[(long long)t * a.X.s_t + (long long)b * a.X.s_b + j];
        if (++b == a.B) { b = 0; ++t; }
        ++nv;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {           // fixed order; absent rows add zeros
      sb += d[u];
      sh = fmaf(d[u], hp[u], sh);
      sx = fmaf(d[u], xv[u], sx);
    }
    r += nv;
    dp += nv * dstep;
  }
  float* p = a.part + (size_t)blockIdx.y * (8 * a.H + 4 * a.I);
  p[n] = sb;
  p[4 * a.H + n] = sh;
  if (j < a.I) p[8 * a.H + k * a.I + j] = sx;
}
// one thread per FOUR consecutive columns of a gate (16-byte loads of dPre, h_{t-1} and x); needs H % 4 == 0, G % 4 == 0,
// I % 4 == 0 or I handled by the tail lanes scalar-wise -- the launcher falls back to the scalar kernel otherwise
static __global__ void colreduce4_kernel(const ColRedArgs a) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;          // column quad of dPre
  const int hq = a.H >> 2;
  if (q >= 4 * hq) return;
  const int k = q / hq, j = (q - k * hq) * 4;
  const long long rows = (long long)a.T * a.B;
  const long long r0 = (long long)blockIdx.y * a.rows_per_split;
  const long long r1 = r0 + a.rows_per_split < rows ? r0 + a.rows_per_split : rows;
  const int nx = a.I - j < 0 ? 0 : (a.I - j > 4 ? 4 : a.I - j);  // how many of the four columns have an input-side term
  const bool x_vec = nx == 4 && ((a.X.s_t | a.X.s_b) & 3) == 0 && (reinterpret_cast<uintptr_t>(a.X.p) & 15) == 0;
  const bool y_vec = ((a.Y.s_t | a.Y.s_b) & 3) == 0 && (reinterpret_cast<uintptr_t>(a.Y.p) & 15) == 0;
  float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), sh = sb, sx = sb;
  int t = (int)(r0 / a.B), b = (int)(r0 - (long long)t * a.B);
  const float* dp = a.dpre + (size_t)r0 * 4 * a.G + (size_t)k * a.G + j;
  const size_t dstep = (size_t)4 * a.G;
  for (long long r = r0; r < r1; r += 2) {
    float4 d[2], hp[2], xv[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      d[u] = hp[u] = xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r + u < r1) {
        d[u] = *reinterpret_cast<const float4*>(dp + u * dstep);
        const float* hrow = t > 0 ? a.Y.p + (long long)(t - 1) * a.Y.s_t + (long long)b * a.Y.s_b
                                  : (a.h0 ? a.h0 + (size_t)b * a.H : nullptr);
        if (hrow) {
          if (t > 0 ? y_vec : true) hp[u] = *reinterpret_cast<const float4*>(hrow + j);
          else hp[u] = make_float4(hrow[j], hrow[j + 1], hrow[j + 2], hrow[j + 3]);
        }
        if (nx) {
          const float* xr = a.X.p + (long long)t * a.X.s_t + (long long)b * a.X.s_b + j;
          if (x_vec) xv[u] = *reinterpret_cast<const float4*>(xr);
          else { xv[u].x = xr[0]; if (nx > 1) xv[u].y = xr[1]; if (nx > 2) xv[u].z = xr[2]; if (nx > 3) xv[u].w = xr[3]; }
        }
        if (++b == a.B) { b = 0; ++t; }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {           // fixed order; absent rows add zeros
      sb.x += d[u].x; sb.y += d[u].y; sb.z += d[u].z; sb.w += d[u].w;
      sh.x = fmaf(d[u].x, hp[u].x, sh.x); sh.y = fmaf(d[u].y, hp[u].y, sh.y);
      sh.z = fmaf(d[u].z, hp[u].z, sh.z); sh.w = fmaf(d[u].w, hp[u].w, sh.w);
      sx.x = fmaf(d[u].x, xv[u].x, sx.x); sx.y = fmaf(d[u].y, xv[u].y, sx.y);
      sx.z = fmaf(d[u].z, xv[u].z, sx.z); sx.w = fmaf(d[u].w, xv[u].w, sx.w);
    }
    dp += 2 * dstep;
  }
  float* p = a.part + (size_t)blockIdx.y * (8 * a.H + 4 * a.I);
  const int n = k * a.H + j;
  p[n] = sb.x; p[n + 1] = sb.y; p[n + 2] = sb.z; p[n + 3] = sb.w;
  p[4 * a.H + n] = sh.x; p[4 * a.H + n + 1] = sh.y; p[4 * a.H + n + 2] = sh.z; p[4 * a.H + n + 3] = sh.w;
  const float sxa[4] = {sx.x, sx.y, sx.z, sx.w};
  for (int e = 0; e < nx; ++e) p[8 * a.H + k * a.I + j + e] = sxa[e];
}
static __global__ void colreduce_final_kernel(const float* __restrict__ part, int nsplit, int H, int I,
                                              float* dbias, float* dDh, float* dDx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, tot = 8 * H + 4 * I;
  if (i >= tot) return;
  float s = 0.f;
  for (int z = 0; z < nsplit; ++z) s += part[(size_t)z * tot + i];
  if (i < 4 * H) dbias[i] = s;
  else if (i < 8 * H) dDh[i - 4 * H] = s;
  else dDx[i - 8 * H] = s;
}

// --------------------------------------------------------------------------------------------- //
// host side
// --------------------------------------------------------------------------------------------- //
inline int g_splits(int M, int N, long long K) {
  const long long tiles = (long long)ceil_div(M, GBM) * ceil_div(N, GBN);
  long long s = (2LL * num_sms() + tiles - 1) / tiles;
  const long long kmax = (K + 255) / 256;
  if (s > kmax) s = kmax;
  if (s > 128) s = 128;
  if (s < 1) s = 1;
  return (int)s;
}
constexpr int kColSplits = 256;
// The tcgen05 accumulate truncates (~3e-8 relative per add): a K = T*B contraction kept in one accumulator set drifts
// with its length (8e-6 at K = 196 608 with 32 splits, measured).  Splits for the time-parallel gradient GEMMs are
// therefore also bounded below so that one split covers at most kTcMaxKPerSplit of K (<= 64 adds per accumulator),
// within a 64 MB budget for the partials.
constexpr long long kTcMaxKPerSplit = 1536;
inline int tc_accuracy_splits(int M, int N, long long K, long long budget_bytes = 64LL << 20) {
  long long s = (K + kTcMaxKPerSplit - 1) / kTcMaxKPerSplit;
  const long long cap = budget_bytes / ((long long)M * N * 4);
  if (s > cap) s = cap;
  if (s > 128) s = 128;
  if (s < 1) s = 1;
  return (int)s;
}

struct GenericSizes {
  int zxp, zp;
  long long n_xp, n_dpre, n_dz, n_dzx, n_state, n_part;
  long long n_at, n_uxt;        // forward: A^T [RH, Hp], Ux^T [RX, Ip]   (B operands of the tensor-core GEMMs are K-major)
  long long n_bmt, n_vxt;       // backward: Bm^T [RH, 4H], Vx^T [RX, 4H]
  long long n_tcpart;           // split-K partials of the per-timestep tensor-core GEMMs
  long long ldt, n_tA, n_tB;    // transposed time-parallel operands: [4H, ldt] and [max(RH,RX), ldt], ldt = round_up(T*B, 4)
  int hp4, ip4;
};
inline GenericSizes generic_sizes(int T, int B, int I, int H, int RX, int RH) {
  GenericSizes s;
  s.zxp = round_up(RX, 4);
  s.zp = round_up(RH, 4);
  const long long rows = (long long)T * B;
  s.n_xp = rows * 4 * H;
  s.n_dpre = rows * 4 * H;
  s.n_dz = rows * s.zp;
  s.n_dzx = rows * s.zxp;
  s.n_state = (long long)B * H;
  long long p = (long long)(g_splits(4 * H, RH, rows) + 1) * 4 * H * RH;       // dBm
  long long q = (long long)(g_splits(H, RH, rows) + 1) * H * RH; if (q > p) p = q;   // dA (+1: h0 slice)
  q = (long long)g_splits(4 * H, RX, rows) * 4 * H * RX; if (q > p) p = q;     // dVx
  q = (long long)g_splits(I, RX, rows) * I * RX; if (q > p) p = q;             // dUx
  q = (long long)kColSplits * (8 * H + 4 * I); if (q > p) p = q;               // column reductions
  q = (long long)tc_accuracy_splits(4 * H, RH, rows) * 4 * H * RH; if (q > p) p = q;   // accuracy-driven splits (tensor-core path)
  q = (long long)tc_accuracy_splits(4 * H, RX, rows) * 4 * H * RX; if (q > p) p = q;
  q = (long long)tc_accuracy_splits(H, RH, rows) * H * RH; if (q > p) p = q;
  q = (long long)tc_accuracy_splits(I, RX, rows) * I * RX; if (q > p) p = q;
  {                                                            // fused dBm | dVx product
    const long long nv = round_up(RH, 32) + RX;
    q = (long long)tc_accuracy_splits(4 * H, (int)nv, rows, 128LL << 20) * 4 * H * nv; if (q > p) p = q;
  }
  s.n_part = p;
  s.hp4 = round_up(H, 4);
  s.ip4 = round_up(I, 4);
  s.n_at = (long long)RH * s.hp4;
  s.n_uxt = (long long)RX * s.ip4;
  s.n_bmt = (long long)RH * 4 * H;
  s.n_vxt = (long long)RX * 4 * H;
  s.n_tcpart = 16LL * B * (RH > H ? RH : H);
  s.ldt = (rows + 3) / 4 * 4;
  s.n_tA = 4LL * H * s.ldt;
  s.n_tB = (long long)(RH > RX ? RH : RX) * s.ldt;
  return s;
}

inline int generic_plan(int T, int B, int I, int H, int RX, int RH, vmlmf_plan* plan) {
  const GenericSizes s = generic_sizes(T, B, I, H, RX, RH);
  plan->path = VMLMF_PATH_G;
  plan->zx_pitch = s.zxp;
  plan->z_pitch = s.zp;
  plan->xp_cols = 4 * H;
  plan->fwd_workspace_bytes = (s.n_xp + 2 * s.n_state + (long long)B * s.zp + s.n_at + s.n_uxt + 2 * (long long)B * s.hp4 + s.n_tcpart + 32) * (long long)sizeof(float);
  plan->bwd_workspace_bytes = (s.n_dpre + s.n_dz + s.n_dzx + 2 * s.n_state + s.n_part + s.n_bmt + s.n_vxt + s.n_tcpart + s.n_tA + s.n_tB + 48) * (long long)sizeof(float);
  return VMLMF_OK;
}

// dst[c, r] = src[r, c]  (src [R, C] row-major, dst row pitch ldd >= R); tiny factors only
static __global__ void transpose_kernel(const float* __restrict__ src, int R, int Cc, float* __restrict__ dst, int ldd) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? src[(size_t)r * Cc + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < Cc && r < ldd) dst[(size_t)c * ldd + r] = (r < R) ? tile[threadIdx.x][i] : 0.f;
  }
}
// dst[c, r] = row_r[c] for r < rows, c < C, where row_r = head + r*head_ld for r < head_n (zeros when head is null)
// and src.row(r - head_n) otherwise; dst row pitch ldd (pad columns r >= rows are not written).
static __global__ void transpose_rows_kernel(RowView src, const float* __restrict__ head, long long head_ld, int head_n,
                                             long long rows, int C, float* __restrict__ dst, long long ldd) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long r = r0 + i;
    const int c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < rows && c < C) {
      if (r < head_n) v = head ? head[r * head_ld + c] : 0.f;
      else v = src.row(r - head_n)[c];
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long r = r0 + threadIdx.x;
    if (c < C && r < rows) dst[(size_t)c * ldd + r] = tile[threadIdx.x][i];
  }
}
inline int transpose_rows_launch(RowView src, const float* head, long long head_ld, int head_n, long long rows, int C,
                                 float* dst, long long ldd, cudaStream_t st) {
  dim3 grid((unsigned)((rows + 31) / 32), ceil_div(C, 32));
  transpose_rows_kernel<<<grid, dim3(32, 8), 0, st>>>(src, head, head_ld, head_n, rows, C, dst, ldd);
  return (int)cudaGetLastError();
}
inline int transpose_launch(const float* src, int R, int Cc, float* dst, int ldd, cudaStream_t st) {
  dim3 grid(ceil_div(Cc, 32), ceil_div(ldd, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(src, R, Cc, dst, ldd);
  return (int)cudaGetLastError();
}
// VMLMF_G_SIMT=1 keeps every GEMM of the generic regime on the SIMT kernel (A/B measurements, tests)
inline bool g_simt_only() {
  const char* e = getenv("VMLMF_G_SIMT");
  return e && e[0] == '1';
}
// VMLMF_TN_OFF=1 keeps the transposed-copy path of the time-parallel gradient GEMMs (A/B measurements, tests)
inline bool g_tn_off() {
  const char* e = getenv("VMLMF_TN_OFF");
  return e && e[0] == '1';
}
inline float* align4(float* p) { return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15); }

// out[M, N] (row pitch ldo) (+)= A[M,K] B[N,K]^T on the tensor cores, split over K when the tile grid alone would
// leave most SMs idle (per-timestep GEMMs: M = batch, N = rank).  `part` holds up to part_cap floats of partials.
inline int tc_gemm_rows(const float* A, long long lda, const float* Bm, long long ldb, int M, int N, int K, float* out,
                        long long ldo, int accumulate, float* part, long long part_cap, cudaStream_t st) {
  int splits = tc::tc_splits(M, N, K, 16);
  while (splits > 1 && (long long)splits * M * N > part_cap) --splits;
  if (splits <= 1) return tc::gemm_tc(A, lda, Bm, ldb, M, N, K, tc::EpiStoreTC{out, ldo, accumulate}, st);
  const int nkb = ceil_div(K, tc::BK), kbs = ceil_div(nkb, splits), nz = ceil_div(nkb, kbs);
  int rc = tc::gemm_tc(A, lda, Bm, ldb, M, N, K, tc::EpiPartialTC{part, M, N}, st, splits);
  if (rc) return rc;
  const long long n = (long long)M * N;
  splitk_reduce_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, nz, M, N, out, ldo, accumulate);
  return (int)cudaGetLastError();
}

struct EpiBias {            // SIMT fallback of EpiBiasTC
  RowView C; const float* bias; int beta;
  __device__ void operator()(int m, int n, int N, const float (&v)[4]) const {
    float* c = C.row(m) + n;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (n + q < N) {
        const float r = v[q] + (bias ? bias[n + q] : 0.f);
        c[q] = beta ? c[q] + r : r;
      }
  }
};
static __global__ void splitk_reduce_bias_kernel(const float* __restrict__ part, int nsplit, int M, int N,
                                                 const float* __restrict__ bias, float* __restrict__ out, long long ldo,
                                                 int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int m = (int)(i / N), c = (int)(i - (long long)m * N);
  float s = bias ? bias[c] : 0.f;
  for (int z = 0; z < nsplit; ++z) s += part[(size_t)z * M * N + i];
  float* o = out + (size_t)m * ldo + c;
  *o = accumulate ? *o + s : s;
}
// public GEMM (vmlmf_gemm_nt): tensor cores when the operands meet the TMA constraints, SIMT otherwise
inline int gemm_nt_public(const float* A, long long lda, const float* Bm, long long ldb, float* C, long long ldc,
                          const float* bias, int M, int N, int K, int accumulate, float* part, long long part_floats,
                          cudaStream_t st) {
  if (!g_simt_only() && tc::encode_fn() && tc::tc_operand_ok(A, lda) && tc::tc_operand_ok(Bm, ldb)) {
    int splits = tc::tc_splits(M, N, K, 16);
    const int acc_splits = tc_accuracy_splits(M, N, K);      // long contractions: bound the adds per accumulator
    if (splits < acc_splits) splits = acc_splits;
    while (splits > 1 && (!part || (long long)splits * M * N > part_floats)) --splits;
    if (splits <= 1) return tc::gemm_tc(A, lda, Bm, ldb, M, N, K, tc::EpiBiasTC{C, ldc, bias, accumulate}, st);
    const int nkb = ceil_div(K, tc::BK), kbs = ceil_div(nkb, splits), nz = ceil_div(nkb, kbs);
    int rc = tc::gemm_tc(A, lda, Bm, ldb, M, N, K, tc::EpiPartialTC{part, M, N}, st, splits);
    if (rc) return rc;
    const long long n = (long long)M * N;
    splitk_reduce_bias_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, nz, M, N, bias, C, ldc, accumulate);
    return (int)cudaGetLastError();
  }
  return gemm_launch<false, true>(plain_view(A, lda), plain_view(Bm, ldb), M, N, K, 1, NIdent{},
                                  EpiBias{plain_view(C, ldc), bias, accumulate}, st);
}

// public TN GEMM (vmlmf_gemm_tn): tensor cores only (the operands must meet the TMA constraints)
inline int gemm_tn_public(const float* At, long long lda, const float* Bt, long long ldb, float* C, long long ldc, int M, int N,
                          long long K, int accumulate, float* part, long long part_floats, cudaStream_t st) {
  if (!tc::encode_fn() || !tc::tc_operand_ok(At, lda) || !tc::tc_operand_ok(Bt, ldb) || K > 0x7fffffffLL) return VMLMF_EUNSUPPORTED;
  int splits = tc::tc_splits(M, N, (int)K, 32);
  const int acc_splits = tc_accuracy_splits(M, N, K);
  if (splits < acc_splits) splits = acc_splits;
  while (splits > 1 && (!part || (long long)splits * M * N > part_floats)) --splits;
  if (splits <= 1) {
    const int rc = tc::gemm_tn(At, lda, Bt, ldb, M, N, K, tc::EpiBiasTC{C, ldc, nullptr, accumulate}, st);
    return rc == tc::kTcNoFit ? VMLMF_EUNSUPPORTED : rc;
  }
  const int nkb = ceil_div((int)K, tc::BK), kbs = ceil_div(nkb, splits), nz = ceil_div(nkb, kbs);
  int rc = tc::gemm_tn(At, lda, Bt, ldb, M, N, K, tc::EpiPartialTC{part, M, N}, st, splits);
  if (rc) return rc == tc::kTcNoFit ? VMLMF_EUNSUPPORTED : rc;
  const long long n = (long long)M * N;
  splitk_reduce_bias_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, nz, M, N, nullptr, C, ldc, accumulate);
  return (int)cudaGetLastError();
}

// ZX = X Ux, pad columns zeroed
static __global__ void zero_pad_cols_kernel(float* z, long long rows, int pitch, int R) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int pad = pitch - R;
  if (pad <= 0 || i >= rows * pad) return;
  z[(i / pad) * pitch + R + (i % pad)] = 0.f;
}
inline int generic_xproj(const float* x, long long xs_t, long long xs_b, const float* Ux, float* zx, int T, int B,
                         int I, int RX, int zx_pitch, cudaStream_t st) {
  const long long rows = (long long)T * B;
  if (rows > 0x7fffffff) return VMLMF_EUNSUPPORTED;
  if (zx_pitch > RX) {
    const long long n = rows * (zx_pitch - RX);
    zero_pad_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(zx, rows, zx_pitch, RX);
  }
  return gemm_launch<false, false>(tb_view(x, xs_t, xs_b, B), plain_view(Ux, RX), (int)rows, RX, I, 1, NIdent{},
                                   EpiStore{plain_view(zx, zx_pitch), 0}, st);
}

#define G_TRY(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)

inline int generic_seq_fwd(const vmlmf_plan* plan, const float* x, long long xs_t, long long xs_b, const float* zx,
                           const float* Ux, const float* Vx, const float* Dx, const float* A, const float* Bm,
                           const float* Dh, const float* bias, const float* h0, const float* c0, float* y,
                           long long ys_t, long long ys_b, float* hT, float* cT, float* gates, float* cs, float* z,
                           void* workspace, int T, int B, int I, int H, int RX, int RH, cudaStream_t st) {
  (void)Ux;
  const GenericSizes s = generic_sizes(T, B, I, H, RX, RH);
  if (plan->zx_pitch != s.zxp || plan->z_pitch != s.zp) return VMLMF_EPLAN;
  if (!workspace) return VMLMF_EWORKSPACE;
  const long long rows = (long long)T * B;
  if (rows * 4 * H > (1LL << 40) || rows > 0x7fffffff) return VMLMF_EUNSUPPORTED;
  float* xp = (float*)workspace;
  float* cbuf[2] = {xp + s.n_xp, xp + s.n_xp + s.n_state};      // running c when not saving
  float* zscratch = align4(xp + s.n_xp + 2 * s.n_state);         // z_t when not saving  [B, zp], 16-byte aligned (TMA operand)
  const bool use_tc = !g_simt_only();
  float* at = align4(zscratch + (size_t)B * s.zp);               // A^T [RH, hp4] for the tensor-core z GEMM
  // time-parallel: XP = ZX Vx^T + bias + x (.) Dx   (tcgen05 3xTF32 GEMM; SIMT when an operand misses the TMA constraints)
  {
    int rc = tc::kTcNoFit;
    if (use_tc) rc = tc::gemm_tc(zx, s.zxp, Vx, RX, (int)rows, 4 * H, RX, tc::EpiXPTC{xp, bias, x, xs_t, xs_b, B, Dx, H, I}, st);
    if (rc == tc::kTcNoFit)
      rc = gemm_launch<false, true>(plain_view(zx, s.zxp), plain_view(Vx, RX), (int)rows, 4 * H, RX, 1, NIdent{},
                                    EpiXP{xp, bias, tb_view(x, xs_t, xs_b, B), Dx, H, I}, st);
    G_TRY(rc);
  }
  float* hpad[2] = {align4(at + s.n_at), nullptr};               // padded h_t ping-pong [B, hp4]
  hpad[1] = hpad[0] + (size_t)B * s.hp4;
  float* tcpart = align4(hpad[1] + (size_t)B * s.hp4);
  const bool tc_step = use_tc && tc::tc_operand_ok(Bm, RH) && (!z || tc::tc_operand_ok(z, s.zp)) && tc::encode_fn() != nullptr;
  if (tc_step) {
    G_TRY(transpose_launch(A, H, RH, at, s.hp4, st));
    if (h0) G_TRY((int)cudaMemcpy2DAsync(hpad[1], (size_t)s.hp4 * 4, h0, (size_t)H * 4, (size_t)H * 4, B, cudaMemcpyDeviceToDevice, st));
  }
  const bool save = gates != nullptr;
  for (int t = 0; t < T; ++t) {
    const float* hprev = t ? y + (size_t)(t - 1) * ys_t : h0;
    const long long hp_sb = t ? ys_b : H;
    const float* cprev = save ? (t ? cs + (size_t)(t - 1) * B * H : c0) : (t ? cbuf[(t - 1) & 1] : c0);
    float* cout = save ? cs + (size_t)t * B * H : cbuf[t & 1];
    float* zdst = save ? z + (size_t)t * B * s.zp : zscratch;   // z_t = h_{t-1} A (zero without initial state)
    if (hprev) {
      if (s.zp > RH) {
        const long long n = (long long)B * (s.zp - RH);
        zero_pad_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(zdst, B, s.zp, RH);
      }
      int rc = tc::kTcNoFit;
      if (tc_step) rc = tc_gemm_rows(hpad[(t + 1) & 1], s.hp4, at, s.hp4, B, RH, H, zdst, s.zp, 0, tcpart, s.n_tcpart, st);
      if (rc == tc::kTcNoFit) {
        if (tc_step) return VMLMF_EUNSUPPORTED;
        rc = gemm_launch<false, false>(RowView{const_cast<float*>(hprev), 0, hp_sb, 0x7fffffff}, plain_view(A, RH), B, RH, H,
                                       1, NIdent{}, EpiStore{plain_view(zdst, s.zp), 0}, st);
      }
      G_TRY(rc);
    } else {
      G_TRY((int)cudaMemsetAsync(zdst, 0, (size_t)B * s.zp * sizeof(float), st));
    }
    const bool last = (t == T - 1);
    EpiGate eg{xp + (size_t)t * B * 4 * H, hprev, hp_sb, cprev, Dh, y + (size_t)t * ys_t, ys_b, cout,
               save ? gates + (size_t)t * B * 4 * H : nullptr, last ? hT : nullptr, last ? cT : nullptr, H};
    int rc = tc::kTcNoFit;
    if (tc_step) {
      tc::EpiGate32 eg32{eg.xp_t, eg.hprev, eg.hp_sb, eg.cprev, eg.Dh, eg.y_t, eg.y_sb, eg.c_out, eg.gates_t, eg.hT, eg.cT, hpad[t & 1], s.hp4, H};
      rc = tc::gemm_tc(zdst, s.zp, Bm, RH, B, 4 * H, RH, eg32, st);       // B = [4][H][RH] view: tile = 4 gates x 32 units
    }
    if (rc == tc::kTcNoFit) {
      if (tc_step) return VMLMF_EUNSUPPORTED;            // the padded h copy is only maintained by the tensor-core epilogue
      rc = gemm_launch<false, true>(plain_view(zdst, s.zp), plain_view(Bm, RH), B, 4 * H, RH, 1, NGate{H}, eg, st);
    }
    G_TRY(rc);
  }
  return VMLMF_OK;
}

// ---- time-parallel half of the backward: every contraction over the T*B rows, shared by regimes G and R2 ----
//   dBm = dPre^T Z, dVx = dPre^T ZX, dA = Hprev^T dZ, dZX = dPre Vx (unless the caller already has it), dUx = X^T dZX,
//   dX = dZX Ux^T + sum_k dPre_k (.) Dx_k, column sums dbias / dDh / dDx.  dPre is [T*B, 4, G] (G >= H: regime R2 pads the
//   gate stride so that its TMA tiles never straddle a gate); pad columns must be zero.
struct TpArgs {
  const float* dpre; int G;
  const float* z; int zp; const float* zx; int zxp;
  float* dz; float* dzx; bool have_dzx;
  const float* x; long long xs_t, xs_b; const float* y; long long ys_t, ys_b; const float* h0;
  const float *Ux, *Vx, *Dx;
  float* dx; long long dxs_t, dxs_b;
  float *dUx, *dVx, *dDx, *dA, *dBm, *dDh, *dbias;
  float* part; long long n_part;       // split-K partials
  float* vxt;                          // Vx^T [RX, 4G] scratch (only when !have_dzx)
  float* tA; float* tB; float* gtmp;   // transposed operands [4G, ldt], [max(RH,RX), ldt]; [4G, max(RH,RX)] when G != H
  long long ldt; float* unused;
  int T, B, I, H, RX, RH; bool use_tc;
  bool vxt_padded = false;             // vxt already holds Vx^T in dPre's gate-padded layout [RX, 4G] (regime R3)
};
// dst[k*H + j, :] = src[k*G + j, :]  (rows of R floats)
static __global__ void compact_gate_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int G, int R) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4LL * H * R) return;
  const int c = (int)(i % R);
  const long long row = i / R;
  const int k = (int)(row / H), j = (int)(row % H);
  dst[i] = src[((size_t)k * G + j) * R + c];
}
// fused dBm | dVx product: part[z][k*G + j][c] over the virtual column axis [RH padded to N1p | RX]  ->  dBm[kH + j, RH], dVx[kH + j, RX]
// (fixed-order sum over the splits; the gate padding G -> H is dropped on the way)
static __global__ void splitk_reduce_pair_kernel(const float* __restrict__ part, int nsplit, int H, int G, int RH, int RX, int N1p,
                                                 float* __restrict__ dBm, float* __restrict__ dVx) {
  const int W = RH + RX, Nv = N1p + RX;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4LL * H * W) return;
  const int c = (int)(i % W);
  const long long row = i / W;
  const int k = (int)(row / H), j = (int)(row % H);
  const size_t src = ((size_t)k * G + j) * Nv + (c < RH ? c : N1p + (c - RH));
  const size_t zs = (size_t)4 * G * Nv;
  float s = 0.f;
  for (int z = 0; z < nsplit; ++z) s += part[(size_t)z * zs + src];
  if (c < RH) dBm[row * RH + c] = s; else dVx[row * RX + (c - RH)] = s;
}
// scratch of generic_bwd_tp for a gate stride G (floats): partials, transposed operands, gate-compaction buffer
struct TpScratch { long long n_part, ldt, n_tA, n_tB, n_gtmp, total; };
inline TpScratch tp_scratch(int T, int B, int I, int H, int G, int RX, int RH) {
  TpScratch s;
  const long long rows = (long long)T * B;
  long long p = (long long)(g_splits(4 * G, RH, rows) + 1) * 4 * G * RH;
  long long q = (long long)(g_splits(H, RH, rows) + 1) * H * RH; if (q > p) p = q;
  q = (long long)g_splits(4 * G, RX, rows) * 4 * G * RX; if (q > p) p = q;
  q = (long long)g_splits(I, RX, rows) * I * RX; if (q > p) p = q;
  q = (long long)kColSplits * (8 * H + 4 * I); if (q > p) p = q;
  q = (long long)tc_accuracy_splits(4 * G, RH, rows) * 4 * G * RH; if (q > p) p = q;
  q = (long long)tc_accuracy_splits(4 * G, RX, rows) * 4 * G * RX; if (q > p) p = q;
  q = (long long)tc_accuracy_splits(H, RH, rows) * H * RH; if (q > p) p = q;
  q = (long long)tc_accuracy_splits(I, RX, rows) * I * RX; if (q > p) p = q;
  {                                                            // fused dBm | dVx product (same budget as the two separate ones)
    const long long nv = round_up(RH, 32) + RX;
    q = (long long)tc_accuracy_splits(4 * G, (int)nv, rows, 128LL << 20) * 4 * G * nv; if (q > p) p = q;
  }
  s.n_part = p;
  s.ldt = (rows + 3) / 4 * 4;
  s.n_tA = 4LL * G * s.ldt;
  s.n_tB = (long long)(RH > RX ? RH : RX) * s.ldt;
  s.n_gtmp = G != H ? 4LL * G * (RH > RX ? RH : RX) : 0;
  s.total = s.n_part + s.n_tA + s.n_tB + s.n_gtmp + 64;
  return s;
}
inline int generic_bwd_tp(TpArgs& a, cudaStream_t st) {
  const int T = a.T, B = a.B, I = a.I, H = a.H, G = a.G, RX = a.RX, RH = a.RH;
  const long long rows = (long long)T * B;
  float* part = a.part;
  const RowView Xv = tb_view(a.x, a.xs_t, a.xs_b, B), Yv = tb_view(a.y, a.ys_t, a.ys_b, B);
  const RowView dPv = plain_view(a.dpre, 4 * G), Zv = plain_view(a.z, a.zp), ZXv = plain_view(a.zx, a.zxp);
  auto reduce_to = [&](int nsplit, long long n, float* out) -> int {
    splitk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, nsplit, n, out);
    return (int)cudaGetLastError();
  };
  // gate-padded [4G, R] result -> [4H, R]
  auto finish_gate = [&](float* tmp, float* out, int R) -> int {
    if (G == H) return 0;
    const long long n = 4LL * H * R;
    compact_gate_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tmp, out, H, G, R);
    return (int)cudaGetLastError();
  };
  float* tA = a.tA;
  float* tB = a.tB;
  const bool tc_tp = a.use_tc && rows >= 256 && tc::encode_fn() != nullptr;   // K = rows contractions on the tensor cores
  // C[M,N] = At[M, rows] Bt[N, rows]^T, split over K with a fixed-order reduce into `out`
  auto tc_tn = [&](int M, int N, float* out) -> int {
    int splits = tc::tc_splits(M, N, (int)rows, 32);
    const int acc_splits = tc_accuracy_splits(M, N, rows);
    if (splits < acc_splits) splits = acc_splits;
    while (splits > 1 && (long long)splits * M * N > a.n_part) --splits;
    const int nkb = ceil_div((int)rows, tc::BK), kbs = ceil_div(nkb, splits), nz = ceil_div(nkb, kbs);
    int rc = tc::gemm_tc(tA, a.ldt, tB, a.ldt, M, N, (int)rows, tc::EpiPartialTC{part, M, N}, st, splits);
    if (rc) return rc;
    return reduce_to(nz, (long long)M * N, out);
  };
  float* gBm = (G == H) ? a.dBm : a.gtmp;
  float* gVx = (G == H) ? a.dVx : a.gtmp;
  // C[M,N] = At[rows, M]^T Bt[rows, N] straight from the row-major activations (MN-major tensor-core operands)
  auto tc_direct = [&](const float* At, long long lda, const float* Bt, long long ldb, int M, int N, float* out) -> int {
    int splits = tc::tc_splits(M, N, (int)rows, 32);
    const int acc_splits = tc_accuracy_splits(M, N, rows);
    if (splits < acc_splits) splits = acc_splits;
    while (splits > 1 && (long long)splits * M * N > a.n_part) --splits;
    const int nkb = ceil_div((int)rows, tc::BK), kbs = ceil_div(nkb, splits), nz = ceil_div(nkb, kbs);
    int rc = tc::gemm_tn(At, lda, Bt, ldb, M, N, rows, tc::EpiPartialTC{part, M, N}, st, splits);
    if (rc) return rc;
    return reduce_to(nz, (long long)M * N, out);
  };
  const bool direct = tc_tp && !g_tn_off() && tc::tc_operand_ok(a.dpre, 4 * G) && tc::tc_operand_ok(a.z, a.zp) &&
                      tc::tc_operand_ok(a.zx, a.zxp);
  if (tc_tp) {
    if (direct) {
      // dBm = dPre^T Z, dVx = dPre^T ZX: no transposed copies, and ONE pass over dPre for both (B operand = [Z | ZX])
      const int N1p = round_up(RH, 32), Nv = N1p + RX, M4 = 4 * G;
      int splits = tc::tc_splits(M4, Nv, (int)rows, 32);
      const int acc_splits = tc_accuracy_splits(M4, Nv, rows, 128LL << 20);
      if (splits < acc_splits) splits = acc_splits;
      while (splits > 1 && (long long)splits * M4 * Nv > a.n_part) --splits;
      if ((long long)M4 * Nv <= a.n_part) {
        const int nkb = ceil_div((int)rows, tc::BK), kbs = ceil_div(nkb, splits), nz = ceil_div(nkb, kbs);
        G_TRY(tc::gemm_tn2(a.dpre, 4 * G, a.z, a.zp, RH, a.zx, a.zxp, RX, M4, rows, tc::EpiPartialTC{part, M4, Nv}, st, splits));
        const long long n = 4LL * H * (RH + RX);
        splitk_reduce_pair_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, nz, H, G, RH, RX, N1p, a.dBm, a.dVx);
        G_TRY((int)cudaGetLastError());
      } else {
        G_TRY(tc_direct(a.dpre, 4 * G, a.z, a.zp, 4 * G, RH, gBm));
        G_TRY(finish_gate(gBm, a.dBm, RH));
        G_TRY(tc_direct(a.dpre, 4 * G, a.zx, a.zxp, 4 * G, RX, gVx));
        G_TRY(finish_gate(gVx, a.dVx, RX));
      }
    } else {
      // dBm = dPre^T Z, dVx = dPre^T ZX share the transposed dPre
      G_TRY(transpose_rows_launch(dPv, nullptr, 0, 0, rows, 4 * G, tA, a.ldt, st));
      G_TRY(transpose_rows_launch(Zv, nullptr, 0, 0, rows, RH, tB, a.ldt, st));
      G_TRY(tc_tn(4 * G, RH, gBm));
      G_TRY(finish_gate(gBm, a.dBm, RH));
      G_TRY(transpose_rows_launch(ZXv, nullptr, 0, 0, rows, RX, tB, a.ldt, st));
      G_TRY(tc_tn(4 * G, RX, gVx));
      G_TRY(finish_gate(gVx, a.dVx, RX));
    }
    // dA = Hprev^T dZ: row (t,b) of Hprev is y[t-1,b], or h0[b] / 0 at t = 0
    int dA_rc = tc::kTcNoFit;
    if (direct && T > 1 && (H & 3) == 0 && tc::tc_operand_ok(a.dz, a.zp)) {
      // y read in place as the rank-3 MN-major operand (any layout with 16-byte row pitches): rows t >= 1 pair y[t-1] with
      // dz[t] -- a pointer offset of B rows on the dz side --, the t = 0 rows pair h0 with dz[0] in one more partial
      const long long rows1 = rows - B;
      int splits = tc::tc_splits(H, RH, (int)rows1, 32);
      const int acc_splits = tc_accuracy_splits(H, RH, rows1);
      if (splits < acc_splits) splits = acc_splits;
      while (splits > 1 && (long long)(splits + 1) * H * RH > a.n_part) --splits;
      const int nkb = ceil_div((int)rows1, tc::BK), kbs = ceil_div(nkb, splits);
      int nz = ceil_div(nkb, kbs);
      dA_rc = tc::gemm_tn_a3(a.y, a.ys_b, a.ys_t, B, T - 1, a.dz + (size_t)B * a.zp, a.zp, H, RH, tc::EpiPartialTC{part, H, RH}, st, splits);
      if (dA_rc == 0) {
        if (a.h0) {
          G_TRY((gemm_launch<true, false>(plain_view(a.h0, H), plain_view(a.dz, a.zp), H, RH, B, 1, NIdent{},
                                          EpiPartial{part + (size_t)nz * H * RH, H, RH}, st)));
          ++nz;
        }
        G_TRY(reduce_to(nz, (long long)H * RH, a.dA));
      } else if (dA_rc != tc::kTcNoFit) {
        return dA_rc;
      }
    }
    if (dA_rc == tc::kTcNoFit) {
      G_TRY(transpose_rows_launch(Yv, a.h0, H, B, rows, H, tA, a.ldt, st));
      G_TRY(transpose_rows_launch(plain_view(a.dz, a.zp), nullptr, 0, 0, rows, RH, tB, a.ldt, st));
      G_TRY(tc_tn(H, RH, a.dA));
    }
  } else {
    // dBm = dPre^T Z   [4G,RH]
    {
      const int sp = g_splits(4 * G, RH, rows);
      G_TRY((gemm_launch<true, false>(dPv, Zv, 4 * G, RH, rows, sp, NIdent{}, EpiPartial{part, 4 * G, RH}, st)));
      G_TRY(reduce_to(sp, (long long)4 * G * RH, gBm));
      G_TRY(finish_gate(gBm, a.dBm, RH));
    }
    // dA = Hprev^T dZ  [H,RH]: rows t>=1 pair y[t-1] with dz[t]; rows of t=0 pair h0 with dz[0]
    {
      const long long r1 = rows - B;
      int sp = 0;
      if (r1 > 0) {
        sp = g_splits(H, RH, r1);
        G_TRY((gemm_launch<true, false>(Yv, plain_view(a.dz + (size_t)B * a.zp, a.zp), H, RH, r1, sp, NIdent{},
                                        EpiPartial{part, H, RH}, st)));
      }
      if (a.h0) {
        G_TRY((gemm_launch<true, false>(plain_view(a.h0, H), plain_view(a.dz, a.zp), H, RH, B, 1, NIdent{},
                                        EpiPartial{part + (size_t)sp * H * RH, H, RH}, st)));
        ++sp;
      }
      if (sp == 0) G_TRY((int)cudaMemsetAsync(a.dA, 0, (size_t)H * RH * sizeof(float), st));
      else G_TRY(reduce_to(sp, (long long)H * RH, a.dA));
    }
    // dVx = dPre^T ZX  [4G,RX]
    {
      const int sp = g_splits(4 * G, RX, rows);
      G_TRY((gemm_launch<true, false>(dPv, ZXv, 4 * G, RX, rows, sp, NIdent{}, EpiPartial{part, 4 * G, RX}, st)));
      G_TRY(reduce_to(sp, (long long)4 * G * RX, gVx));
      G_TRY(finish_gate(gVx, a.dVx, RX));
    }
  }
  // dZX = dPre Vx    [T*B,RX]   (regime R2 forms it inside its recurrence kernel)
  if (!a.have_dzx) {
    if (a.zxp > RX) G_TRY((int)cudaMemsetAsync(a.dzx, 0, (size_t)rows * a.zxp * sizeof(float), st));
    int rc = tc::kTcNoFit;
    if (a.vxt_padded) {
      // pad units of dPre are zero, pad columns of vxt are zero: contract over the whole padded gate axis
      if (!(a.use_tc && tc::tc_operand_ok(a.dpre, 4 * G) && tc::tc_operand_ok(a.dzx, a.zxp))) return VMLMF_EUNSUPPORTED;
      // few output tiles, a long contraction (K = 4G): split K over the idle SMs, fixed-order reduce
      int splits = tc::tc_splits((int)rows, RX, 4 * G, 16);
      while (splits > 1 && (long long)splits * rows * RX > a.n_part) --splits;
      if (splits > 1) {
        const int nkb = ceil_div(4 * G, tc::BK), kbs = ceil_div(nkb, splits), nz = ceil_div(nkb, kbs);
        rc = tc::gemm_tc(a.dpre, 4 * G, a.vxt, 4 * G, (int)rows, RX, 4 * G, tc::EpiPartialTC{part, (int)rows, RX}, st, splits);
        if (rc == 0) {
          splitk_reduce_rows_kernel<<<(unsigned)((rows * RX + 255) / 256), 256, 0, st>>>(part, nz, (int)rows, RX, a.dzx, a.zxp, 0);
          rc = (int)cudaGetLastError();
        }
      } else {
        rc = tc::gemm_tc(a.dpre, 4 * G, a.vxt, 4 * G, (int)rows, RX, 4 * G, tc::EpiStoreTC{a.dzx, a.zxp, 0}, st);
      }
      if (rc == tc::kTcNoFit) return VMLMF_EUNSUPPORTED;
    } else if (a.use_tc && tc::tc_operand_ok(a.dpre, 4 * G) && tc::tc_operand_ok(a.dzx, a.zxp)) {
      G_TRY(transpose_launch(a.Vx, 4 * H, RX, a.vxt, 4 * H, st));
      rc = tc::gemm_tc(a.dpre, 4 * G, a.vxt, 4 * H, (int)rows, RX, 4 * H, tc::EpiStoreTC{a.dzx, a.zxp, 0}, st);
    }
    if (rc == tc::kTcNoFit)
      rc = gemm_launch<false, false>(dPv, plain_view(a.Vx, RX), (int)rows, RX, 4 * H, 1, NIdent{},
                                     EpiStore{plain_view(a.dzx, a.zxp), 0}, st);
    G_TRY(rc);
  }
  // dUx = X^T dZX    [I,RX]
  if (tc_tp) {
    G_TRY(transpose_rows_launch(Xv, nullptr, 0, 0, rows, I, tA, a.ldt, st));
    G_TRY(transpose_rows_launch(plain_view(a.dzx, a.zxp), nullptr, 0, 0, rows, RX, tB, a.ldt, st));
    G_TRY(tc_tn(I, RX, a.dUx));
  } else {
    const int sp = g_splits(I, RX, rows);
    G_TRY((gemm_launch<true, false>(Xv, plain_view(a.dzx, a.zxp), I, RX, rows, sp, NIdent{}, EpiPartial{part, I, RX}, st)));
    G_TRY(reduce_to(sp, (long long)I * RX, a.dUx));
  }
  // dX = dZX Ux^T + sum_k dPre[:, kG:kG+I] (.) Dx_k
  if (a.dx) {
    int rc = tc::kTcNoFit;
    if (a.use_tc && tc::tc_operand_ok(a.dzx, a.zxp) && tc::tc_operand_ok(a.Ux, RX))
      rc = tc::gemm_tc(a.dzx, a.zxp, a.Ux, RX, (int)rows, I, RX, EpiDXTC{a.dx, a.dxs_t, a.dxs_b, B, a.dpre, a.Dx, G, I}, st);
    if (rc == tc::kTcNoFit)
      rc = gemm_launch<false, true>(plain_view(a.dzx, a.zxp), plain_view(a.Ux, RX), (int)rows, I, RX, 1, NIdent{},
                                    EpiDX{tb_view(a.dx, a.dxs_t, a.dxs_b, B), a.dpre, a.Dx, G, I}, st);
    G_TRY(rc);
  }
  // dbias, dDh, dDx
  {
    long long rps = (rows + kColSplits - 1) / kColSplits;
    if (rps < 1) rps = 1;
    const int nsp = (int)((rows + rps - 1) / rps);
    ColRedArgs ca{a.dpre, Yv, a.h0, Xv, part, T, B, H, I, rps, G};
    const bool quad = (H & 3) == 0 && (G & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dpre) & 15) == 0 &&
                      (!a.h0 || (reinterpret_cast<uintptr_t>(a.h0) & 15) == 0) && ((a.ys_t | a.ys_b) & 3) == 0 &&
                      (reinterpret_cast<uintptr_t>(a.y) & 15) == 0;
    if (quad) colreduce4_kernel<<<dim3(ceil_div(H, 128), nsp), 128, 0, st>>>(ca);
    else colreduce_kernel<<<dim3(ceil_div(4 * H, 128), nsp), 128, 0, st>>>(ca);
    G_TRY((int)cudaGetLastError());
    colreduce_final_kernel<<<ceil_div(8 * H + 4 * I, 256), 256, 0, st>>>(part, nsp, H, I, a.dbias, a.dDh, a.dDx);
    G_TRY((int)cudaGetLastError());
  }
  return VMLMF_OK;
}

inline int generic_seq_bwd(const vmlmf_plan* plan, const float* x, long long xs_t, long long xs_b, const float* zx,
                           const float* Ux, const float* Vx, const float* Dx, const float* A, const float* Bm,
                           const float* Dh, const float* h0, const float* c0, const float* y, long long ys_t,
                           long long ys_b, const float* gates, const float* cs, const float* z, const float* dy,
                           long long dys_t, long long dys_b, const float* dhT, const float* dcT, float* dx,
                           long long dxs_t, long long dxs_b, float* dh0, float* dc0, float* dUx, float* dVx,
                           float* dDx, float* dA, float* dBm, float* dDh, float* dbias, void* workspace, int T,
                           int B, int I, int H, int RX, int RH, cudaStream_t st) {
  const GenericSizes s = generic_sizes(T, B, I, H, RX, RH);
  if (plan->zx_pitch != s.zxp || plan->z_pitch != s.zp) return VMLMF_EPLAN;
  if (!workspace) return VMLMF_EWORKSPACE;
  const long long rows = (long long)T * B;
  if (rows > 0x7fffffff) return VMLMF_EUNSUPPORTED;
  float* dpre = (float*)workspace;
  float* dz = dpre + s.n_dpre;
  float* dzx = dz + s.n_dz;
  float* dh = dzx + s.n_dzx;
  float* dc = dh + s.n_state;
  float* part = dc + s.n_state;
  const size_t sb = (size_t)B * H * sizeof(float);
  if (dhT) G_TRY((int)cudaMemcpyAsync(dh, dhT, sb, cudaMemcpyDeviceToDevice, st));
  else G_TRY((int)cudaMemsetAsync(dh, 0, sb, st));
  if (dcT) G_TRY((int)cudaMemcpyAsync(dc, dcT, sb, cudaMemcpyDeviceToDevice, st));
  else G_TRY((int)cudaMemsetAsync(dc, 0, sb, st));
  if (s.zp > RH) G_TRY((int)cudaMemsetAsync(dz, 0, (size_t)s.n_dz * sizeof(float), st));

  const bool use_tc = !g_simt_only() && tc::encode_fn() != nullptr;
  float* bmt = align4(part + s.n_part);                          // Bm^T [RH, 4H]
  float* vxt = align4(bmt + s.n_bmt);                            // Vx^T [RX, 4H]
  float* tcpart = align4(vxt + s.n_vxt);
  const bool tc_step = use_tc && tc::tc_operand_ok(dpre, 4 * H) && tc::tc_operand_ok(dz, s.zp) && tc::tc_operand_ok(A, RH);
  if (tc_step) G_TRY(transpose_launch(Bm, 4 * H, RH, bmt, 4 * H, st));
  const int nel = B * H;
  for (int t = T - 1; t >= 0; --t) {
    DpreArgs da{gates + (size_t)t * B * 4 * H, cs + (size_t)t * B * H, t ? cs + (size_t)(t - 1) * B * H : c0,
                dy ? dy + (size_t)t * dys_t : nullptr, dys_b, Dh, dh, dc, dpre + (size_t)t * B * 4 * H, B, H};
    dpre_step_kernel<<<ceil_div(nel, 256), 256, 0, st>>>(da);
    G_TRY((int)cudaGetLastError());
    float* dzt = dz + (size_t)t * B * s.zp;
    // dz_t = dPre_t Bm        [B,4H] x [4H,RH]
    int rc = tc::kTcNoFit;
    if (tc_step) rc = tc_gemm_rows(dpre + (size_t)t * B * 4 * H, 4 * H, bmt, 4 * H, B, RH, 4 * H, dzt, s.zp, 0, tcpart, s.n_tcpart, st);
    if (rc == tc::kTcNoFit)
      rc = gemm_launch<false, false>(plain_view(dpre + (size_t)t * B * 4 * H, 4 * H), plain_view(Bm, RH), B, RH, 4 * H, 1,
                                     NIdent{}, EpiStore{plain_view(dzt, s.zp), 0}, st);
    G_TRY(rc);
    // dh_{t-1} = (sum_k dpre_k Dh_k) + dz_t A^T
    rc = tc::kTcNoFit;
    if (tc_step) rc = tc_gemm_rows(dzt, s.zp, A, RH, B, H, RH, dh, H, 1, tcpart, s.n_tcpart, st);
    if (rc == tc::kTcNoFit)
      rc = gemm_launch<false, true>(plain_view(dzt, s.zp), plain_view(A, RH), B, H, RH, 1, NIdent{},
                                    EpiStore{plain_view(dh, H), 1}, st);
    G_TRY(rc);
  }
  if (dh0) G_TRY((int)cudaMemcpyAsync(dh0, dh, sb, cudaMemcpyDeviceToDevice, st));
  if (dc0) G_TRY((int)cudaMemcpyAsync(dc0, dc, sb, cudaMemcpyDeviceToDevice, st));

  TpArgs tp{dpre, H, z, s.zp, zx, s.zxp, dz, dzx, false, x, xs_t, xs_b, y, ys_t, ys_b, h0, Ux, Vx, Dx, dx, dxs_t, dxs_b,
            dUx, dVx, dDx, dA, dBm, dDh, dbias, part, s.n_part, vxt, align4(tcpart + s.n_tcpart), nullptr, nullptr, s.ldt, nullptr,
            T, B, I, H, RX, RH, use_tc};
  tp.tB = align4(tp.tA + s.n_tA);
  G_TRY(generic_bwd_tp(tp, st));
  return VMLMF_OK;
}

}  // namespace vmlmf
