// xproj.cuh -- K1 (small-shape variant): zx[t,b,:] = x[t,b,:] Ux for all T*B rows at once.
// Time-parallel half of `torch.matmul(x, self.u_x)` (V/models/vmlmf.py:98, vmlmf_group.py:98,
// vmlmf_lm.py:246).  In regime R1 the contraction is [T*B, I<=256] x [I, RX<=16]: K and N are
// far below one tcgen05 tile, the op is bound by reading x once (I floats/row) and SIMT FMAs,
// so it is a shared-memory tiled SIMT kernel; the large-shape x projection (regime G) goes
// through the GEMM path instead.
#pragma once
#include "common.cuh"

namespace vmlmf {

// block = 128 threads; each thread produces 4 consecutive outputs of one row.
// smem: Us[I][pitch] | xs[ROWS][I+1]
// `order`: 0 = arbitrary (time, batch) strides, rows walked as t*B + b;  1 = x is one contiguous [B,T,I] block,
// 2 = one contiguous [T,B,I] block.  In the contiguous cases rows are walked in MEMORY order, so a block's ROWS rows
// are a single contiguous span streamed with independent 16-byte loads (the 32-byte zx rows go wherever t*B + b says).
static __global__ void __launch_bounds__(128) xproj_small_kernel(const float* __restrict__ x, long long xs_t,
                                                          long long xs_b, const float* __restrict__ Ux,
                                                          float* __restrict__ zx, int T, int B, int I,
                                                          int RX, int pitch, int order) {
  extern __shared__ __align__(16) float smem[];
  const int G = pitch >> 2;                 // threads per row
  const int ROWS = blockDim.x / G;          // rows per block (a multiple of 4)
  float* Us = smem;                         // [I][pitch], zero padded
  float* xs = smem + I * pitch;             // [ROWS][I+1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
  for (int i = tid; i < I * pitch; i += blockDim.x) {
    const int jj = i / pitch, r = i % pitch;
    Us[i] = r < RX ? __ldg(Ux + (size_t)jj * RX + r) : 0.f;
  }
  const long long nrows = (long long)T * B;
  const bool vec = order != 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  for (long long row0 = (long long)blockIdx.x * ROWS; row0 < nrows; row0 += (long long)gridDim.x * ROWS) {
    __syncthreads();
    if (order != 0) {
      const float* src = x + row0 * I;
      const long long left = nrows - row0;
      const int n = (int)(left < ROWS ? left : ROWS) * I;          // floats in this block's span
      int e = tid * 4;
      if (vec) {
        const int n4 = n & ~3;
        for (; e + 3 * 512 < n4; e += 4 * 512) {                    // four independent 16-byte loads per thread
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(src + e + u * 512));
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e0 = e + u * 512;
            int rr = e0 / I, jj = e0 - rr * I;
            const float w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              xs[rr * (I + 1) + jj] = w[c];
              if (++jj == I) { jj = 0; ++rr; }
            }
          }
        }
        for (; e < n4; e += 512) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src + e));
          int rr = e / I, jj = e - rr * I;
          const float w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            xs[rr * (I + 1) + jj] = w[c];
            if (++jj == I) { jj = 0; ++rr; }
          }
        }
        for (int t1 = n4 + tid; t1 < n; t1 += blockDim.x) xs[(t1 / I) * (I + 1) + t1 % I] = __ldg(src + t1);
      } else {
        for (int t1 = tid; t1 < n; t1 += blockDim.x) xs[(t1 / I) * (I + 1) + t1 % I] = __ldg(src + t1);
      }
    } else {
      for (int rr = warp; rr < ROWS; rr += NW) {
        const long long row = row0 + rr;
        if (row < nrows) {
          const long long t = row / B, b = row % B;
          const float* src = x + t * xs_t + b * xs_b;
          for (int jj = lane; jj < I; jj += 32) xs[rr * (I + 1) + jj] = __ldg(src + jj);
        }
      }
    }
    __syncthreads();
    const int rr = tid / G, rg = tid % G;
    long long row = row0 + rr;
    if (rr < ROWS && row < nrows) {
      if (order == 1) row = (row % T) * B + row / T;      // memory order is (b, t): zx row is t*B + b
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* xr = xs + rr * (I + 1);
      for (int jj = 0; jj < I; ++jj) {
        const float xv = xr[jj];
        const float4 u = *reinterpret_cast<const float4*>(Us + jj * pitch + 4 * rg);
        acc.x = fmaf(xv, u.x, acc.x); acc.y = fmaf(xv, u.y, acc.y);
        acc.z = fmaf(xv, u.z, acc.z); acc.w = fmaf(xv, u.w, acc.w);
      }
      *reinterpret_cast<float4*>(zx + row * pitch + 4 * rg) = acc;
    }
  }
}

// ---- K0 / K5: diagonal corrections of the vector-multiplication terms and their chain rule ----
// one warp per (gate k, row j): D[k,j] = dia[j] - <u[j,:], v[kH+j,:]>
static __global__ void diag_fwd_kernel(const float* __restrict__ u, const float* __restrict__ v,
                                       const float* __restrict__ dia, float* __restrict__ D, int n, int H, int R) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= 4 * n) return;
  const int k = w / n, j = w - k * n;
  const float* ur = u + (size_t)j * R;
  const float* vr = v + ((size_t)k * H + j) * R;
  float s = 0.f;
  for (int r = lane; r < R; r += 32) s = fmaf(__ldg(ur + r), __ldg(vr + r), s);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) D[w] = __ldg(dia + j) - s;
}
// one thread per element of dv [4H, R]; the first n*R threads also produce du, the first n ddia
static __global__ void diag_bwd_kernel(const float* __restrict__ u, const float* __restrict__ v,
                                       const float* __restrict__ dD, float* __restrict__ du, float* __restrict__ dv,
                                       float* __restrict__ ddia, int n, int H, int R) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)4 * H * R) return;
  const int row = (int)(i / R), r = (int)(i - (long long)row * R);
  const int k = row / H, j = row - k * H;
  dv[i] = j < n ? -__ldg(dD + k * n + j) * __ldg(u + (size_t)j * R + r) : 0.f;
  if (i < (long long)n * R) {                 // here row = j < n <= H (k == 0), column r
    float s = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) s = fmaf(__ldg(dD + kk * n + row), __ldg(v + ((size_t)kk * H + row) * R + r), s);
    du[i] = -s;
  }
  if (i < n) ddia[i] = __ldg(dD + i) + __ldg(dD + n + i) + __ldg(dD + 2 * n + i) + __ldg(dD + 3 * n + i);
}


// ---- K0 / K5 for a whole plain cell in one launch each way (vmlmf_pack_plain_fwd / _bwd) ----
// forward: warps [0,4I) -> Dx, warps [4I, 4I+4H) -> Dh, the remaining warps -> bias = b_x + b_h (32 elements each)
static __global__ void pack_plain_fwd_kernel(const float* __restrict__ u_x, const float* __restrict__ v_x,
                                             const float* __restrict__ dia_x, const float* __restrict__ u_h,
                                             const float* __restrict__ v_h, const float* __restrict__ dia_h,
                                             const float* __restrict__ b_x, const float* __restrict__ b_h,
                                             float* __restrict__ Dx, float* __restrict__ Dh, float* __restrict__ bias,
                                             int I, int H, int RX, int RH) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= 4 * I + 4 * H) {
    const int i = (w - 4 * I - 4 * H) * 32 + lane;
    if (i < 4 * H) bias[i] = __ldg(b_x + i) + __ldg(b_h + i);
    return;
  }
  const bool xside = w < 4 * I;
  if (!xside) w -= 4 * I;
  const int n = xside ? I : H, R = xside ? RX : RH;
  const float* u = xside ? u_x : u_h;
  const float* v = xside ? v_x : v_h;
  const int k = w / n, j = w - k * n;
  const float* ur = u + (size_t)j * R;
  const float* vr = v + ((size_t)k * H + j) * R;
  float s = 0.f;
  for (int r = lane; r < R; r += 32) s = fmaf(__ldg(ur + r), __ldg(vr + r), s);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) (xside ? Dx : Dh)[w] = __ldg((xside ? dia_x : dia_h) + j) - s;
}

// backward, in place on the canonical gradients: one thread per element of dVx [4H,RX], then of dBm [4H,RH], then of
// db_h [4H].   dv[kH+j,r] -= dD[k,j] u[j,r] (j < n);  du[j,r] -= sum_k dD[k,j] v[kH+j,r];  ddia[j] = sum_k dD[k,j]
static __global__ void pack_plain_bwd_kernel(const float* __restrict__ u_x, const float* __restrict__ v_x,
                                             const float* __restrict__ u_h, const float* __restrict__ v_h,
                                             const float* __restrict__ dDx, const float* __restrict__ dDh,
                                             float* __restrict__ dUx, float* __restrict__ dVx, float* __restrict__ dA,
                                             float* __restrict__ dBm, float* __restrict__ ddia_x,
                                             float* __restrict__ ddia_h, const float* __restrict__ dbias,
                                             float* __restrict__ db_h, int I, int H, int RX, int RH) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nx = (long long)4 * H * RX, nh = (long long)4 * H * RH;
  if (i >= nx + nh) {
    i -= nx + nh;
    if (i < 4 * H) db_h[i] = __ldg(dbias + i);
    return;
  }
  const bool xside = i < nx;
  if (!xside) i -= nx;
  const int n = xside ? I : H, R = xside ? RX : RH;
  const float* u = xside ? u_x : u_h;
  const float* v = xside ? v_x : v_h;
  const float* dD = xside ? dDx : dDh;
  float* du = xside ? dUx : dA;
  float* dv = xside ? dVx : dBm;
  float* ddia = xside ? ddia_x : ddia_h;
  const int row = (int)(i / R), r = (int)(i - (long long)row * R);
  const int k = row / H, j = row - k * H;
  if (j < n) dv[i] -= __ldg(dD + k * n + j) * __ldg(u + (size_t)j * R + r);
  if (i < (long long)n * R) {                 // here row = j < n <= H (k == 0), column r
    float s = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) s = fmaf(__ldg(dD + kk * n + row), __ldg(v + ((size_t)kk * H + row) * R + r), s);
    du[i] -= s;
  }
  if (i < n) ddia[i] = __ldg(dD + i) + __ldg(dD + n + i) + __ldg(dD + 2 * n + i) + __ldg(dD + 3 * n + i);
}

}  // namespace vmlmf
