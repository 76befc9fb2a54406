// xproj.cuh -- K1 (small-shape variant): zx[t,b,:] = x[t,b,:] Ux for all T*B rows at once, and the parameter
// packing kernels K0 / K5.  Time-parallel half of `torch.matmul(x, self.u_x)` (V/models/vmlmf.py:98,
// vmlmf_group.py:98, vmlmf_lm.py:246).  In regime R1 the contraction is [T*B, I<=512] x [I, RX<=128]: K and N are
// far below one tcgen05 tile and the op is bound by reading x once (I floats per row), so it is a shared-memory
// staged warp-MMA kernel (mma.sync m16n8k8, 3xTF32); the large-shape x projection (regime G) goes through the
// tcgen05 GEMM path instead.
#pragma once
#include "common.cuh"
#include "seq_mma.cuh"

namespace vmlmf {

// block = 128 threads; each thread produces 4 consecutive outputs of one row.
// smem: Us[I][pitch] | xs[ROWS][I+1]
// `order`: 0 = arbitrary (time, batch) strides, rows walked as t*B + b;  1 = x is one contiguous [B,T,I] block,
// 2 = one contiguous [T,B,I] block.  In the contiguous cases rows are walked in MEMORY order, so a block's 64 rows
// are a single contiguous span copied to shared memory with independent 16-byte loads (the 32-byte zx rows go wherever
// t*B + b says).  The product itself runs on mma.sync m16n8k8 with the 3xTF32 split (fp32-accurate): a warp owns 16
// rows, A fragments come from the staged rows, B fragments (Ux, split once per block) from shared memory -- a fifth of
// the instructions of the SIMT inner product, which leaves the kernel bound by reading x once.
// block = 128 threads = 4 warps x 16 rows.  smem: Uf[nk][NT][32] float4 {b0hi, b1hi, b0lo, b1lo} | xs[stages][64][I]
// (stages = 3 when the pipelined path fits, else 1)
constexpr int kXprojRows = 64;
static __global__ void __launch_bounds__(128) xproj_small_kernel(const float* __restrict__ x, long long xs_t,
                                                          long long xs_b, const float* __restrict__ Ux,
                                                          float* __restrict__ zx, int T, int B, int I,
                                                          int RX, int pitch, int order, int stages) {
  extern __shared__ __align__(16) float smem[];
  const int nk = (I + 7) >> 3, NT = (pitch + 7) >> 3;
  float4* Uf = reinterpret_cast<float4*>(smem);
  float* xs = smem + (size_t)nk * NT * 32 * 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  for (int i = tid; i < nk * NT * 32; i += blockDim.x) {
    const int ln = i & 31, f = i >> 5, nt = f % NT, ks = f / NT;
    const int k0 = 8 * ks + (ln & 3), k1 = k0 + 4, n = 8 * nt + (ln >> 2);
    const float b0 = (k0 < I && n < RX) ? __ldg(Ux + (size_t)k0 * RX + n) : 0.f;
    const float b1 = (k1 < I && n < RX) ? __ldg(Ux + (size_t)k1 * RX + n) : 0.f;
    const float b0h = tf32_rna(b0), b1h = tf32_rna(b1);
    Uf[i] = make_float4(b0h, b1h, b0 - b0h, b1 - b1h);
  }
  const long long nrows = (long long)T * B;
  const bool vec = order != 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const size_t stage_floats = (size_t)kXprojRows * I;

  // C[16 x 8 NT] = A[16 x 8 nk] B for this warp's 16 rows of the tile staged at xt
  auto compute = [&](const float* xt, long long row0, int nr) {
    const int r0 = 16 * warp + g, r1 = r0 + 8;
    const bool v0 = r0 < nr, v1 = r1 < nr;
    const float* x0 = xt + r0 * I;
    const float* x1 = xt + r1 * I;
    for (int nt = 0; nt < NT; ++nt) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const float4* uf = Uf + (size_t)nt * 32 + lane;
#pragma unroll 2
      for (int ks = 0; ks < nk; ++ks) {
        const int k0 = 8 * ks + q, k1 = k0 + 4;
        const float av[4] = {(v0 && k0 < I) ? x0[k0] : 0.f, (v1 && k0 < I) ? x1[k0] : 0.f,
                             (v0 && k1 < I) ? x0[k1] : 0.f, (v1 && k1 < I) ? x1[k1] : 0.f};
        float ah[4], al[4];
        split4(av, ah, al);
        const float4 b = uf[(size_t)ks * NT * 32];
        mma_3x(acc, ah, al, b.x, b.y, b.z, b.w);
      }
      // accumulator fragment: (row g, cols 2q, 2q+1), (row g+8, same cols) of n-tile nt
      const int c = 8 * nt + 2 * q;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        if (!(hf ? v1 : v0) || c >= pitch) continue;
        long long row = row0 + (hf ? r1 : r0);
        if (order == 1) {                                   // memory order is (b, t): zx row is t*B + b
          if (nrows <= 0xffffffffLL) {                      // 32-bit division: a 64-bit one costs ~100 instructions
            const unsigned r32 = (unsigned)row, bq = r32 / (unsigned)T;
            row = (long long)(r32 - bq * (unsigned)T) * B + bq;
          } else {
            row = (row % T) * B + row / T;
          }
        }
        float* o = zx + row * pitch + c;                    // pitch % 4 == 0 and c even: 8-byte aligned
        *reinterpret_cast<float2*>(o) = make_float2(acc[2 * hf], acc[2 * hf + 1]);
      }
    }
  };

  const long long step = (long long)gridDim.x * kXprojRows;
  if (vec && stages >= 3) {
    // contiguous, 16-byte aligned input: three tiles in flight per block (cp.async.cg 16 B), loads overlap the MMAs
    auto issue = [&](long long row0, int stage) {
      if (row0 < nrows) {
        const long long left = nrows - row0;
        const int n = (int)(left < kXprojRows ? left : kXprojRows) * I, n4 = n & ~3;
        const float* src = x + row0 * I;
        float* dst = xs + (size_t)stage * stage_floats;
        for (int e = tid * 4; e < n4; e += 4 * 128)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + e)), "l"(src + e) : "memory");
        for (int t1 = n4 + tid; t1 < n; t1 += 128) dst[t1] = __ldg(src + t1);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");   // always: keeps the group count uniform
    };
    long long row0 = (long long)blockIdx.x * kXprojRows;
    issue(row0, 0);
    issue(row0 + step, 1);
    for (int i = 0; row0 < nrows; ++i, row0 += step) {
      issue(row0 + 2 * step, (i + 2) % 3);              // that buffer was consumed in iteration i-1 (barrier below)
      asm volatile("cp.async.wait_group 2;" ::: "memory");
      __syncthreads();
      const long long left = nrows - row0;
      compute(xs + (size_t)(i % 3) * stage_floats, row0, (int)(left < kXprojRows ? left : kXprojRows));
      __syncthreads();
    }
    return;
  }
  for (long long row0 = (long long)blockIdx.x * kXprojRows; row0 < nrows; row0 += step) {
    __syncthreads();
    const long long left = nrows - row0;
    const int nr = (int)(left < kXprojRows ? left : kXprojRows);
    if (order != 0) {
      const float* src = x + row0 * I;
      for (int t1 = tid; t1 < nr * I; t1 += blockDim.x) xs[t1] = __ldg(src + t1);
    } else {
      for (int rr = warp; rr < nr; rr += NW) {
        const long long row = row0 + rr;
        long long t, b;
        if (nrows <= 0xffffffffLL) { const unsigned tq = (unsigned)row / (unsigned)B; t = tq; b = (unsigned)row - tq * (unsigned)B; }
        else { t = row / B; b = row % B; }
        const float* src = x + t * xs_t + b * xs_b;
        for (int jj = lane; jj < I; jj += 32) xs[rr * I + jj] = __ldg(src + jj);
      }
    }
    __syncthreads();
    compute(xs, row0, nr);
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
