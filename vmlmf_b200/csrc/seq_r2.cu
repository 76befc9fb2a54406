// seq_r2.cu -- host side of regime R2 (seq_r2.cuh): geometry, operand packing, tensor maps, launches.
#include "seq_r2.cuh"

#include <stdlib.h>

#include "../../include/vmlmf_b200.h"

#include "seq_r2_host.cuh"

namespace vmlmf {
namespace r2 {

namespace {

inline long long al64(long long n) { return (n + 63) / 64 * 64; }      // 256-byte granules (TMA bases need 16)

// ---- operand packing (once per call; the factors are tiny next to the activations) ----
// at[r, j]   = A[j, r]                          [RHr, Hp]   (B operand of phase Z: N = z column r, K = unit j)
// w2[k,j,q]  = Bm[kH+j, q]        q <  RH       [4, Hp, KPp] (B operand of phase G: N = (gate k, unit j), K = [z | zx])
//            = Vx[kH+j, q-KZP]    KZP <= q < KZP+RX
// zero elsewhere; every value split into tf32 hi and lo parts.
__global__ void pack_fwd_kernel(const float* __restrict__ A, const float* __restrict__ Bm, const float* __restrict__ Vx,
                                float* __restrict__ at_hi, float* __restrict__ at_lo, float* __restrict__ w2_hi,
                                float* __restrict__ w2_lo, int H, int RH, int RX, int Hp, int RHr, int KZP, int KPp) {
  const long long n_at = (long long)RHr * Hp, n_w2 = 4LL * Hp * KPp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_at + n_w2; i += (long long)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < n_at) {
      const int r = (int)(i / Hp), j = (int)(i % Hp);
      if (r < RH && j < H) v = __ldg(A + (size_t)j * RH + r);
      const float hi = split_hi(v);
      at_hi[i] = hi;
      at_lo[i] = split_lo(v, hi);
    } else {
      const long long e = i - n_at;
      const int q = (int)(e % KPp);
      const int j = (int)((e / KPp) % Hp), k = (int)(e / ((long long)KPp * Hp));
      if (j < H) {
        if (q < RH) v = __ldg(Bm + ((size_t)k * H + j) * RH + q);
        else if (q >= KZP && q < KZP + RX) v = __ldg(Vx + ((size_t)k * H + j) * RX + (q - KZP));
      }
      const float hi = split_hi(v);
      w2_hi[e] = hi;
      w2_lo[e] = split_lo(v, hi);
    }
  }
}

// backward operands:
// w2t[n, k, j] = Bm[kH+j, n]        n <  RH       [KPp, 4, Hp]  (B operand of phase 1: N = column n of [dz | dzx], K = (gate, unit))
//              = Vx[kH+j, n-KZP]    KZP <= n < KZP+RX
// ap[j, q]     = A[j, q]                          [Hp, KZP]     (B operand of phase 2: N = unit j, K = z column q)
__global__ void pack_bwd_kernel(const float* __restrict__ A, const float* __restrict__ Bm, const float* __restrict__ Vx,
                                float* __restrict__ w2t_hi, float* __restrict__ w2t_lo, float* __restrict__ ap_hi,
                                float* __restrict__ ap_lo, int H, int RH, int RX, int Hp, int KZP, int KPp) {
  const long long n_w = (long long)KPp * 4 * Hp, n_a = (long long)Hp * KZP;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_w + n_a; i += (long long)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < n_w) {
      const int j = (int)(i % Hp), k = (int)((i / Hp) & 3), n = (int)(i / (4LL * Hp));
      if (j < H) {
        if (n < RH) v = __ldg(Bm + ((size_t)k * H + j) * RH + n);
        else if (n >= KZP && n < KZP + RX) v = __ldg(Vx + ((size_t)k * H + j) * RX + (n - KZP));
      }
      const float hi = split_hi(v);
      w2t_hi[i] = hi;
      w2t_lo[i] = split_lo(v, hi);
    } else {
      const long long e = i - n_w;
      const int q = (int)(e % KZP), j = (int)(e / KZP);
      if (j < H && q < RH) v = __ldg(A + (size_t)j * RH + q);
      const float hi = split_hi(v);
      ap_hi[e] = hi;
      ap_lo[e] = split_lo(v, hi);
    }
  }
}

// hop[b, j] = h0[b, j] (0 when h0 is null or j >= H), as tf32 hi / lo
__global__ void prep_state_kernel(const float* __restrict__ h0, float* __restrict__ hop_hi, float* __restrict__ hop_lo,
                                  int B, int H, int Hp) {
  const long long n = (long long)B * Hp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / Hp), j = (int)(i % Hp);
    const float v = (h0 && j < H) ? h0[(size_t)b * H + j] : 0.f;
    const float hi = split_hi(v);
    hop_hi[i] = hi;
    hop_lo[i] = split_lo(v, hi);
  }
}

// elementwise tf32 split of a dense buffer
__global__ void split_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = src[i], h = split_hi(v);
    hi[i] = h;
    lo[i] = split_lo(v, h);
  }
}

inline int ew_grid(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = 8LL * num_sms();
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

// ---- tensor maps ----
int map_2d(CUtensorMap* m, const float* p, long long cols, long long rows, long long ld) {
  return tc::make_map_2d(m, p, rows, cols, ld);                       // box 32 x 128
}
// [d3][d2][d1][cols] view, strides in floats; box = 32 x 1 x box2 x 1
int map_4d(CUtensorMap* m, const float* p, long long cols, long long d1, long long d2, long long d3, long long s1, long long s2,
           long long s3, int box2) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  if (!fn) return tc::kTcNoFit;
  cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)d1, (cuuint64_t)d2, (cuuint64_t)d3};
  cuuint64_t strides[3] = {(cuuint64_t)s1 * 4, (cuuint64_t)s2 * 4, (cuuint64_t)s3 * 4};
  cuuint32_t box[4] = {BK, 1, (cuuint32_t)box2, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : tc::kTcNoFit;
}
// [d2][d1][cols] view, strides in floats; box = 32 x box1 x box2
int map_3d(CUtensorMap* m, const float* p, long long cols, long long d1, long long d2, long long s1, long long s2,
           int box1, int box2) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  if (!fn) return tc::kTcNoFit;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)s1 * 4, (cuuint64_t)s2 * 4};
  cuuint32_t box[3] = {BK, (cuuint32_t)box1, (cuuint32_t)box2};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : tc::kTcNoFit;
}

// Cooperative launch: every CTA of the grid is resident at once (grid <= SMs, one CTA per SM by shared-memory size), which
// the group barriers inside the kernels rely on.
template <class Kern, class... Args>
int launch_coop(Kern kern, int grid, int smem_bytes, cudaStream_t st, Args... args) {
  // set on every launch: kernels that differ only in a template flag share this function's instantiation (same pointer
  // type), and the attribute is per device; the call is a few microseconds against a launch that runs T timesteps
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return (int)e;
  void* argv[] = {(void*)&args...};
  return (int)cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kThreads), argv, smem_bytes, st);
}

}  // namespace

// VMLMF_NO_R2=1 keeps the launch-per-timestep generic regime (A/B measurements, tests); read at plan time
bool r2_disabled() {
  const char* e = getenv("VMLMF_NO_R2");
  return e && e[0] == '1';
}

Geom geom(int T, int B, int I, int H, int RX, int RH) {
  (void)I;
  Geom g;
  g.ntiles = ceil_div(B, BM);
  // widest group that still has work for every CTA: enough CTAs to cover the SMs, at least 32 units per CTA
  int cs = num_sms() / g.ntiles;
  if (cs > kMaxGroup) cs = kMaxGroup;
  if (cs < 1) cs = 1;
  if (cs > ceil_div(H, 32)) cs = ceil_div(H, 32);
  const char* force = getenv("VMLMF_R2_CLUSTER");     // tests: force a group size
  if (force && atoi(force) >= 1 && atoi(force) <= kMaxGroup && atoi(force) <= ceil_div(H, 32)) cs = atoi(force);
  // at most 512 units (64 tf32 k-steps) per CTA in the K-split z product: the tensor core's accumulate truncates, so
  // the rounding error of one accumulator grows with the number of k-steps (measured 3e-8 per step)
  if (cs < ceil_div(H, 512)) cs = ceil_div(H, 512);
  g.HS = 32 * ceil_div(H, 32 * cs);
  g.CS = ceil_div(H, g.HS);
  g.Hp = g.CS * g.HS;
  g.zp = round_up(RH, 4);
  g.zxp = round_up(RX, 4);
  g.RHr = round_up(RH, 8);
  g.KZP = round_up(RH, 32);
  g.KXP = round_up(RX, 32);
  g.KPp = g.KZP + g.KXP;
  int ncl = num_sms() / g.CS;
  if (ncl > g.ntiles) ncl = g.ntiles;
  if (ncl < 1) ncl = 1;
  g.ncl = ncl;
  // forward workspace (floats)
  long long o = 0;
  g.o_hop_hi = o; o += al64((long long)B * g.Hp);
  g.o_hop_lo = o; o += al64((long long)B * g.Hp);
  g.o_zop_hi = o; o += al64((long long)B * g.zp);
  g.o_zop_lo = o; o += al64((long long)B * g.zp);
  g.o_zpart = o; o += al64((long long)g.ncl * g.CS * BM * g.zp);
  g.o_zx_hi = o; o += al64((long long)T * B * g.zxp);
  g.o_zx_lo = o; o += al64((long long)T * B * g.zxp);
  g.o_at_hi = o; o += al64((long long)g.RHr * g.Hp);
  g.o_at_lo = o; o += al64((long long)g.RHr * g.Hp);
  g.o_w2_hi = o; o += al64(4LL * g.Hp * g.KPp);
  g.o_w2_lo = o; o += al64(4LL * g.Hp * g.KPp);
  g.o_cbuf = o; o += al64(2LL * B * H);
  g.o_sync = o; o += al64(32LL * g.ncl);
  g.fwd_floats = o + 64;
  // backward: phase 1 contracts over 4 * HS values per CTA; at most 16 K tiles (64 tf32 k-steps) per accumulator
  g.KSPLIT = ceil_div(4 * (g.HS / 32), 16);
  g.NP = g.CS * g.KSPLIT;
  o = 0;
  g.b_dpre = o; o += al64((long long)T * B * 4 * g.Hp);
  g.b_dz = o; o += al64((long long)T * B * g.zp);
  g.b_dzx = o; o += al64((long long)T * B * g.zxp);
  g.b_dpo_hi = o; o += al64((long long)B * 4 * g.Hp);
  g.b_dpo_lo = o; o += al64((long long)B * 4 * g.Hp);
  g.b_dzo_hi = o; o += al64((long long)B * g.zp);
  g.b_dzo_lo = o; o += al64((long long)B * g.zp);
  g.b_dhrun = o; o += al64((long long)B * g.Hp);
  g.b_dcrun = o; o += al64((long long)B * g.Hp);
  g.b_part = o; o += al64(g.NP > 1 ? (long long)g.ncl * g.NP * BM * g.KPp : 0);
  g.b_w2t_hi = o; o += al64((long long)g.KPp * 4 * g.Hp);
  g.b_w2t_lo = o; o += al64((long long)g.KPp * 4 * g.Hp);
  g.b_ap_hi = o; o += al64((long long)g.Hp * g.KZP);
  g.b_ap_lo = o; o += al64((long long)g.Hp * g.KZP);
  g.b_sync = o; o += al64(32LL * g.ncl);
  g.bwd_floats = o + 64;
  return g;
}

bool fits(int T, int B, int I, int H, int RX, int RH) {
  if (r2_disabled()) return false;
  if (I > H || (long long)T * B > 0x7fffffffLL) return false;
  if (tc::encode_fn() == nullptr) return false;
  return true;
}

int launch_fwd(const FwdCall& c, void* workspace, cudaStream_t st) {
  const Geom g = geom(c.T, c.B, c.I, c.H, c.RX, c.RH);
  float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  float *hop_hi = ws + g.o_hop_hi, *hop_lo = ws + g.o_hop_lo, *zop_hi = ws + g.o_zop_hi, *zop_lo = ws + g.o_zop_lo;
  float *zx_hi = ws + g.o_zx_hi, *zx_lo = ws + g.o_zx_lo, *at_hi = ws + g.o_at_hi, *at_lo = ws + g.o_at_lo;
  float *w2_hi = ws + g.o_w2_hi, *w2_lo = ws + g.o_w2_lo;
  const bool save = c.gates != nullptr;

  pack_fwd_kernel<<<ew_grid((long long)g.RHr * g.Hp + 4LL * g.Hp * g.KPp), 256, 0, st>>>(
      c.A, c.Bm, c.Vx, at_hi, at_lo, w2_hi, w2_lo, c.H, c.RH, c.RX, g.Hp, g.RHr, g.KZP, g.KPp);
  prep_state_kernel<<<ew_grid((long long)c.B * g.Hp), 256, 0, st>>>(c.h0, hop_hi, hop_lo, c.B, c.H, g.Hp);
  split_kernel<<<ew_grid((long long)c.T * c.B * g.zxp), 256, 0, st>>>(c.zx, zx_hi, zx_lo, (long long)c.T * c.B * g.zxp);
  int rc = (int)cudaGetLastError();
  if (rc) return rc;

  CUtensorMap m_hop_hi, m_hop_lo, m_at_hi, m_at_lo, m_zop_hi, m_zop_lo, m_zx_hi, m_zx_lo, m_w2_hi, m_w2_lo;
  if (map_2d(&m_hop_hi, hop_hi, g.Hp, c.B, g.Hp) || map_2d(&m_hop_lo, hop_lo, g.Hp, c.B, g.Hp) ||
      map_2d(&m_at_hi, at_hi, g.Hp, g.RHr, g.Hp) || map_2d(&m_at_lo, at_lo, g.Hp, g.RHr, g.Hp) ||
      map_2d(&m_zop_hi, zop_hi, g.zp, c.B, g.zp) || map_2d(&m_zop_lo, zop_lo, g.zp, c.B, g.zp) ||
      map_3d(&m_zx_hi, zx_hi, g.zxp, c.B, c.T, g.zxp, (long long)c.B * g.zxp, BM, 1) ||
      map_3d(&m_zx_lo, zx_lo, g.zxp, c.B, c.T, g.zxp, (long long)c.B * g.zxp, BM, 1) ||
      map_3d(&m_w2_hi, w2_hi, g.KPp, g.Hp, 4, g.KPp, (long long)g.Hp * g.KPp, 32, 4) ||
      map_3d(&m_w2_lo, w2_lo, g.KPp, g.Hp, 4, g.KPp, (long long)g.Hp * g.KPp, 32, 4))
    return VMLMF_EUNSUPPORTED;

  FwdArgs a;
  a.x = c.x; a.xs_t = c.xs_t; a.xs_b = c.xs_b;
  a.Dx = c.Dx; a.Dh = c.Dh; a.bias = c.bias; a.h0 = c.h0; a.c0 = c.c0;
  a.y = c.y; a.ys_t = c.ys_t; a.ys_b = c.ys_b; a.hT = c.hT; a.cT = c.cT;
  a.gates = c.gates; a.cs = save ? c.cs : ws + g.o_cbuf; a.z = c.z;
  a.hop_hi = hop_hi; a.hop_lo = hop_lo; a.zop_hi = zop_hi; a.zop_lo = zop_lo; a.zpart = ws + g.o_zpart;
  a.T = c.T; a.B = c.B; a.I = c.I; a.H = c.H; a.RX = c.RX; a.RH = c.RH;
  a.Hp = g.Hp; a.HS = g.HS; a.CS = g.CS; a.zp = g.zp; a.KZP = g.KZP; a.save = save ? 1 : 0;
  a.sync = reinterpret_cast<unsigned int*>(ws + g.o_sync);
  cudaError_t me = cudaMemsetAsync(a.sync, 0, 32 * sizeof(unsigned int) * (size_t)g.ncl, st);
  if (me != cudaSuccess) return (int)me;
  const int grid = g.ncl * g.CS;
  if (save)
    return launch_coop(r2_fwd_kernel<true>, grid, kSmemBytes, st, m_hop_hi, m_hop_lo, m_at_hi, m_at_lo, m_zop_hi, m_zop_lo,
                       m_zx_hi, m_zx_lo, m_w2_hi, m_w2_lo, a);
  return launch_coop(r2_fwd_kernel<false>, grid, kSmemBytes, st, m_hop_hi, m_hop_lo, m_at_hi, m_at_lo, m_zop_hi, m_zop_lo,
                     m_zx_hi, m_zx_lo, m_w2_hi, m_w2_lo, a);
}

int launch_bwd(const BwdCall& c, void* workspace, BwdOut* out, cudaStream_t st) {
  const Geom g = geom(c.T, c.B, c.I, c.H, c.RX, c.RH);
  float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  float *dpo_lo = ws + g.b_dpo_lo, *dzo_hi = ws + g.b_dzo_hi, *dzo_lo = ws + g.b_dzo_lo, *dpre = ws + g.b_dpre;
  float *w2t_hi = ws + g.b_w2t_hi, *w2t_lo = ws + g.b_w2t_lo, *ap_hi = ws + g.b_ap_hi, *ap_lo = ws + g.b_ap_lo;
  // pad units (H <= j < Hp) are never written by the kernel but are read as tensor-core operands (against zero weights):
  // they must hold finite values.  The lo copy is small; of dPre itself only the pad columns are cleared.
  cudaError_t e = cudaMemsetAsync(dpo_lo, 0, (size_t)c.B * 4 * g.Hp * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  if (g.Hp > c.H) {
    e = cudaMemset2DAsync(dpre + c.H, (size_t)g.Hp * sizeof(float), 0, (size_t)(g.Hp - c.H) * sizeof(float),
                          (size_t)c.T * c.B * 4, st);
    if (e != cudaSuccess) return (int)e;
  }
  pack_bwd_kernel<<<ew_grid((long long)g.KPp * 4 * g.Hp + (long long)g.Hp * g.KZP), 256, 0, st>>>(
      c.A, c.Bm, c.Vx, w2t_hi, w2t_lo, ap_hi, ap_lo, c.H, c.RH, c.RX, g.Hp, g.KZP, g.KPp);
  int rc = (int)cudaGetLastError();
  if (rc) return rc;
  CUtensorMap m_dpre, m_dpo_lo, m_w2t_hi, m_w2t_lo, m_dzo_hi, m_dzo_lo, m_ap_hi, m_ap_lo;
  if (map_4d(&m_dpre, dpre, g.Hp, 4, c.B, c.T, g.Hp, 4LL * g.Hp, 4LL * g.Hp * c.B, BM) ||
      map_3d(&m_dpo_lo, dpo_lo, g.Hp, 4, c.B, g.Hp, 4LL * g.Hp, 1, BM) ||
      map_3d(&m_w2t_hi, w2t_hi, g.Hp, 4, g.KPp, g.Hp, 4LL * g.Hp, 1, 128) || map_3d(&m_w2t_lo, w2t_lo, g.Hp, 4, g.KPp, g.Hp, 4LL * g.Hp, 1, 128) ||
      map_2d(&m_dzo_hi, dzo_hi, g.zp, c.B, g.zp) || map_2d(&m_dzo_lo, dzo_lo, g.zp, c.B, g.zp) ||
      map_2d(&m_ap_hi, ap_hi, g.KZP, g.Hp, g.KZP) || map_2d(&m_ap_lo, ap_lo, g.KZP, g.Hp, g.KZP))
    return VMLMF_EUNSUPPORTED;
  BwdArgs a;
  a.gates = c.gates; a.cs = c.cs; a.c0 = c.c0; a.dy = c.dy; a.dys_t = c.dys_t; a.dys_b = c.dys_b;
  a.dhT = c.dhT; a.dcT = c.dcT; a.Dh = c.Dh; a.dh0 = c.dh0; a.dc0 = c.dc0;
  a.dpre = dpre; a.dz_all = ws + g.b_dz; a.dzx_all = ws + g.b_dzx;
  a.dpo_lo = dpo_lo; a.dzo_hi = dzo_hi; a.dzo_lo = dzo_lo;
  a.dhrun = ws + g.b_dhrun; a.dcrun = ws + g.b_dcrun; a.part = ws + g.b_part;
  a.T = c.T; a.B = c.B; a.H = c.H; a.RX = c.RX; a.RH = c.RH;
  a.Hp = g.Hp; a.HS = g.HS; a.CS = g.CS; a.zp = g.zp; a.zxp = g.zxp; a.KZP = g.KZP; a.KPp = g.KPp; a.KSPLIT = g.KSPLIT;
  out->dpre = a.dpre; out->G = g.Hp; out->dz = a.dz_all; out->dzx = a.dzx_all; out->after = ws + g.bwd_floats;
  a.sync = reinterpret_cast<unsigned int*>(ws + g.b_sync);
  e = cudaMemsetAsync(a.sync, 0, 32 * sizeof(unsigned int) * (size_t)g.ncl, st);
  if (e != cudaSuccess) return (int)e;
  return launch_coop(r2_bwd_kernel<0>, g.ncl * g.CS, kSmemBytesBwd, st, m_dpre, m_dpo_lo, m_w2t_hi, m_w2t_lo, m_dzo_hi, m_dzo_lo,
                     m_ap_hi, m_ap_lo, a);
}

}  // namespace r2
}  // namespace vmlmf

#ifdef VMLMF_R2_TRACE
// debug builds only (tools/trace_r2.py): copy the cycle trace to the host and reset it
extern "C" int vmlmf_r2_trace_read(long long* dst, int max_events) {
  int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, vmlmf::r2::g_r2_trace_n, sizeof(int));
  if (n > max_events) n = max_events;
  cudaMemcpyFromSymbol(dst, vmlmf::r2::g_r2_trace, (size_t)n * 2 * sizeof(long long));
  const int zero = 0;
  cudaMemcpyToSymbol(vmlmf::r2::g_r2_trace_n, &zero, sizeof(int));
  return n;
}
#endif
