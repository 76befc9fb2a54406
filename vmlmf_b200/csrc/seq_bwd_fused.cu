// seq_bwd_fused.cu -- instantiations and host launcher of the fused R1M backward (seq_bwd_fused.cuh).
#include "seq_bwd_fused.cuh"
#include "seq_r1_launch.cuh"

namespace vmlmf {

namespace {
int ks_of(int RX, int RH) { return ceil_div(RH + RX + 1, 8); }
int nz_of(int RH) { return ceil_div(RH, 8); }
int tmem_cols_of(int NW) {
  int c = kFusedCols * ceil_div(NW, 4), p = 32;
  while (p < c) p <<= 1;
  return p;
}
// CTAs that can share an SM: shared memory (227 KB, 1 KB reserved per CTA), tensor memory (512 columns), the register
// file (128 registers per thread by __launch_bounds__) and 2048 threads.  Analytic so that vmlmf_seq_plan (no device
// calls) and the launch agree on the number of per-CTA gradient partials.
int occ_of(int NW, int KS, int I, int RX) {
  const size_t smem = seq_bwd_fused_smem_bytes(NW, KS, I, RX) + 1024;
  int occ = (int)((size_t)227 * 1024 / smem);
  occ = occ < 512 / tmem_cols_of(NW) ? occ : 512 / tmem_cols_of(NW);
  occ = occ < 65536 / (NW * 32 * 128) ? occ : 65536 / (NW * 32 * 128);
  return occ < 1 ? 1 : occ;
}
int grid_of(int B, int NW, int KS, int I, int RX) {
  const int nt = ceil_div(B, 16), cap = num_sms() * occ_of(NW, KS, I, RX);
  return nt < cap ? nt : cap;
}
}  // namespace

bool bwd_fused_fits(int I, int H, int RX, int RH) {
  if (H > 256 || (H & 3) || I > H || RH > 16) return false;
  const int KS = ks_of(RX, RH), NZ = nz_of(RH);
  if (KS > 2 || NZ > KS) return false;
  return seq_bwd_fused_smem_bytes(ceil_div(H, 16), KS, I, RX) <= 227 * 1024;
}

// blocks of dux_rows_kernel: at most three per SM (three pipeline stages of shared memory each); two partials per block
int dux_grid(int T, int B) {
  const long long tiles = ((long long)T * B + kDuxRows - 1) / kDuxRows;
  return (int)(tiles < 3LL * num_sms() ? tiles : 3LL * num_sms());
}

// workspace: [per-CTA gradient partials | dzx rows | per-block dUx partials], each part a multiple of 4 floats
long long bwd_fused_workspace_floats(int T, int B, int I, int H, int RX, int RH) {
  if (!bwd_fused_fits(I, H, RX, RH)) return 0;
  const GradLayout L(I, H, RX, RH);
  const long long zxp = round_up(RX, 4);
  const long long part = round_up((int)((long long)grid_of(B, ceil_div(H, 16), ks_of(RX, RH), I, RX) * L.total + 8), 4);
  return part + (long long)T * B * zxp + 2LL * dux_grid(T, B) * I * zxp;
}

template <int KS, int NZ>
static int launch_t(const SeqBwdFusedArgs& a, int NW, int G, cudaStream_t st) {
  auto kern = seq_bwd_fused_kernel<KS, NZ>;
  const size_t smem = seq_bwd_fused_smem_bytes(NW, KS, a.I, a.RX);
  static PerDevice attr;                                // the attribute is per device
  int& attr_done = attr.cur();
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_done = 1;
  }
  kern<<<G, NW * 32, smem, st>>>(a, tmem_cols_of(NW));
  return (int)cudaGetLastError();
}

int launch_bwd_fused(const SeqBwdFusedArgs& a0, const GradOut& out, void* workspace, cudaStream_t st) {
  const int KS = ks_of(a0.RX, a0.RH), NZ = nz_of(a0.RH), NW = ceil_div(a0.H, 16);
  const int G = grid_of(a0.B, NW, KS, a0.I, a0.RX);
  const GradLayout L(a0.I, a0.H, a0.RX, a0.RH);
  SeqBwdFusedArgs a = a0;
  const int zxp = a0.zxp;
  const long long part = round_up((int)((long long)G * L.total + 8), 4);
  a.partial = (float*)workspace;
  a.dzc = a.partial + part;                             // dzx rows, in the row order of x when x is one contiguous block
  float* pbuf = a.dzc + (size_t)a0.T * a0.B * zxp;
  const bool x_bt = a0.xs_t == a0.I && a0.xs_b == (long long)a0.T * a0.I;
  const bool x_tb = a0.xs_b == a0.I && a0.xs_t == (long long)a0.B * a0.I;
  a.dz_bt = x_bt ? 1 : 0;
  int rc = kMmaNoFit;
  if (KS == 1) rc = launch_t<1, 1>(a, NW, G, st);
  else if (KS == 2 && NZ == 1) rc = launch_t<2, 1>(a, NW, G, st);
  else if (KS == 2 && NZ == 2) rc = launch_t<2, 2>(a, NW, G, st);
  if (rc) return rc;
  reduce_partials_kernel<<<ceil_div(L.total, kReduceElems), 256, 0, st>>>(a.partial, G, L, out);
  // dUx = X^T dZX (overwrites the unwritten dUx slice the reduce just summed)
  const int GX = dux_grid(a0.T, a0.B);
  const size_t stage_bytes = (size_t)kDuxRows * (a0.I + zxp) * sizeof(float);
  const bool pipelined = (x_bt || x_tb) && (reinterpret_cast<uintptr_t>(a0.x) & 15) == 0 &&
                         3 * (kDuxStages * stage_bytes + 1024) <= 227 * 1024;
  const size_t dsm = (pipelined ? kDuxStages : 1) * stage_bytes;
  DuxArgs d{a0.x, a0.xs_t, a0.xs_b, a.dzc, pbuf, a0.T, a0.B, a0.I, zxp, (x_bt || x_tb) ? 1 : 0, pipelined ? kDuxStages : 1};
  static PerDevice dux_attr_pd;
  int& dux_attr = dux_attr_pd.cur();
  if (!dux_attr) {
    cudaError_t e = cudaFuncSetAttribute(dux_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    dux_attr = 1;
  }
  dux_rows_kernel<<<GX, kDuxThreads, dsm, st>>>(d);
  dux_reduce_kernel<<<ceil_div(a0.I * a0.RX, 4), 256, 0, st>>>(pbuf, 2 * GX, a0.I, a0.RX, zxp, out.dUx);
  return (int)cudaGetLastError();
}

}  // namespace vmlmf
