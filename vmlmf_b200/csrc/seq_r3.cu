// seq_r3.cu -- host side of regime R3 (seq_r3.cuh): geometry, slice-major operand packing, tensor maps, launches.
#include "seq_r3.cuh"

#include <stdlib.h>

#include "../../include/vmlmf_b200.h"

#include "seq_r2_host.cuh"

namespace vmlmf {
namespace r3 {

namespace {

inline long long al64(long long n) { return (n + 63) / 64 * 64; }

// ---- operand packing (once per call) ----
// p[s*128 + n, q*8 + u]   = A[8s+u, q*128+n]     phase Z B operand of CTA s: chunk q of the z columns in k-step slot q
// w2[s*32 + k*8 + u, q]   = Bm[kH + 8s+u, q]      phase G B operand of CTA s: N = (gate k, unit u), K = z column q
// zero outside the matrices; every value split into tf32 hi and lo parts.
__global__ void pack3_fwd_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ p_hi,
                                 float* __restrict__ p_lo, float* __restrict__ w2_hi, float* __restrict__ w2_lo, int H, int RH,
                                 int CS, int KZP) {
  const long long n_p = (long long)CS * 128 * 32, n_w = (long long)CS * 32 * KZP;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_p + n_w; i += (long long)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < n_p) {
      const int col = (int)(i & 31), n = (int)((i >> 5) & 127), s = (int)(i >> 12);
      const int q = col >> 3, u = col & 7;
      const int j = 8 * s + u, r = q * 128 + n;
      if (j < H && r < RH) v = __ldg(A + (size_t)j * RH + r);
      p_hi[i] = split_hi(v);
      p_lo[i] = split_lo(v, v);
    } else {
      const long long e = i - n_p;
      const int q = (int)(e % KZP);
      const int row = (int)(e / KZP);
      const int s = row >> 5, k = (row >> 3) & 3, u = row & 7;
      const int j = 8 * s + u;
      if (j < H && q < RH) v = __ldg(Bm + ((size_t)k * H + j) * RH + q);
      w2_hi[e] = split_hi(v);
      w2_lo[e] = split_lo(v, v);
    }
  }
}
// w2t[n, s*32 + k*8 + u]  = Bm[kH + 8s+u, n]      phase 1 B operand: N = dz column n, K = CTA s's (gate, unit) slice
// ap[j, q]                = A[j, q]               phase 2 B operand: N = unit j, K = z column q      [Hp, KZP]
__global__ void pack3_bwd_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ w2t_hi,
                                 float* __restrict__ w2t_lo, float* __restrict__ ap_hi, float* __restrict__ ap_lo, int H, int RH,
                                 int Hp, int KZP) {
  const long long n_w = (long long)KZP * 4 * Hp, n_a = (long long)Hp * KZP;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_w + n_a; i += (long long)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < n_w) {
      const int col = (int)(i % (4 * Hp)), n = (int)(i / (4 * Hp));
      const int s = col >> 5, k = (col >> 3) & 3, u = col & 7;
      const int j = 8 * s + u;
      if (j < H && n < RH) v = __ldg(Bm + ((size_t)k * H + j) * RH + n);
      w2t_hi[i] = split_hi(v);
      w2t_lo[i] = split_lo(v, v);
    } else {
      const long long e = i - n_w;
      const int q = (int)(e % KZP), j = (int)(e / KZP);
      if (j < H && q < RH) v = __ldg(A + (size_t)j * RH + q);
      ap_hi[e] = split_hi(v);
      ap_lo[e] = split_lo(v, v);
    }
  }
}
// hop tile of CTA j / 8 (tile-major, zeroed before): [hi | lo][row b][column j % 8] = h0[b, j]
__global__ void prep3_state_kernel(const float* __restrict__ h0, float* __restrict__ hop, int B, int H) {
  const long long n = (long long)B * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / H), j = (int)(i % H);
    const float v = h0[i];
    float* t = hop + (size_t)(j >> 3) * kTileFloats + b * 32 + (j & 7);
    t[0] = split_hi(v);
    t[1024] = split_lo(v, v);
  }
}
// vxt[r, k*G + j] = Vx[kH + j, r]  (zero for j >= H): the B operand of dzx = dPre Vx with dPre in its gate-padded layout
__global__ void vxt_pad_kernel(const float* __restrict__ Vx, float* __restrict__ vxt, int H, int G, int RX) {
  const long long n = (long long)RX * 4 * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % (4 * G)), r = (int)(i / (4 * G));
    const int k = col / G, j = col - k * G;
    vxt[i] = j < H ? __ldg(Vx + ((size_t)k * H + j) * RX + r) : 0.f;
  }
}

inline int ew_grid(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = 8LL * num_sms();
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

// [rows][cols] row-major view with row pitch ld (floats); box = 32 columns x box_rows
int map2(CUtensorMap* m, const float* p, long long cols, long long rows, long long ld, int box_rows) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  if (!fn) return tc::kTcNoFit;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {BK, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : tc::kTcNoFit;
}

template <class Kern, class... Args>
int launch_coop(Kern kern, int grid, int smem_bytes, cudaStream_t st, Args... args) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return (int)e;
  void* argv[] = {(void*)&args...};
  return (int)cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kThreads), argv, smem_bytes, st);
}

bool r3_disabled() {
  const char* e = getenv("VMLMF_NO_R3");
  return e && e[0] == '1';
}

}  // namespace

Geom geom(int T, int B, int I, int H, int RX, int RH) {
  (void)I;
  Geom g;
  g.CS = ceil_div(H, UB);
  g.Hp = g.CS * UB;
  g.zp = round_up(RH, 4);
  g.zxp = round_up(RX, 4);
  g.RHr = round_up(RH, 8);
  g.KZP = round_up(RH, 32);
  g.nkz = ceil_div(RH, BK);
  g.nch = ceil_div(g.RHr, 128);
  // shared memory: stationary factor rows + an activation ring of S stages of G K tiles (8 KB each); one TMA box and one
  // barrier per stage, so the bigger the stage the fewer barrier round trips of the issuing thread
  const int stat_f = 2 * kPTile + g.nkz * kAStage;
  const int stat_b = g.nch * 2 * kPTile + g.nkz * 2 * kApTile;
  auto ring = [&](int stat, int* S, int* G) {
    const int room = kSmemMax - kBarBytes - kXbufBytes - stat;
    for (int gg = 4; gg >= 1; gg >>= 1) {
      int ss = room / (gg * kAStage);
      if (ss > kMaxStages) ss = kMaxStages;
      const int need = ceil_div(g.nkz, gg) + 1;            // every stage of a step resident at once is as deep as it gets
      if (ss > need) ss = need;
      if (ss >= 2 || gg == 1) { *S = ss; *G = gg; return; }
    }
  };
  ring(stat_f, &g.S_fwd, &g.G_fwd);
  ring(stat_b, &g.S_bwd, &g.G_bwd);
  g.smem_fwd = g.S_fwd * g.G_fwd * kAStage + stat_f + kBarBytes + kXbufBytes + 1024;
  g.smem_bwd = g.S_bwd * g.G_bwd * kAStage + stat_b + kBarBytes + kXbufBytes + 1024;
  long long o = 0;
  g.o_xp = o; o += al64((long long)T * B * 4 * H);
  g.o_hop = o; o += al64((long long)g.CS * kTileFloats);
  g.o_zop = o; o += al64((long long)g.nkz * kTileFloats);
  g.o_zpart = o; o += al64((long long)g.CS * RB * g.zp);
  g.o_p_hi = o; o += al64((long long)g.CS * 128 * 32);
  g.o_p_lo = o; o += al64((long long)g.CS * 128 * 32);
  g.o_w2_hi = o; o += al64((long long)g.CS * 32 * g.KZP);
  g.o_w2_lo = o; o += al64((long long)g.CS * 32 * g.KZP);
  g.o_cbuf = o; o += al64(2LL * B * H);
  g.o_sync = o; o += 64;
  g.fwd_floats = o + 64;
  o = 0;
  g.b_dpre = o; o += al64((long long)T * B * 4 * g.Hp);
  g.b_dz = o; o += al64((long long)T * B * g.zp);
  g.b_dzx = o; o += al64((long long)T * B * g.zxp);
  g.b_dpo = o; o += al64((long long)g.CS * kTileFloats);
  g.b_dzo = o; o += al64((long long)g.nkz * kTileFloats);
  g.b_dhrun = o; o += al64((long long)B * g.Hp);
  g.b_dcrun = o; o += al64((long long)B * g.Hp);
  g.b_part = o; o += al64((long long)g.CS * RB * g.zp);
  g.b_w2t_hi = o; o += al64((long long)g.KZP * 4 * g.Hp);
  g.b_w2t_lo = o; o += al64((long long)g.KZP * 4 * g.Hp);
  g.b_ap_hi = o; o += al64((long long)g.Hp * g.KZP);
  g.b_ap_lo = o; o += al64((long long)g.Hp * g.KZP);
  g.b_vxt = o; o += al64((long long)RX * 4 * g.Hp);
  g.b_sync = o; o += 64;
  g.bwd_floats = o + 64;
  return g;
}

bool fits(int T, int B, int I, int H, int RX, int RH) {
  if (r3_disabled()) return false;
  if (B > RB || I > H || (long long)T * B > 0x7fffffffLL) return false;
  if (tc::encode_fn() == nullptr) return false;
  const int cs = ceil_div(H, UB);
  if (cs < 2 || cs > num_sms()) return false;               // one CTA per SM, all co-resident
  if (round_up(RH, 8) > 512) return false;                  // the phase Z tile holds four 128-column chunks
  const Geom g = geom(T, B, I, H, RX, RH);
  return g.S_fwd >= 2 && g.S_bwd >= 2;
}

int launch_fwd(const FwdCall& c, void* workspace, cudaStream_t st) {
  const Geom g = geom(c.T, c.B, c.I, c.H, c.RX, c.RH);
  float* ws = ws_base(workspace);
  float *hop = ws + g.o_hop, *zop = ws + g.o_zop;
  float *p_hi = ws + g.o_p_hi, *p_lo = ws + g.o_p_lo, *w2_hi = ws + g.o_w2_hi, *w2_lo = ws + g.o_w2_lo;
  const bool save = c.gates != nullptr;
  // rows >= B, pad columns and pad units of the operand tiles are never written by the kernel: zero them once (hop and zop are adjacent)
  cudaError_t me = cudaMemsetAsync(hop, 0, (size_t)(g.o_zpart - g.o_hop) * sizeof(float), st);
  if (me != cudaSuccess) return (int)me;
  pack3_fwd_kernel<<<ew_grid((long long)g.CS * 128 * 32 + (long long)g.CS * 32 * g.KZP), 256, 0, st>>>(
      c.A, c.Bm, p_hi, p_lo, w2_hi, w2_lo, c.H, c.RH, g.CS, g.KZP);
  if (c.h0) prep3_state_kernel<<<ew_grid((long long)c.B * c.H), 256, 0, st>>>(c.h0, hop, c.B, c.H);
  int rc = (int)cudaGetLastError();
  if (rc) return rc;
  CUtensorMap m_hop, m_zop, m_p_hi, m_p_lo, m_w2_hi, m_w2_lo;
  if (map2(&m_hop, hop, 32, (long long)g.CS * 64, 32, 64) || map2(&m_zop, zop, 32, (long long)g.nkz * 64, 32, 64 * g.G_fwd) ||
      map2(&m_p_hi, p_hi, 32, (long long)g.CS * 128, 32, BM) || map2(&m_p_lo, p_lo, 32, (long long)g.CS * 128, 32, BM) ||
      map2(&m_w2_hi, w2_hi, g.KZP, (long long)g.CS * 32, g.KZP, 32) || map2(&m_w2_lo, w2_lo, g.KZP, (long long)g.CS * 32, g.KZP, 32))
    return VMLMF_EUNSUPPORTED;
  FwdArgs3 a;
  a.xp = c.xp; a.Dh = c.Dh; a.h0 = c.h0; a.c0 = c.c0;
  a.y = c.y; a.ys_t = c.ys_t; a.ys_b = c.ys_b; a.hT = c.hT; a.cT = c.cT;
  a.gates = c.gates; a.cs = save ? c.cs : ws + g.o_cbuf; a.z = c.z;
  a.hop = hop; a.zop = zop; a.zpart = ws + g.o_zpart;
  a.sync = reinterpret_cast<unsigned int*>(ws + g.o_sync);
  a.T = c.T; a.B = c.B; a.H = c.H; a.RH = c.RH;
  a.CS = g.CS; a.zp = g.zp; a.S = g.S_fwd; a.G = g.G_fwd;
  me = cudaMemsetAsync(a.sync, 0, 32 * sizeof(unsigned int), st);
  if (me != cudaSuccess) return (int)me;
  if (save)
    return launch_coop(r3_fwd_kernel<true>, g.CS, g.smem_fwd, st, m_hop, m_zop, m_p_hi, m_p_lo, m_w2_hi, m_w2_lo, a);
  return launch_coop(r3_fwd_kernel<false>, g.CS, g.smem_fwd, st, m_hop, m_zop, m_p_hi, m_p_lo, m_w2_hi, m_w2_lo, a);
}

int launch_bwd(const BwdCall& c, void* workspace, BwdOut* out, cudaStream_t st) {
  const Geom g = geom(c.T, c.B, c.I, c.H, c.RX, c.RH);
  float* ws = ws_base(workspace);
  float *dpre = ws + g.b_dpre, *dpo = ws + g.b_dpo, *dzo = ws + g.b_dzo;
  float *w2t_hi = ws + g.b_w2t_hi, *w2t_lo = ws + g.b_w2t_lo, *ap_hi = ws + g.b_ap_hi, *ap_lo = ws + g.b_ap_lo;
  // rows >= B / pad columns / pad units of the operand tiles and the pad units of dPre (read by the time-parallel GEMMs) are
  // never written by the kernel: they must hold zeros (dpo and dzo are adjacent)
  cudaError_t e = cudaMemsetAsync(dpo, 0, (size_t)(g.b_dhrun - g.b_dpo) * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  if (g.Hp > c.H) {
    e = cudaMemset2DAsync(dpre + c.H, (size_t)g.Hp * sizeof(float), 0, (size_t)(g.Hp - c.H) * sizeof(float),
                          (size_t)c.T * c.B * 4, st);
    if (e != cudaSuccess) return (int)e;
  }
  pack3_bwd_kernel<<<ew_grid((long long)g.KZP * 4 * g.Hp + (long long)g.Hp * g.KZP), 256, 0, st>>>(
      c.A, c.Bm, w2t_hi, w2t_lo, ap_hi, ap_lo, c.H, c.RH, g.Hp, g.KZP);
  int rc = (int)cudaGetLastError();
  if (rc) return rc;
  CUtensorMap m_dpo, m_dzo, m_w2t_hi, m_w2t_lo, m_ap_hi, m_ap_lo;
  if (map2(&m_dpo, dpo, 32, (long long)g.CS * 64, 32, 64) || map2(&m_dzo, dzo, 32, (long long)g.nkz * 64, 32, 64 * g.G_bwd) ||
      map2(&m_w2t_hi, w2t_hi, 4LL * g.Hp, g.KZP, 4LL * g.Hp, BM) || map2(&m_w2t_lo, w2t_lo, 4LL * g.Hp, g.KZP, 4LL * g.Hp, BM) ||
      map2(&m_ap_hi, ap_hi, g.KZP, g.Hp, g.KZP, 16) || map2(&m_ap_lo, ap_lo, g.KZP, g.Hp, g.KZP, 16))
    return VMLMF_EUNSUPPORTED;
  BwdArgs3 a;
  a.gates = c.gates; a.cs = c.cs; a.c0 = c.c0; a.dy = c.dy; a.dys_t = c.dys_t; a.dys_b = c.dys_b;
  a.dhT = c.dhT; a.dcT = c.dcT; a.Dh = c.Dh; a.dh0 = c.dh0; a.dc0 = c.dc0;
  a.dpre = dpre; a.dz_all = ws + g.b_dz; a.dpo = dpo; a.dzo = dzo;
  a.dhrun = ws + g.b_dhrun; a.dcrun = ws + g.b_dcrun; a.part = ws + g.b_part;
  a.sync = reinterpret_cast<unsigned int*>(ws + g.b_sync);
  a.T = c.T; a.B = c.B; a.H = c.H; a.RH = c.RH;
  a.Hp = g.Hp; a.CS = g.CS; a.zp = g.zp; a.S = g.S_bwd; a.G = g.G_bwd;
  float* vxt = ws + g.b_vxt;
  vxt_pad_kernel<<<ew_grid((long long)c.RX * 4 * g.Hp), 256, 0, st>>>(c.Vx, vxt, c.H, g.Hp, c.RX);
  out->dpre = dpre; out->G = g.Hp; out->dz = a.dz_all; out->dzx = ws + g.b_dzx; out->vxt = vxt; out->after = ws + g.bwd_floats;
  e = cudaMemsetAsync(a.sync, 0, 32 * sizeof(unsigned int), st);
  if (e != cudaSuccess) return (int)e;
  return launch_coop(r3_bwd_kernel, g.CS, g.smem_bwd, st, m_dpo, m_dzo, m_w2t_hi, m_w2t_lo, m_ap_hi, m_ap_lo, a);
}

}  // namespace r3
}  // namespace vmlmf

#ifdef VMLMF_R2_TRACE
// debug builds only (tools/trace_r2.py): this translation unit's copy of the cycle trace (10 warps x 96 slots, zero = unused)
extern "C" int vmlmf_r3_trace_read(long long* dst, int max_events) {
  static long long host[2 * 960];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host, vmlmf::r2::g_r2_trace, sizeof(host));
  int n = 0;
  for (int i = 0; i < 960 && n < max_events; ++i)
    if (host[2 * i + 1] != 0) { dst[2 * n] = host[2 * i]; dst[2 * n + 1] = host[2 * i + 1]; ++n; }
  static long long zeros[2 * 960];
  cudaMemcpyToSymbol(vmlmf::r2::g_r2_trace, zeros, sizeof(zeros));
  return n;
}
#endif
