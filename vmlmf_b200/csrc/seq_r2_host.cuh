// seq_r2_host.cuh -- host interface of regime R2 (implemented in seq_r2.cu, device code in seq_r2.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vmlmf {
namespace r2 {

// geometry of one call: cluster size, hidden slice per CTA, padded sizes, workspace offsets (in floats)
struct Geom {
  int ntiles, CS, HS, Hp, zp, zxp, RHr, KZP, KXP, KPp, ncl;
  long long o_hop_hi, o_hop_lo, o_zop_hi, o_zop_lo, o_zpart, o_zx_hi, o_zx_lo, o_at_hi, o_at_lo, o_w2_hi, o_w2_lo, o_cbuf, o_sync;
  long long fwd_floats;
  // backward
  int KSPLIT, NP;            // K segments of phase 1 per CTA (<= 64 tf32 k-steps per accumulator), partials per tile
  long long b_dpre, b_dz, b_dzx, b_dpo_hi, b_dpo_lo, b_dzo_hi, b_dzo_lo, b_dhrun, b_dcrun, b_part, b_w2t_hi, b_w2t_lo,
      b_ap_hi, b_ap_lo, b_sync;
  long long bwd_floats;      // recurrence part only; the time-parallel GEMMs' scratch follows it in the caller's workspace
};
Geom geom(int T, int B, int I, int H, int RX, int RH);
bool fits(int T, int B, int I, int H, int RX, int RH);

struct FwdCall {
  const float* x; long long xs_t, xs_b;
  const float* zx;                                   // [T*B, zxp]
  const float *Vx, *Dx, *A, *Bm, *Dh, *bias, *h0, *c0;
  float* y; long long ys_t, ys_b;
  float *hT, *cT, *gates, *cs, *z;
  int T, B, I, H, RX, RH;
};
int launch_fwd(const FwdCall& c, void* workspace, cudaStream_t st);

struct BwdCall {
  const float *Vx, *A, *Bm, *Dh, *c0;
  const float *gates, *cs;
  const float* dy; long long dys_t, dys_b;
  const float *dhT, *dcT;
  float *dh0, *dc0;
  int T, B, I, H, RX, RH;
};
// what the recurrence kernel leaves behind for the time-parallel half of the backward
struct BwdOut { float* dpre; int G; float* dz; float* dzx; float* after; };
int launch_bwd(const BwdCall& c, void* workspace, BwdOut* out, cudaStream_t st);

}  // namespace r2
}  // namespace vmlmf

// regime R3 (seq_r3.cuh / seq_r3.cu): small batches (B <= 32), factors stationary in shared memory
namespace vmlmf {
namespace r3 {

struct Geom {
  int CS, Hp, zp, zxp, RHr, KZP, nkz, nch, S_fwd, G_fwd, S_bwd, G_bwd, smem_fwd, smem_bwd;
  long long o_xp, o_hop, o_zop, o_zpart, o_p_hi, o_p_lo, o_w2_hi, o_w2_lo, o_cbuf, o_sync;
  long long fwd_floats;
  long long b_dpre, b_dz, b_dzx, b_dpo, b_dzo, b_dhrun, b_dcrun, b_part, b_w2t_hi, b_w2t_lo, b_ap_hi, b_ap_lo, b_vxt, b_sync;
  long long bwd_floats;      // recurrence part; the time-parallel GEMMs' scratch follows it
};
Geom geom(int T, int B, int I, int H, int RX, int RH);
bool fits(int T, int B, int I, int H, int RX, int RH);

struct FwdCall {
  const float* xp;                                   // [T*B, 4H] = zx Vx^T + bias + x (.) Dx (formed by the caller)
  const float *A, *Bm, *Dh, *h0, *c0;
  float* y; long long ys_t, ys_b;
  float *hT, *cT, *gates, *cs, *z;
  int T, B, I, H, RX, RH;
};
// xp lives at Geom::o_xp of the (256-byte aligned) workspace: ws_base(workspace) + o_xp
inline float* ws_base(void* workspace) { return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255); }
int launch_fwd(const FwdCall& c, void* workspace, cudaStream_t st);

struct BwdCall {
  const float *Vx, *A, *Bm, *Dh, *c0;
  const float *gates, *cs;
  const float* dy; long long dys_t, dys_b;
  const float *dhT, *dcT;
  float *dh0, *dc0;
  int T, B, I, H, RX, RH;
};
// dzx is only a buffer here: the caller forms dzx = dPre Vx (time-parallel) with vxt = Vx^T in dPre's gate-padded layout
struct BwdOut { float* dpre; int G; float* dz; float* dzx; float* vxt; float* after; };
int launch_bwd(const BwdCall& c, void* workspace, BwdOut* out, cudaStream_t st);

}  // namespace r3
}  // namespace vmlmf
