// seq_r1_launch.cuh -- host-side launchers for the R1 kernels, instantiated per x-rank in
// seq_r1_rx*.cu so the translation units compile in parallel.
#pragma once
#include "seq_r1.cuh"

namespace vmlmf {

constexpr int kFwdBT = 4;             // sequences per CTA tile, forward
constexpr int kBwdBT = 2;             // sequences per CTA tile, backward
constexpr int kMaxCtasPerSM = 4;      // cap used to size the backward partial workspace
// Small batches (the reference's own 64 / 81 sequences: cfg1, the CLI default) are latency-bound: a step costs what ONE
// CTA's instruction stream costs, and with 4 sequences per CTA only B/4 of the 148 SMs work at all.  One sequence per CTA
// cuts the per-step instruction count ~3x and spreads the batch over B SMs (cfg1: 0.97 -> see DESIGN.md us per step).
// (template ranks below 4 keep the wide tiles: the kernels' shared rows are float4-aligned only when BT * pow2(RH_T) % 4 == 0)
inline int r1_fwd_bt(int B, int rh_t) { return (rh_t >= 4 && ceil_div(B, kFwdBT) * 2 <= num_sms()) ? 1 : kFwdBT; }
inline int r1_bwd_bt(int B, int rh_t) { return (rh_t >= 4 && ceil_div(B, kBwdBT) * 2 <= num_sms()) ? 1 : kBwdBT; }

inline int r1_fwd_smem_bytes(int RH_T, int RX_T, int NT, int BT = kFwdBT) {
  const int NW = NT / 32, NV = BT * next_pow2(RH_T), NZX = BT * round_up(RX_T, 4);
  return (3 * NW * NV + NW * NZX) * (int)sizeof(float);
}
inline int r1_bwd_smem_bytes(int RH_T, int RX_T, int NT, int BT = kBwdBT) {
  const int NW = NT / 32, NV = BT * next_pow2(RH_T + RX_T);
  const int NIN = BT * (next_pow2(RH_T) + round_up(RX_T, 4));
  return (3 * NW * NV + NW * NIN) * (int)sizeof(float);
}

template <int RH_T, int RX_T, int BT>
int launch_fwd_r1_bt(const SeqFwdArgs& a, bool save, cudaStream_t st) {
  const int NT = round_up(a.H, 32);
  const int smem = r1_fwd_smem_bytes(RH_T, RX_T, NT, BT);
  const int ntiles = ceil_div(a.B, BT);
  auto go = [&](auto kern, int variant) -> int {
    static PerDevice occ_cache[2][9];                   // [variant][NT/32], per device (and per BT: template)
    int& slot = occ_cache[variant][NT / 32].cur();
    int occ = slot;
    if (occ == 0) {
      cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
      if (e != cudaSuccess) return (int)e;
      if (occ < 1) occ = 1;
      slot = occ;
    }
    const int grid = ntiles < num_sms() * occ ? ntiles : num_sms() * occ;
    kern<<<grid, NT, smem, st>>>(a);
    return (int)cudaGetLastError();
  };
  if (save) return go(seq_fwd_r1_kernel<RH_T, RX_T, BT, true, 256, 1>, 1);
  return go(seq_fwd_r1_kernel<RH_T, RX_T, BT, false, 256, 1>, 0);
}
template <int RH_T, int RX_T>
int launch_fwd_r1(const SeqFwdArgs& a, bool save, cudaStream_t st) {
  return r1_fwd_bt(a.B, RH_T) == 1 ? launch_fwd_r1_bt<RH_T, RX_T, 1>(a, save, st) : launch_fwd_r1_bt<RH_T, RX_T, kFwdBT>(a, save, st);
}

template <int RH_T, int RX_T, int BT>
int launch_bwd_r1_bt(const SeqBwdArgs& a, const GradOut& out, cudaStream_t st) {
  const int NT = round_up(a.H, 32);
  const int smem = r1_bwd_smem_bytes(RH_T, RX_T, NT, BT);
  const int ntiles = ceil_div(a.B, BT);
  auto kern = seq_bwd_r1_kernel<RH_T, RX_T, BT, 256, 1>;
  static PerDevice occ_cache[9];
  int& slot = occ_cache[NT / 32].cur();
  int occ = slot;
  cudaError_t e;
  if (occ == 0) {
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) occ = 1;
    if (occ > kMaxCtasPerSM) occ = kMaxCtasPerSM;
    slot = occ;
  }
  const int grid = ntiles < num_sms() * occ ? ntiles : num_sms() * occ;
  kern<<<grid, NT, smem, st>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  const GradLayout L(a.I, a.H, a.RX, a.RH);
  reduce_partials_kernel<<<ceil_div(L.total, kReduceElems), 256, 0, st>>>(a.partial, grid, L, out);
  return (int)cudaGetLastError();
}
template <int RH_T, int RX_T>
int launch_bwd_r1(const SeqBwdArgs& a, const GradOut& out, cudaStream_t st) {
  return r1_bwd_bt(a.B, RH_T) == 1 ? launch_bwd_r1_bt<RH_T, RX_T, 1>(a, out, st) : launch_bwd_r1_bt<RH_T, RX_T, kBwdBT>(a, out, st);
}

// one entry per compiled x-rank; RH_T is dispatched inside
int launch_fwd_r1_rx4(int RH_T, const SeqFwdArgs& a, bool save, cudaStream_t st);
int launch_fwd_r1_rx8(int RH_T, const SeqFwdArgs& a, bool save, cudaStream_t st);
int launch_fwd_r1_rx16(int RH_T, const SeqFwdArgs& a, bool save, cudaStream_t st);
int launch_bwd_r1_rx4(int RH_T, const SeqBwdArgs& a, const GradOut& o, cudaStream_t st);
int launch_bwd_r1_rx8(int RH_T, const SeqBwdArgs& a, const GradOut& o, cudaStream_t st);
int launch_bwd_r1_rx16(int RH_T, const SeqBwdArgs& a, const GradOut& o, cudaStream_t st);

}  // namespace vmlmf
