// seq_r1_launch.cuh -- host-side launchers for the R1 kernels, instantiated per x-rank in
// seq_r1_rx*.cu so the translation units compile in parallel.
#pragma once
#include "seq_r1.cuh"

namespace vmlmf {

constexpr int kFwdBT = 4;             // sequences per CTA tile, forward
constexpr int kBwdBT = 2;             // sequences per CTA tile, backward
constexpr int kMaxCtasPerSM = 4;      // cap used to size the backward partial workspace

inline int r1_fwd_smem_bytes(int RH_T, int RX_T, int NT) {
  const int NW = NT / 32, NV = kFwdBT * next_pow2(RH_T), NZX = kFwdBT * round_up(RX_T, 4);
  return (3 * NW * NV + NW * NZX) * (int)sizeof(float);
}
inline int r1_bwd_smem_bytes(int RH_T, int RX_T, int NT) {
  const int NW = NT / 32, NV = kBwdBT * next_pow2(RH_T + RX_T);
  const int NIN = kBwdBT * (next_pow2(RH_T) + round_up(RX_T, 4));
  return (3 * NW * NV + NW * NIN) * (int)sizeof(float);
}

template <int RH_T, int RX_T>
int launch_fwd_r1(const SeqFwdArgs& a, bool save, cudaStream_t st) {
  const int NT = round_up(a.H, 32);
  const int smem = r1_fwd_smem_bytes(RH_T, RX_T, NT);
  const int ntiles = ceil_div(a.B, kFwdBT);
  auto go = [&](auto kern, int variant) -> int {
    static PerDevice occ_cache[2][9];                   // [variant][NT/32], per device
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
  if (save) return go(seq_fwd_r1_kernel<RH_T, RX_T, kFwdBT, true, 256, 1>, 1);
  return go(seq_fwd_r1_kernel<RH_T, RX_T, kFwdBT, false, 256, 1>, 0);
}

template <int RH_T, int RX_T>
int launch_bwd_r1(const SeqBwdArgs& a, const GradOut& out, cudaStream_t st) {
  const int NT = round_up(a.H, 32);
  const int smem = r1_bwd_smem_bytes(RH_T, RX_T, NT);
  const int ntiles = ceil_div(a.B, kBwdBT);
  auto kern = seq_bwd_r1_kernel<RH_T, RX_T, kBwdBT, 256, 1>;
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

// one entry per compiled x-rank; RH_T is dispatched inside
int launch_fwd_r1_rx4(int RH_T, const SeqFwdArgs& a, bool save, cudaStream_t st);
int launch_fwd_r1_rx8(int RH_T, const SeqFwdArgs& a, bool save, cudaStream_t st);
int launch_fwd_r1_rx16(int RH_T, const SeqFwdArgs& a, bool save, cudaStream_t st);
int launch_bwd_r1_rx4(int RH_T, const SeqBwdArgs& a, const GradOut& o, cudaStream_t st);
int launch_bwd_r1_rx8(int RH_T, const SeqBwdArgs& a, const GradOut& o, cudaStream_t st);
int launch_bwd_r1_rx16(int RH_T, const SeqBwdArgs& a, const GradOut& o, cudaStream_t st);

}  // namespace vmlmf
