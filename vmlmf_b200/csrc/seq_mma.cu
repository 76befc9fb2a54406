// seq_mma.cu -- instantiations and host launcher of the warp-MMA R1 forward kernel (seq_mma.cuh).
#include "seq_mma.cuh"
#include "seq_r1_launch.cuh"

namespace vmlmf {

template <int KS, int NZ>
static int launch_fwd_mma_t(const SeqFwdMmaArgs& a, bool save, cudaStream_t st) {
  const int NW = ceil_div(a.s.H, 16);
  const size_t smem = seq_fwd_mma_smem_bytes(NW, KS, NZ);
  if (smem > 227 * 1024) return kMmaNoFit;
  auto go = [&](auto kern, int variant) -> int {
    // attribute + occupancy are per (instantiation, NW): cached so a launch costs one driver call
    // (benign race: every thread computes the same values)
    static PerDevice occ_cache[2][17];               // [variant][NW], per device: the two variants share one pointer type
    int& slot = occ_cache[variant][NW].cur();
    int occ = slot;
    if (occ == 0) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return (int)e;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NW * 32, smem);
      if (e != cudaSuccess) return (int)e;
      if (occ < 1) occ = 1;
      slot = occ;
    }
    const int ntiles = ceil_div(a.s.B, 16);
    const int grid = ntiles < num_sms() * occ ? ntiles : num_sms() * occ;
    kern<<<grid, NW * 32, smem, st>>>(a);
    return (int)cudaGetLastError();
  };
  if (save) return go(seq_fwd_mma_kernel<KS, NZ, true>, 1);
  return go(seq_fwd_mma_kernel<KS, NZ, false>, 0);
}

bool fwd_mma_fits(int I, int H, int RX, int RH) {
  if (H > 256 || (H & 3) || I > H || RH > 16) return false;
  const int KS = ceil_div(RH + RX + 1, 8), NZ = ceil_div(RH, 8);
  if (KS > 4 || (NZ == 1 && KS > 3) || (NZ == 2 && KS < 2)) return false;
  return seq_fwd_mma_smem_bytes(ceil_div(H, 16), KS, NZ) <= 227 * 1024;
}

int launch_fwd_mma(const SeqFwdMmaArgs& a, bool save, cudaStream_t st) {
  const int RH = a.s.RH, RX = a.s.RX;
  if (a.s.H > 256 || RH > 16) return kMmaNoFit;
  const int KS = ceil_div(RH + RX + 1, 8), NZ = ceil_div(RH, 8);
  if (NZ == 1) {
    if (KS == 1) return launch_fwd_mma_t<1, 1>(a, save, st);
    if (KS == 2) return launch_fwd_mma_t<2, 1>(a, save, st);
    if (KS == 3) return launch_fwd_mma_t<3, 1>(a, save, st);
  } else if (NZ == 2) {
    if (KS == 2) return launch_fwd_mma_t<2, 2>(a, save, st);
    if (KS == 3) return launch_fwd_mma_t<3, 2>(a, save, st);
    if (KS == 4) return launch_fwd_mma_t<4, 2>(a, save, st);
  }
  return kMmaNoFit;
}

}  // namespace vmlmf
