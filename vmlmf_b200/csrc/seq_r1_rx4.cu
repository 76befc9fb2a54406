// seq_r1_rx4.cu -- R1 kernel instantiations for x-side template rank 4.
#include "seq_r1_launch.cuh"
namespace vmlmf {
int launch_fwd_r1_rx4(int RH_T, const SeqFwdArgs& a, bool save, cudaStream_t st) {
  switch (RH_T) {
    case 2: return launch_fwd_r1<2, 4>(a, save, st);
    case 4: return launch_fwd_r1<4, 4>(a, save, st);
    case 6: return launch_fwd_r1<6, 4>(a, save, st);
    case 8: return launch_fwd_r1<8, 4>(a, save, st);
    case 12: return launch_fwd_r1<12, 4>(a, save, st);
    case 16: return launch_fwd_r1<16, 4>(a, save, st);
    default: return -3;
  }
}
int launch_bwd_r1_rx4(int RH_T, const SeqBwdArgs& a, const GradOut& o, cudaStream_t st) {
  switch (RH_T) {
    case 2: return launch_bwd_r1<2, 4>(a, o, st);
    case 4: return launch_bwd_r1<4, 4>(a, o, st);
    case 6: return launch_bwd_r1<6, 4>(a, o, st);
    case 8: return launch_bwd_r1<8, 4>(a, o, st);
    case 12: return launch_bwd_r1<12, 4>(a, o, st);
    case 16: return launch_bwd_r1<16, 4>(a, o, st);
    default: return -3;
  }
}
}  // namespace vmlmf
