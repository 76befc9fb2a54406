// generic.cuh -- regime G (any shape): placeholder until the per-step GEMM path lands.
#pragma once
#include "../../include/vmlmf_b200.h"
#include "common.cuh"

namespace vmlmf {

inline int generic_plan(int, int, int, int, int, int, vmlmf_plan*) { return VMLMF_EUNSUPPORTED; }
inline int generic_xproj(const float*, long long, long long, const float*, float*, int, int, int, int, int,
                         cudaStream_t) { return VMLMF_EUNSUPPORTED; }
template <class... Ts> inline int generic_seq_fwd(Ts...) { return VMLMF_EUNSUPPORTED; }
template <class... Ts> inline int generic_seq_bwd(Ts...) { return VMLMF_EUNSUPPORTED; }

}  // namespace vmlmf
