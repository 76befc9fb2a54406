// seq_bwd_mma.cu -- instantiations and host launcher of the warp-MMA R1 backward (seq_bwd_mma.cuh).
#include "seq_bwd_mma.cuh"
#include "seq_r1_launch.cuh"

namespace vmlmf {

namespace {
struct Shape { int KS, NZ, NW; };
Shape shape_of(int H, int RX, int RH) { return Shape{ceil_div(RH + RX + 1, 8), ceil_div(RH, 8), ceil_div(H, 16)}; }
long long n_blocks(int T, int B) { return (long long)T * ceil_div(B, 16); }
int grad_ctas(long long blocks) { return (int)(blocks < num_sms() ? blocks : num_sms()); }
}  // namespace

bool bwd_mma_fits(int I, int H, int RX, int RH) {
  if (H > 256 || (H & 3) || I > H || RH > 16) return false;
  const Shape s = shape_of(H, RX, RH);
  if (s.KS > 4 || s.NZ > 2 || (s.NZ == 1 && s.KS > 3) || (s.NZ == 2 && s.KS < 2)) return false;
  if (seq_bwd_mma_smem_bytes(s.NW, s.KS) > 227 * 1024) return false;
  if (grad_rows_smem_bytes(s.KS, I, RX) > 200 * 1024) return false;
  return true;
}

long long bwd_mma_workspace_floats(int T, int B, int I, int H, int RX, int RH) {
  if (!bwd_mma_fits(I, H, RX, RH)) return 0;
  const Shape s = shape_of(H, RX, RH);
  const long long rows = (long long)T * B;
  const GradLayout L(I, H, RX, RH);
  return (long long)frag_floats(T, B, H, 4) + rows * 8 * s.KS + (long long)grad_ctas(n_blocks(T, B)) * L.total;
}

template <int KS, int NZ>
static int launch_a(const SeqBwdMmaArgs& a, int NW, cudaStream_t st) {
  auto kern = seq_bwd_mma_kernel<KS, NZ>;
  const size_t smem = seq_bwd_mma_smem_bytes(NW, KS);
  static PerDevice attr;                                // the attribute is per device
  int& attr_done = attr.cur();
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_done = 1;
  }
  const int ntiles = ceil_div(a.B, 16);
  const int grid = ntiles < num_sms() ? ntiles : num_sms();
  kern<<<grid, NW * 32, smem, st>>>(a);
  return (int)cudaGetLastError();
}

template <int KS, int NT_MAX>
static int launch_b_t(const GradRowsArgs& g, int NW, int grid, cudaStream_t st) {
  auto kern = grad_rows_kernel<KS, NT_MAX>;
  const size_t smem = grad_rows_smem_bytes(KS, g.I, g.RX);
  static PerDevice attr;
  int& attr_done = attr.cur();
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_done = 1;
  }
  kern<<<grid, NW * 32, smem, st>>>(g);
  return (int)cudaGetLastError();
}

template <int KS>
static int launch_b(const GradRowsArgs& g, int NW, int grid, cudaStream_t st) {
  // <= 8 warps: compile for 256 threads so the wide (KS >= 3) accumulator sets stay in registers
  return NW <= 8 ? launch_b_t<KS, 256>(g, NW, grid, st) : launch_b_t<KS, 512>(g, NW, grid, st);
}

int launch_bwd_mma(const SeqBwdMmaArgs& a0, const GradRowsArgs& g0, const GradOut& out, void* workspace, int* n_parts,
                   cudaStream_t st) {
  const Shape s = shape_of(a0.H, a0.RX, a0.RH);
  const long long rows = (long long)a0.T * a0.B;
  const GradLayout L(g0.I, a0.H, a0.RX, a0.RH);
  float* ws = (float*)workspace;
  SeqBwdMmaArgs a = a0;
  GradRowsArgs g = g0;
  a.dpre = ws;
  a.dzc = ws + frag_floats(a.T, a.B, a.H, 4);
  g.dpre = a.dpre;
  g.dzc = a.dzc;
  g.partial = a.dzc + rows * 8 * s.KS;
  const long long nb = n_blocks(a.T, a.B);
  const int G = grad_ctas(nb);
  g.blocks_per_cta = (int)((nb + G - 1) / G);
  *n_parts = G;
  int rc = kMmaNoFit;
  if (s.NZ == 1) {
    if (s.KS == 1) rc = launch_a<1, 1>(a, s.NW, st);
    else if (s.KS == 2) rc = launch_a<2, 1>(a, s.NW, st);
    else if (s.KS == 3) rc = launch_a<3, 1>(a, s.NW, st);
  } else {
    if (s.KS == 2) rc = launch_a<2, 2>(a, s.NW, st);
    else if (s.KS == 3) rc = launch_a<3, 2>(a, s.NW, st);
    else if (s.KS == 4) rc = launch_a<4, 2>(a, s.NW, st);
  }
  if (rc) return rc;
  switch (s.KS) {
    case 1: rc = launch_b<1>(g, s.NW, G, st); break;
    case 2: rc = launch_b<2>(g, s.NW, G, st); break;
    case 3: rc = launch_b<3>(g, s.NW, G, st); break;
    case 4: rc = launch_b<4>(g, s.NW, G, st); break;
  }
  if (rc) return rc;
  reduce_partials_kernel<<<ceil_div(L.total, kReduceElems), 256, 0, st>>>(g.partial, G, L, out);
  return (int)cudaGetLastError();
}

}  // namespace vmlmf
