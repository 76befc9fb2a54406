// vmlmf_api.cu -- the C ABI (include/vmlmf_b200.h): argument checks, regime selection, launches.
#include "../../include/vmlmf_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "generic.cuh"
#include "seq_bwd_fused.cuh"
#include "seq_bwd_mma.cuh"
#include "seq_mma.cuh"
#include "seq_r1_launch.cuh"
#include "seq_r2_host.cuh"
#include "tail.cuh"
#include "xproj.cuh"

using namespace vmlmf;

namespace {

const int kMmaMinBatch = 1536;

// smallest compiled template rank >= r, or -1
int pick(const int* list, int n, int r) {
  for (int i = 0; i < n; ++i)
    if (list[i] >= r) return list[i];
  return -1;
}
const int kRH[] = {2, 4, 6, 8, 12, 16};
const int kRX[] = {4, 8, 16};

struct R1Choice { int rh_t, rx_t; bool ok; };

R1Choice choose_r1(int I, int H, int RX, int RH) {
  R1Choice c{pick(kRH, 6, RH), pick(kRX, 3, RX), false};
  c.ok = c.rh_t > 0 && c.rx_t > 0 && (c.rh_t + c.rx_t) <= 24 && H <= 256 && I <= H;
  return c;
}

// VMLMF_R1_SIMT=1 forces the SIMT R1 kernels (A/B measurements, tests); read at plan time
bool simt_only() {
  const char* e = getenv("VMLMF_R1_SIMT");
  return e && e[0] == '1';
}
// VMLMF_MMA_MIN_BATCH overrides the batch size from which the warp-MMA path is planned (tests force it to 1)
// VMLMF_BWD_SPLIT=1 selects the two-kernel R1M backward (dPre through HBM) instead of the fused one
bool use_fused_bwd(int I, int H, int RX, int RH) {
  const char* e = getenv("VMLMF_BWD_SPLIT");
  return !(e && e[0] == '1') && bwd_fused_fits(I, H, RX, RH);
}
int mma_min_batch() {
  const char* e = getenv("VMLMF_MMA_MIN_BATCH");
  return e ? atoi(e) : kMmaMinBatch;
}

// any of the (nullable) [B,H] state pointers not 8-byte aligned (the warp-MMA kernels move them as float2)
bool misaligned8(const void* a, const void* b, const void* c, const void* d, const void* e, const void* f) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(f)) & 7) != 0;
}

int check_dims(int T, int B, int I, int H, int RX, int RH) {
  if (T <= 0 || B <= 0 || I <= 0 || H <= 0 || RX <= 0 || RH <= 0) return VMLMF_EINVAL;
  if (H < I) return VMLMF_ESHAPE;
  return VMLMF_OK;
}

}  // namespace

extern "C" {

int vmlmf_abi_version(void) { return VMLMF_ABI_VERSION; }

const char* vmlmf_strerror(int code) {
  switch (code) {
    case VMLMF_OK: return "ok";
    case VMLMF_EINVAL: return "vmlmf: invalid argument (null pointer, non-positive size or bad stride)";
    case VMLMF_ESHAPE: return "vmlmf: hidden_size must be >= input_size";
    case VMLMF_EUNSUPPORTED: return "vmlmf: shape outside every compiled regime";
    case VMLMF_EWORKSPACE: return "vmlmf: workspace missing or too small";
    case VMLMF_EPLAN: return "vmlmf: plan does not match the arguments";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "vmlmf: unknown error";
}

int vmlmf_seq_plan(int T, int B, int I, int H, int RX, int RH, vmlmf_plan* plan) {
  if (!plan) return VMLMF_EINVAL;
  const int rc = check_dims(T, B, I, H, RX, RH);
  if (rc) return rc;
  *plan = vmlmf_plan{};
  plan->gates_bytes = (long long)T * B * 4 * H * (long long)sizeof(float);
  plan->cs_bytes = (long long)T * B * H * (long long)sizeof(float);
  // warp-MMA kernels tile 16 sequences per CTA: they win once the batch fills most SMs (measured crossover
  // on B200 between B = 1024 and 2048 at H = 256); below that the SIMT kernels (4 / 2
  // sequences per CTA) have the lower per-step latency.
  if (!simt_only() && B >= mma_min_batch() && fwd_mma_fits(I, H, RX, RH) && bwd_mma_fits(I, H, RX, RH)) {
    plan->path = VMLMF_PATH_R1M;
    plan->zx_pitch = round_up(RX, 4);
    plan->z_pitch = 8 * ceil_div(RH, 8);
    plan->gates_bytes = (long long)frag_floats(T, B, H, 4) * (long long)sizeof(float);
    plan->cs_bytes = (long long)frag_floats(T, B, H, 1) * (long long)sizeof(float);
    plan->bwd_workspace_bytes = bwd_mma_workspace_floats(T, B, I, H, RX, RH) * (long long)sizeof(float);
    if (use_fused_bwd(I, H, RX, RH)) {      // fused backward: no dPre buffer, only the reduced dzc rows and the partials
      plan->bwd_workspace_bytes = bwd_fused_workspace_floats(T, B, I, H, RX, RH) * (long long)sizeof(float);
      plan->reserved[2] = 1;                // the backward variant is fixed here: the workspace was sized for it
    }
    return VMLMF_OK;
  }
  const R1Choice c = choose_r1(I, H, RX, RH);
  if (c.ok) {
    plan->path = VMLMF_PATH_R1;
    plan->zx_pitch = round_up(c.rx_t, 4);
    plan->z_pitch = next_pow2(c.rh_t);
    plan->xp_cols = 0;
    plan->fwd_workspace_bytes = 0;
    const GradLayout L(I, H, RX, RH);
    const long long ntiles = ceil_div(B, r1_bwd_bt(B, c.rh_t));
    const long long cap = (long long)num_sms() * kMaxCtasPerSM;
    plan->bwd_workspace_bytes = (ntiles < cap ? ntiles : cap) * L.total * (long long)sizeof(float);
    plan->reserved[0] = c.rh_t;
    plan->reserved[1] = c.rx_t;
    return VMLMF_OK;
  }
  const int grc = generic_plan(T, B, I, H, RX, RH, plan);
  if (grc) return grc;
  // beyond the register-resident kernels: the persistent tcgen05 recurrence (one launch for all T steps) when TMA is
  // available; the launch-per-timestep generic regime otherwise.  Pitches and the saved-state layouts are shared.
  if (r3::fits(T, B, I, H, RX, RH)) {
    // small batches: weight-stationary variant (factors resident in shared memory, XP / dzx time-parallel around the launch)
    plan->path = VMLMF_PATH_R3;
    plan->xp_cols = 0;
    const r3::Geom g = r3::geom(T, B, I, H, RX, RH);
    plan->fwd_workspace_bytes = (g.fwd_floats + 128) * (long long)sizeof(float);
    plan->bwd_workspace_bytes = (g.bwd_floats + tp_scratch(T, B, I, H, g.Hp, RX, RH).total + 192) * (long long)sizeof(float);
  } else if (r2::fits(T, B, I, H, RX, RH)) {
    plan->path = VMLMF_PATH_R2;
    plan->xp_cols = 0;
    const r2::Geom g = r2::geom(T, B, I, H, RX, RH);
    plan->fwd_workspace_bytes = (g.fwd_floats + 64) * (long long)sizeof(float);
    plan->bwd_workspace_bytes = (g.bwd_floats + tp_scratch(T, B, I, H, g.Hp, RX, RH).total + 128) * (long long)sizeof(float);
  }
  return VMLMF_OK;
}

int vmlmf_diag_fwd(const float* u, const float* v, const float* dia, float* D, int n, int H, int R, void* stream) {
  if (!u || !v || !dia || !D || n <= 0 || H <= 0 || R <= 0) return VMLMF_EINVAL;
  if (n > H) return VMLMF_ESHAPE;
  const long long threads = 4LL * n * 32;
  diag_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(u, v, dia, D, n, H, R);
  return (int)cudaGetLastError();
}

int vmlmf_diag_bwd(const float* u, const float* v, const float* dD, float* du, float* dv, float* ddia, int n, int H,
                   int R, void* stream) {
  if (!u || !v || !dD || !du || !dv || !ddia || n <= 0 || H <= 0 || R <= 0) return VMLMF_EINVAL;
  if (n > H) return VMLMF_ESHAPE;
  const long long threads = 4LL * H * R;
  diag_bwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(u, v, dD, du, dv, ddia, n, H, R);
  return (int)cudaGetLastError();
}

int vmlmf_pack_plain_fwd(const float* u_x, const float* v_x, const float* dia_x, const float* u_h, const float* v_h,
                         const float* dia_h, const float* b_x, const float* b_h, float* Dx, float* Dh, float* bias,
                         int I, int H, int RX, int RH, void* stream) {
  if (!u_x || !v_x || !dia_x || !u_h || !v_h || !dia_h || !b_x || !b_h || !Dx || !Dh || !bias) return VMLMF_EINVAL;
  if (I <= 0 || H <= 0 || RX <= 0 || RH <= 0) return VMLMF_EINVAL;
  if (I > H) return VMLMF_ESHAPE;
  const long long threads = (4LL * I + 4LL * H + ceil_div(4 * H, 32)) * 32;
  pack_plain_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      u_x, v_x, dia_x, u_h, v_h, dia_h, b_x, b_h, Dx, Dh, bias, I, H, RX, RH);
  return (int)cudaGetLastError();
}

int vmlmf_pack_plain_bwd(const float* u_x, const float* v_x, const float* u_h, const float* v_h, const float* dDx,
                         const float* dDh, float* dUx, float* dVx, float* dA, float* dBm, float* ddia_x,
                         float* ddia_h, const float* dbias, float* db_h, int I, int H, int RX, int RH, void* stream) {
  if (!u_x || !v_x || !u_h || !v_h || !dDx || !dDh || !dUx || !dVx || !dA || !dBm || !ddia_x || !ddia_h || !dbias || !db_h)
    return VMLMF_EINVAL;
  if (I <= 0 || H <= 0 || RX <= 0 || RH <= 0) return VMLMF_EINVAL;
  if (I > H) return VMLMF_ESHAPE;
  const long long threads = 4LL * H * RX + 4LL * H * RH + 4LL * H;
  pack_plain_bwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      u_x, v_x, u_h, v_h, dDx, dDh, dUx, dVx, dA, dBm, ddia_x, ddia_h, dbias, db_h, I, H, RX, RH);
  return (int)cudaGetLastError();
}

int vmlmf_gemm_nt(const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc,
                  const float* bias, int M, int N, int K, int accumulate, void* workspace, long long workspace_bytes,
                  void* stream) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || lda < K || ldb < K || ldc < N) return VMLMF_EINVAL;
  return gemm_nt_public(A, lda, B, ldb, C, ldc, bias, M, N, K, accumulate, (float*)workspace,
                        workspace ? workspace_bytes / (long long)sizeof(float) : 0, (cudaStream_t)stream);
}

int vmlmf_gemm_tn(const float* At, long long lda, const float* Bt, long long ldb, float* C, long long ldc, int M, int N,
                  long long K, int accumulate, void* workspace, long long workspace_bytes, void* stream) {
  if (!At || !Bt || !C || M <= 0 || N <= 0 || K <= 0 || lda < M || ldb < N || ldc < N) return VMLMF_EINVAL;
  return gemm_tn_public(At, lda, Bt, ldb, C, ldc, M, N, K, accumulate, (float*)workspace,
                        workspace ? workspace_bytes / (long long)sizeof(float) : 0, (cudaStream_t)stream);
}

int vmlmf_xproj_fwd(const float* x, long long xs_t, long long xs_b, const float* Ux, float* zx, int T,
                    int B, int I, int RX, int zx_pitch, void* stream) {
  if (!x || !Ux || !zx || T <= 0 || B <= 0 || I <= 0 || RX <= 0) return VMLMF_EINVAL;
  if (zx_pitch < RX || (zx_pitch & 3)) return VMLMF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t uf_bytes = (size_t)ceil_div(I, 8) * ceil_div(zx_pitch, 8) * 32 * 4 * sizeof(float);
  const size_t tile_bytes = (size_t)kXprojRows * I * sizeof(float);
  if (zx_pitch <= 128 && I <= 512 && uf_bytes + tile_bytes <= 200 * 1024) {
    const int order = (xs_t == I && xs_b == (long long)T * I) ? 1 : ((xs_b == I && xs_t == (long long)B * I) ? 2 : 0);
    // three tiles in flight per block when the input is one contiguous block and three blocks still fit an SM
    const int stages = (order != 0 && 3 * (uf_bytes + 3 * tile_bytes + 1024) <= 227 * 1024) ? 3 : 1;
    const size_t smem = uf_bytes + stages * tile_bytes;
    static PerDevice attr;                                // the attribute is per device
    int& attr_done = attr.cur();
    if (!attr_done) {
      cudaError_t e = cudaFuncSetAttribute(xproj_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return (int)e;
      attr_done = 1;
    }
    const long long nrows = (long long)T * B;
    long long grid = (nrows + kXprojRows - 1) / kXprojRows;
    // resident blocks per SM: one wave, every block walks its share.  The query is a driver call: cached per
    // shared-memory size (256-byte buckets; a benign race, every writer stores the same value)
    static PerDevice occ_cache[1024];
    int& occ_slot = occ_cache[smem / 256 < 1023 ? smem / 256 : 1023].cur();
    if (occ_slot == 0) {
      int q = 1;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, xproj_small_kernel, 128, smem) != cudaSuccess || q < 1) q = 1;
      occ_slot = q;
    }
    const int occ = occ_slot;
    if (grid > (long long)num_sms() * occ) grid = (long long)num_sms() * occ;
    xproj_small_kernel<<<(int)grid, 128, smem, st>>>(x, xs_t, xs_b, Ux, zx, T, B, I, RX, zx_pitch, order, stages);
    return (int)cudaGetLastError();
  }
  return generic_xproj(x, xs_t, xs_b, Ux, zx, T, B, I, RX, zx_pitch, st);
}

int vmlmf_seq_fwd(const vmlmf_plan* plan, const float* x, long long xs_t, long long xs_b, const float* zx,
                  const float* Ux, const float* Vx, const float* Dx, const float* A, const float* Bm,
                  const float* Dh, const float* bias, const float* h0, const float* c0, float* y,
                  long long ys_t, long long ys_b, float* hT, float* cT, float* gates, float* cs, float* z,
                  void* workspace, int T, int B, int I, int H, int RX, int RH, void* stream) {
  if (!plan || !x || !zx || !Ux || !Vx || !Dx || !A || !Bm || !Dh || !bias || !hT || !cT) return VMLMF_EINVAL;
  const int rc = check_dims(T, B, I, H, RX, RH);
  if (rc) return rc;
  const bool save = gates || cs || z;
  if (save && !(gates && cs && z)) return VMLMF_EINVAL;
  // y may be null for a last-step-only caller (V/models/vmlmf.py:354-355 reads y[:, -1] alone): inference on the
  // persistent kernels only -- backward and the generic regime read h_{t-1} back from y
  if (!y && (save || plan->path == VMLMF_PATH_G || plan->path == VMLMF_PATH_R2 || plan->path == VMLMF_PATH_R3)) return VMLMF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (plan->path == VMLMF_PATH_R1) {
    const R1Choice c = choose_r1(I, H, RX, RH);
    if (!c.ok || plan->zx_pitch != round_up(c.rx_t, 4) || plan->z_pitch != next_pow2(c.rh_t)) return VMLMF_EPLAN;
    SeqFwdArgs a{x, xs_t, xs_b, zx, Vx, Dx, A, Bm, Dh, bias, h0, c0, y, ys_t, ys_b, hT, cT, gates, cs, z,
                 T, B, I, H, RX, RH};
    switch (c.rx_t) {
      case 4: return launch_fwd_r1_rx4(c.rh_t, a, save, st);
      case 8: return launch_fwd_r1_rx8(c.rh_t, a, save, st);
      case 16: return launch_fwd_r1_rx16(c.rh_t, a, save, st);
    }
    return VMLMF_EUNSUPPORTED;
  }
  if (plan->path == VMLMF_PATH_R1M) {
    if (!fwd_mma_fits(I, H, RX, RH) || plan->z_pitch != 8 * ceil_div(RH, 8) || plan->zx_pitch != round_up(RX, 4))
      return VMLMF_EPLAN;
    if ((ys_t & 1) || (ys_b & 1) || (reinterpret_cast<uintptr_t>(y) & 7)) return VMLMF_EINVAL;
    if (misaligned8(h0, c0, hT, cT, nullptr, nullptr)) return VMLMF_EINVAL;   // read / written as float2
    SeqFwdArgs a{x, xs_t, xs_b, zx, Vx, Dx, A, Bm, Dh, bias, h0, c0, y, ys_t, ys_b, hT, cT, gates, cs, z,
                 T, B, I, H, RX, RH};
    return launch_fwd_mma(SeqFwdMmaArgs{a, plan->z_pitch, plan->zx_pitch}, save, st);
  }
  if (plan->path == VMLMF_PATH_R2) {
    if (!r2::fits(T, B, I, H, RX, RH) || plan->zx_pitch != round_up(RX, 4) || plan->z_pitch != round_up(RH, 4)) return VMLMF_EPLAN;
    if (!workspace) return VMLMF_EWORKSPACE;
    r2::FwdCall c{x, xs_t, xs_b, zx, Vx, Dx, A, Bm, Dh, bias, h0, c0, y, ys_t, ys_b, hT, cT, gates, cs, z, T, B, I, H, RX, RH};
    return r2::launch_fwd(c, workspace, st);
  }
  if (plan->path == VMLMF_PATH_R3) {
    if (!r3::fits(T, B, I, H, RX, RH) || plan->zx_pitch != round_up(RX, 4) || plan->z_pitch != round_up(RH, 4)) return VMLMF_EPLAN;
    if (!workspace) return VMLMF_EWORKSPACE;
    const r3::Geom g = r3::geom(T, B, I, H, RX, RH);
    float* xp = r3::ws_base(workspace) + g.o_xp;
    // time-parallel: XP = ZX Vx^T + bias + x (.) Dx (tcgen05 3xTF32 GEMM; SIMT when an operand misses the TMA constraints)
    const int rows = T * B;
    int rc3 = tc::gemm_tc(zx, plan->zx_pitch, Vx, RX, rows, 4 * H, RX, tc::EpiXPTC{xp, bias, x, xs_t, xs_b, B, Dx, H, I}, st);
    if (rc3 == tc::kTcNoFit)
      rc3 = gemm_launch<false, true>(plain_view(zx, plan->zx_pitch), plain_view(Vx, RX), rows, 4 * H, RX, 1, NIdent{},
                                     EpiXP{xp, bias, tb_view(x, xs_t, xs_b, B), Dx, H, I}, st);
    if (rc3) return rc3;
    r3::FwdCall c{xp, A, Bm, Dh, h0, c0, y, ys_t, ys_b, hT, cT, gates, cs, z, T, B, I, H, RX, RH};
    return r3::launch_fwd(c, workspace, st);
  }
  if (plan->path == VMLMF_PATH_G)
    return generic_seq_fwd(plan, x, xs_t, xs_b, zx, Ux, Vx, Dx, A, Bm, Dh, bias, h0, c0, y, ys_t, ys_b, hT, cT,
                           gates, cs, z, workspace, T, B, I, H, RX, RH, st);
  return VMLMF_EPLAN;
}

int vmlmf_seq_bwd(const vmlmf_plan* plan, const float* x, long long xs_t, long long xs_b, const float* zx,
                  const float* Ux, const float* Vx, const float* Dx, const float* A, const float* Bm,
                  const float* Dh, const float* h0, const float* c0, const float* y, long long ys_t,
                  long long ys_b, const float* gates, const float* cs, const float* z, const float* dy,
                  long long dys_t, long long dys_b, const float* dhT, const float* dcT, float* dx,
                  long long dxs_t, long long dxs_b, float* dh0, float* dc0, float* dUx, float* dVx,
                  float* dDx, float* dA, float* dBm, float* dDh, float* dbias, void* workspace, int T, int B,
                  int I, int H, int RX, int RH, void* stream) {
  if (!plan || !x || !zx || !Ux || !Vx || !Dx || !A || !Bm || !Dh || !y || !gates || !cs || !z) return VMLMF_EINVAL;
  if (!dUx || !dVx || !dDx || !dA || !dBm || !dDh || !dbias) return VMLMF_EINVAL;
  const int rc = check_dims(T, B, I, H, RX, RH);
  if (rc) return rc;
  if (plan->bwd_workspace_bytes > 0 && !workspace) return VMLMF_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  if (plan->path == VMLMF_PATH_R1) {
    const R1Choice c = choose_r1(I, H, RX, RH);
    if (!c.ok || plan->zx_pitch != round_up(c.rx_t, 4) || plan->z_pitch != next_pow2(c.rh_t)) return VMLMF_EPLAN;
    SeqBwdArgs a{x, xs_t, xs_b, zx, Ux, Vx, Dx, A, Bm, Dh, h0, c0, y, ys_t, ys_b, gates, cs, z,
                 dy, dys_t, dys_b, dhT, dcT, dx, dxs_t, dxs_b, dh0, dc0, (float*)workspace, T, B, I, H, RX, RH};
    GradOut o{dUx, dVx, dDx, dA, dBm, dDh, dbias};
    switch (c.rx_t) {
      case 4: return launch_bwd_r1_rx4(c.rh_t, a, o, st);
      case 8: return launch_bwd_r1_rx8(c.rh_t, a, o, st);
      case 16: return launch_bwd_r1_rx16(c.rh_t, a, o, st);
    }
    return VMLMF_EUNSUPPORTED;
  }
  if (plan->path == VMLMF_PATH_R1M) {
    if (!bwd_mma_fits(I, H, RX, RH) || plan->z_pitch != 8 * ceil_div(RH, 8) || plan->zx_pitch != round_up(RX, 4))
      return VMLMF_EPLAN;
    if ((ys_t & 1) || (ys_b & 1) || (reinterpret_cast<uintptr_t>(y) & 7)) return VMLMF_EINVAL;
    GradOut o2{dUx, dVx, dDx, dA, dBm, dDh, dbias};
    if (misaligned8(h0, c0, dhT, dcT, dh0, dc0)) return VMLMF_EINVAL;      // read / written as float2
    if (plan->reserved[2] == 1) {
      if (!bwd_fused_fits(I, H, RX, RH)) return VMLMF_EPLAN;
      // fused: recurrence + weight-gradient accumulation (accumulators in tensor memory) + dX; then dA / dUx
      SeqBwdFusedArgs fa{gates, cs, c0, dy, dys_t, dys_b, dhT, dcT, Ux, Vx, Dx, A, Bm, Dh, x, xs_t, xs_b, y, ys_t, ys_b, h0,
                         z, zx, plan->z_pitch, plan->zx_pitch, nullptr, 0, dx, dxs_t, dxs_b, dh0, dc0, nullptr, T, B, I, H, RX, RH};
      return launch_bwd_fused(fa, o2, workspace, st);
    }
    // reverse-time recurrence (K3a) + time-parallel gradient accumulation (K3b)
    SeqBwdMmaArgs ba{gates, cs, c0, dy, dys_t, dys_b, dhT, dcT, Vx, A, Bm, Dh, nullptr, nullptr, dh0, dc0, T, B, H, RX, RH};
    GradRowsArgs gr{nullptr, nullptr, z, zx, plan->z_pitch, plan->zx_pitch, y, ys_t, ys_b, h0, x, xs_t, xs_b, Ux, Dx,
                    dx, dxs_t, dxs_b, nullptr, T, B, I, H, RX, RH, 0};
    GradOut o{dUx, dVx, dDx, dA, dBm, dDh, dbias};
    int nparts = 0;
    return launch_bwd_mma(ba, gr, o, workspace, &nparts, st);
  }
  if (plan->path == VMLMF_PATH_R2) {
    if (!r2::fits(T, B, I, H, RX, RH) || plan->zx_pitch != round_up(RX, 4) || plan->z_pitch != round_up(RH, 4)) return VMLMF_EPLAN;
    // persistent reverse-time recurrence, then the contractions over all T*B rows
    r2::BwdCall bc{Vx, A, Bm, Dh, c0, gates, cs, dy, dys_t, dys_b, dhT, dcT, dh0, dc0, T, B, I, H, RX, RH};
    r2::BwdOut bo;
    int rc2 = r2::launch_bwd(bc, workspace, &bo, st);
    if (rc2) return rc2;
    const TpScratch ts = tp_scratch(T, B, I, H, bo.G, RX, RH);
    float* part = align4(bo.after);
    TpArgs tp{bo.dpre, bo.G, z, plan->z_pitch, zx, plan->zx_pitch, bo.dz, bo.dzx, true, x, xs_t, xs_b, y, ys_t, ys_b, h0, Ux, Vx, Dx,
              dx, dxs_t, dxs_b, dUx, dVx, dDx, dA, dBm, dDh, dbias, part, ts.n_part, nullptr, nullptr, nullptr, nullptr, ts.ldt, nullptr,
              T, B, I, H, RX, RH, true};
    tp.tA = align4(part + ts.n_part);
    tp.tB = align4(tp.tA + ts.n_tA);
    tp.gtmp = align4(tp.tB + ts.n_tB);
    return generic_bwd_tp(tp, st);
  }
  if (plan->path == VMLMF_PATH_R3) {
    if (!r3::fits(T, B, I, H, RX, RH) || plan->zx_pitch != round_up(RX, 4) || plan->z_pitch != round_up(RH, 4)) return VMLMF_EPLAN;
    r3::BwdCall bc{Vx, A, Bm, Dh, c0, gates, cs, dy, dys_t, dys_b, dhT, dcT, dh0, dc0, T, B, I, H, RX, RH};
    r3::BwdOut bo;
    int rc3 = r3::launch_bwd(bc, workspace, &bo, st);
    if (rc3) return rc3;
    const TpScratch ts = tp_scratch(T, B, I, H, bo.G, RX, RH);
    float* part = align4(bo.after);
    TpArgs tp{bo.dpre, bo.G, z, plan->z_pitch, zx, plan->zx_pitch, bo.dz, bo.dzx, false, x, xs_t, xs_b, y, ys_t, ys_b, h0, Ux, Vx, Dx,
              dx, dxs_t, dxs_b, dUx, dVx, dDx, dA, dBm, dDh, dbias, part, ts.n_part, bo.vxt, nullptr, nullptr, nullptr, ts.ldt, nullptr,
              T, B, I, H, RX, RH, true};
    tp.tA = align4(part + ts.n_part);
    tp.tB = align4(tp.tA + ts.n_tA);
    tp.gtmp = align4(tp.tB + ts.n_tB);
    tp.vxt_padded = true;
    return generic_bwd_tp(tp, st);
  }
  if (plan->path == VMLMF_PATH_G)
    return generic_seq_bwd(plan, x, xs_t, xs_b, zx, Ux, Vx, Dx, A, Bm, Dh, h0, c0, y, ys_t, ys_b, gates, cs, z,
                           dy, dys_t, dys_b, dhT, dcT, dx, dxs_t, dxs_b, dh0, dc0, dUx, dVx, dDx, dA, dBm, dDh,
                           dbias, workspace, T, B, I, H, RX, RH, st);
  return VMLMF_EPLAN;
}

// ---------------- callers either side of the recurrence (tail.cuh) ----------------

long long vmlmf_softmax_nll_workspace_bytes(long long rows, int C) {
  if (rows <= 0 || C <= 0) return 0;
  return (C <= 2048 ? (rows + 7) / 8 : rows) * (long long)sizeof(float);
}

int vmlmf_softmax_nll_fwd(const float* scores, long long ld, const long long* labels, float* lse, float* loss,
                          float scale, void* workspace, long long rows, int C, void* stream) {
  if (!scores || !labels || !lse || !loss || !workspace || rows <= 0 || C <= 0 || ld < C) return VMLMF_EINVAL;
  if (rows > 0x7fffffffLL) return VMLMF_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  float* partials = (float*)workspace;
  long long nparts;
  if (C <= 2048) {
    nparts = (rows + 7) / 8;
    softmax_nll_fwd_kernel<true><<<(unsigned)nparts, kTailThreads, 0, st>>>(scores, ld, labels, lse, partials, rows, C);
  } else {
    nparts = rows;
    softmax_nll_fwd_kernel<false><<<(unsigned)nparts, kTailThreads, 0, st>>>(scores, ld, labels, lse, partials, rows, C);
  }
  sum_scale_kernel<<<1, kTailThreads, 0, st>>>(partials, nparts, scale, loss);
  return (int)cudaGetLastError();
}

int vmlmf_softmax_nll_bwd(const float* scores, long long ld, const long long* labels, const float* lse,
                          const float* dloss, float scale, float* dscores, long long ldd, long long rows, int C,
                          void* stream) {
  if (!scores || !labels || !lse || !dscores || rows <= 0 || C <= 0 || ld < C || ldd < C) return VMLMF_EINVAL;
  if (rows > 0x7fffffffLL) return VMLMF_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (C <= 2048)
    softmax_nll_bwd_kernel<true><<<(unsigned)((rows + 7) / 8), kTailThreads, 0, st>>>(scores, ld, labels, lse, dloss, scale,
                                                                                     dscores, ldd, rows, C);
  else
    softmax_nll_bwd_kernel<false><<<(unsigned)rows, kTailThreads, 0, st>>>(scores, ld, labels, lse, dloss, scale, dscores,
                                                                          ldd, rows, C);
  return (int)cudaGetLastError();
}

namespace {
inline int head_np(int N) { return N <= 8 ? 8 : (N <= 16 ? 16 : 32); }
inline int head_rows_per_block(int B) {
  const int r = ceil_div(B, 2 * num_sms());
  return r < 8 ? 8 : (r > 256 ? 256 : r);
}
}  // namespace

int vmlmf_head_fwd(const float* h, long long ldh, const float* W, const float* bias, float* out, int B, int K, int N,
                   void* stream) {
  if (!h || !W || !out || B <= 0 || K <= 0 || N <= 0 || ldh < K) return VMLMF_EINVAL;
  if (N > 32 || K > 1024) return VMLMF_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int NP = head_np(N);
  const size_t smem = (size_t)NP * K * sizeof(float);
  int grid = ceil_div(ceil_div(B, 2), kTailThreads / 32);
  if (grid > 2 * num_sms()) grid = 2 * num_sms();
#define VMLMF_HEAD_FWD(NP_)                                                                                          \
  {                                                                                                                  \
    if (smem > 48 * 1024) {                                                                                          \
      cudaError_t e = cudaFuncSetAttribute(head_fwd_kernel<NP_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return (int)e;                                                                           \
    }                                                                                                                \
    head_fwd_kernel<NP_><<<grid, kTailThreads, smem, st>>>(h, ldh, W, bias, out, B, K, N);                           \
  }
  if (NP == 8) VMLMF_HEAD_FWD(8) else if (NP == 16) VMLMF_HEAD_FWD(16) else VMLMF_HEAD_FWD(32)
#undef VMLMF_HEAD_FWD
  return (int)cudaGetLastError();
}

long long vmlmf_head_bwd_workspace_bytes(int B, int K, int N) {
  if (B <= 0 || K <= 0 || N <= 0 || N > 32) return 0;
  const int NP = head_np(N);
  return (long long)ceil_div(B, head_rows_per_block(B)) * (NP * K + NP) * (long long)sizeof(float);
}

int vmlmf_head_bwd(const float* h, long long ldh, const float* W, const float* dout, float* dh, long long lddh,
                   float* dW, float* db, void* workspace, int B, int K, int N, void* stream) {
  if (!h || !W || !dout || !dW || !workspace || B <= 0 || K <= 0 || N <= 0 || ldh < K || (dh && lddh < K))
    return VMLMF_EINVAL;
  if (N > 32 || K > 1024) return VMLMF_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int NP = head_np(N), rpb = head_rows_per_block(B), grid = ceil_div(B, rpb);
  const size_t smem = (size_t)rpb * NP * sizeof(float);
  const int KJ = ceil_div(K, kTailThreads);
  float* partials = (float*)workspace;
#define VMLMF_HEAD_BWD(NP_, KJ_) \
  head_bwd_kernel<NP_, KJ_><<<grid, kTailThreads, smem, st>>>(h, ldh, W, dout, dh, lddh, partials, B, K, N, rpb)
#define VMLMF_HEAD_BWD_NP(NP_)                                    \
  {                                                               \
    if (KJ == 1) VMLMF_HEAD_BWD(NP_, 1);                          \
    else if (KJ == 2) VMLMF_HEAD_BWD(NP_, 2);                     \
    else VMLMF_HEAD_BWD(NP_, 4);                                  \
  }
  if (NP == 8) VMLMF_HEAD_BWD_NP(8) else if (NP == 16) VMLMF_HEAD_BWD_NP(16) else VMLMF_HEAD_BWD_NP(32)
#undef VMLMF_HEAD_BWD_NP
#undef VMLMF_HEAD_BWD
  head_reduce_kernel<<<ceil_div(N * K + N, 16), kTailThreads, 0, st>>>(partials, grid, NP, K, N, dW, db);
  return (int)cudaGetLastError();
}

int vmlmf_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                    float eps, const float* step_dev, int step, void* stream) {
  if (!p || !g || !m || !v || n <= 0 || (!step_dev && step < 1)) return VMLMF_EINVAL;
  adam_kernel<<<tail_grid(n), kTailThreads, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, step_dev, step);
  return (int)cudaGetLastError();
}

int vmlmf_embed_dropout_fwd(const long long* tok, const float* W, const unsigned char* mask, float scale, float* out,
                            long long ldo, long long rows, int E, int V, void* stream) {
  if (!tok || !W || !out || rows <= 0 || E <= 0 || V <= 0 || ldo < E) return VMLMF_EINVAL;
  embed_dropout_kernel<<<tail_grid(rows * ldo), kTailThreads, 0, (cudaStream_t)stream>>>(tok, W, mask, scale, out, ldo, rows, E, V);
  return (int)cudaGetLastError();
}

int vmlmf_p2p_adam_step(float* p, float* m, float* v, const float* const* peer_grads, int world, long long n, float scale,
                        float lr, float beta1, float beta2, float eps, const float* step_dev, int step, void* stream) {
  if (!p || !m || !v || !peer_grads || world < 1 || world > 16 || n <= 0 || (!step_dev && step < 1)) return VMLMF_EINVAL;
  p2p_adam_kernel<<<tail_grid(n), kTailThreads, 0, (cudaStream_t)stream>>>(p, m, v, peer_grads, world, n, scale, lr, beta1, beta2,
                                                                          eps, step_dev, step);
  return (int)cudaGetLastError();
}

long long vmlmf_sgd_clip_workspace_bytes(long long n) { return n <= 0 ? 0 : (long long)tail_grid(n) * (long long)sizeof(float); }

int vmlmf_sgd_clip_step(float* p, float* g, long long n, float lr, float max_norm, int scale_grads, float* norm_out,
                        void* workspace, void* stream) {
  if (!p || !g || !workspace || n <= 0) return VMLMF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = tail_grid(n);
  float* partials = (float*)workspace;
  sumsq_kernel<<<grid, kTailThreads, 0, st>>>(g, n, partials);
  sgd_clip_kernel<<<grid, kTailThreads, 0, st>>>(p, g, n, lr, max_norm, partials, grid, scale_grads, norm_out);
  return (int)cudaGetLastError();
}

}  // extern "C"
