// seq_r3.cuh -- regime R3: the persistent tcgen05 recurrence for SMALL batches (B <= 32), weight-stationary.
//
// The LM at the reference's own batch (V/lm_test.py: 20 streams, H=650, ranks 300; time loop V/models/vmlmf_lm.py:272-280,
// step :222-269) is a latency chain, not a throughput problem: one 128-row tensor-core tile is 84 % padding, and in regime
// R2 every CTA re-streams its 0.6 MB slice of the factors through TMA on every timestep.  Here the factors never move:
//   * ONE group of CS = ceil(H / 8) CTAs (cooperative launch, one per SM) owns the whole batch for all T steps; CTA s owns
//     the 8 hidden units [8s, 8s+8) and keeps ITS rows of the factors in shared memory for the whole call
//     (forward: Bm rows of its 4 x 8 gate outputs [32 x RH] + its 8 rows of A; backward: the same two slices transposed),
//     as tf32 hi / lo pairs: ~110 KB per CTA at the LM shape, 9 MB over the group;
//   * per step only activations move: the CTA's [32 x 8] slice of h (phase Z), the reduced z [32 x RH] (phase G).  They
//     live in global memory tile-major ([K tile][hi | lo][32 rows][32 fp32]), so one TMA box brings a group of K tiles;
//     a tile is 32 rows hi + 32 rows lo (the tensor core still computes 128-row tiles: rows 0..31 = hi, 32..63 = lo of the
//     SAME instruction -- see "ONE MMA per k-step" below --, rows 64..127 are whatever follows in shared memory and only reach
//     accumulator lanes nobody reads);
//   * the x side is taken out of the serial loop: XP = zx Vx^T + bias + x (.) Dx is one time-parallel tcgen05 GEMM before the
//     launch (700 rows at the LM batch), the gate epilogue adds it; likewise dzx = dPre Vx after the backward launch.
// Per timestep and CTA (3xTF32, accumulators in tensor memory):
//   phase Z   partial z_s[32, RH] = h_{t-1}[:, 8 units] * A[8 units, :]          1 k-step, N = RH in 128-column chunks
//   exchange  z = sum_s z_s: partials through L2, fixed-order reduce spread over ALL CTAs of the group (each CTA sums a
//             1/CS share of the [B, RH] outputs: 8 lanes x (CS/8) partials each, then a butterfly), two group barriers
//   phase G   pre[32, 4 x 8] = z * Bm[4 x 8 rows, :]^T                            K = RH, N = 32
// Backward mirrors it: phase 1 (dz partial = dPre_t[:, 4 x 8] * Bm[4 x 8 rows, :], K = 32), exchange, phase 2
// (dh_{t-1}[32, 8] = dz_t * A[8 units, :]^T, K = RH, N = 16), gate-gradient algebra of step t-1 in the epilogue.
// Proxy ordering: every thread that writes an operand tile with ordinary stores executes fence.proxy.async before the (CTA or
// group) barrier that precedes the TMA read -- the writer-side fence is the documented pattern; the producer adds none.
// Warp roles as in seq_r2.cuh: warp 0 TMA producer, warp 1 MMA issuer (both run warp-uniform and elect one lane per asynchronous
// instruction), warps 2-9 epilogue -- warps 4 and 8 own tensor-memory lanes 0..31 (hi rows: hi*hi + hi*lo), warps 5 and 9 lanes
// 32..63 (lo rows: lo*hi, handed to warps 4 / 8 through shared memory); all eight help with the reduce.
#pragma once
#include "seq_r2.cuh"

namespace vmlmf {
namespace r3 {

using namespace r2;

// Debug-only cycle trace (-DVMLMF_R2_TRACE): lane 0 of a warp of CTA 0 appends (event, clock) pairs to the warp's own slice of
// the trace buffer -- a plain store, no atomics, so that tracing does not stretch the latency chain it measures.
#ifdef VMLMF_R2_TRACE
#define R3_TRACE(ev)                                                                                           \
  do {                                                                                                         \
    if (blockIdx.x == 0 && lane == 0 && t >= 8 && t < 11 && trc < 96) {                                        \
      const int i__ = warp * 96 + trc++;                                                                       \
      g_r2_trace[2 * i__] = (ev) + 1000 * t; g_r2_trace[2 * i__ + 1] = clock64();                              \
    }                                                                                                          \
  } while (0)
#else
#define R3_TRACE(ev) do {} while (0)
#endif

constexpr int UB = 8;                        // hidden units per CTA
constexpr int RB = 32;                       // batch rows (one tensor-memory lane quarter, one TMA box)
constexpr int kATile = RB * BK * 4;          // 4 KB: [32 rows x 32 fp32] activation tile (hi or lo)
constexpr int kAStage = 2 * kATile;          // one K tile of an activation: hi | lo
constexpr int kTileFloats = kAStage / 4;     // 2048
constexpr int kPTile = BM * BK * 4;          // 16 KB: [128 rows x 32 fp32] weight tile
constexpr int kApTile = 16 * BK * 4;         // 2 KB: [16 rows x 32 fp32] (phase 2 B operand: 8 units + 8 spare rows)
constexpr int kMaxStages = 4;
constexpr int kBarBytes = 512;
constexpr int kXbufBytes = 4 * 32 * 32 * 4;  // lo*hi hand-over between the lane quarters: 2 warp pairs x 2 buffers x [32 x 32] fp32
constexpr int kSmemMax = 232448 - 1024;      // 227 KB minus the alignment slack

struct Bars3 {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t accf[2];
  uint64_t acce[2];
  uint64_t wbar;               // stationary weights landed
  uint32_t tmem_slot;
};
static_assert(sizeof(Bars3) <= kBarBytes, "barrier block");

// Activations that feed the tensor core live in global memory TILE-MAJOR: [K tile][hi | lo][32 rows][32 fp32] -- exactly the
// bytes of the shared-memory stage, so that ONE TMA box (32 columns x 64 * G rows of a [tiles * 64, 32] view) brings G K tiles
// with one barrier.  (A barrier wait + tcgen05 fence + commit costs ~250 cycles of the single issuing thread per use, an MMA
// ~50: tools/ubench_mma.cu.  Per-K-tile barriers made the MMA phases 3x longer than their instructions.)
// element (plane p, row r, column c) of a row-major [32, K] activation:
__host__ __device__ __forceinline__ size_t tile_off(int p, int r, int c) { return ((size_t)(c >> 5) * 2 + p) * 1024 + r * 32 + (c & 31); }
// (gate k, unit j) inside the K = 32 slice of CTA j / 8 (phase 1 A operand; slice-major weight rows use the same order)
__host__ __device__ __forceinline__ int slice_col(int k, int j) { return k * 8 + (j & 7); }

// Fixed-order sum of the NP partial copies of a [rows, 4 * w4] matrix (partial q at part + q * RB * pitch).  Output element e
// (one float4) belongs to CTA e / epc; inside the CTA eight adjacent lanes share one output, lane pg adding the partials
// q = pg, pg + 8, ... in order, then a butterfly over the eight lanes (same order every run: deterministic).
template <class Store, class Trace>
__device__ __forceinline__ void reduce_partials(const float* part, int NP, int CS, int s_rank, int rows, int pitch, int w4, Store&& store, Trace&& tr) {
  const int et = (int)threadIdx.x - 64;
  const int o = et >> 3, pg = et & 7;
  const int E = rows * w4;
  const int epc = (E + CS - 1) / CS;
  const int e0 = s_rank * epc;
  const int cnt = min(epc, E - e0);
  const size_t qstride4 = (size_t)RB * pitch / 4;
  for (int base = 0; base < cnt; base += 32) {
    const int oo = base + o;
    const bool on = oo < cnt;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int r = 0, c4 = 0;
    if (on) {
      const int e = e0 + oo;
      r = e / w4;
      c4 = e - r * w4;
      const float4* src = reinterpret_cast<const float4*>(part + (size_t)r * pitch) + c4;
      for (int q = pg; q < NP; q += 96) {
        float4 pv[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) pv[i] = q + 8 * i < NP ? __ldcg(src + (size_t)(q + 8 * i) * qstride4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 12; ++i) { acc.x += pv[i].x; acc.y += pv[i].y; acc.z += pv[i].z; acc.w += pv[i].w; }
      }
    }
    tr(33);
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, d);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, d);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, d);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, d);
    }
    if (on && pg == 0) store(r, c4 * 4, acc);
  }
}
// v and its tf32 lo part into the tile-major operand buffer (4 consecutive columns never straddle a tile: c % 4 == 0)
__device__ __forceinline__ void store_split4(float* op, int r, int c, const float4 v) {
  float4 lo;
  lo.x = split_lo(v.x, v.x); lo.y = split_lo(v.y, v.y); lo.z = split_lo(v.z, v.z); lo.w = split_lo(v.w, v.w);
  *reinterpret_cast<float4*>(op + tile_off(0, r, c)) = v;           // the tensor core truncates: the value is its own hi part
  *reinterpret_cast<float4*>(op + tile_off(1, r, c)) = lo;
}
// 32 lanes x 32 consecutive fp32 columns of the accumulator -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32_raw(uint32_t taddr, float (&v)[32]) {
  float g0[8], g1[8], g2[8], g3[8];
  tc::tmem_ld8_raw(taddr, g0); tc::tmem_ld8_raw(taddr + 8, g1); tc::tmem_ld8_raw(taddr + 16, g2); tc::tmem_ld8_raw(taddr + 24, g3);
  tc::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = g0[i]; v[8 + i] = g1[i]; v[16 + i] = g2[i]; v[24 + i] = g3[i]; }
}
// The producer and the MMA warp run their loops with all 32 lanes (warp-uniform control flow, warp-uniform operands) and
// elect one lane right at each asynchronous instruction.  Inside an `if (lane == 0)` region the compiler keeps descriptors
// and addresses in per-thread registers and wraps EVERY tcgen05.mma / TMA in an elect + five R2UR.BROADCAST + loop sequence
// (~135 cycles per MMA measured, against ~45 for the uniform form: tools/ubench_mma.cu).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_elect(uint32_t acc, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  if (elect_one()) mma_tf32_ss(acc, da, db, idesc, accumulate);
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
  if (elect_one()) mma_commit(bar);
}
// ONE MMA per tf32 k-step computes all three 3xTF32 products.  The batch fills 32 of the 128 accumulator rows and every operand
// lives in shared memory as a hi tile directly followed by its lo tile, so
//   * the A descriptor at the hi tile makes rows 0..31 = a_hi and rows 32..63 = a_lo (rows 64..127: don't care);
//   * the B descriptor at the hi tile with N doubled makes columns [0, n) = b_hi and [NB, NB + n) = b_lo (NB = rows of the hi tile).
// Accumulator: lanes 0..31 x [0, n) = hi*hi, lanes 0..31 x [NB, NB+n) = hi*lo, lanes 32..63 x [0, n) = lo*hi (lo*lo is computed
// and ignored).  The three terms stay in separate accumulator cells as before (the tensor core's accumulate truncates: the
// small terms must not ride on the large sum).
__device__ __forceinline__ void issue_ktile(uint32_t a_addr, uint32_t b_addr, uint32_t acc, uint32_t idesc, int ksteps, bool first) {
  const uint64_t da = make_desc(a_addr), db = make_desc(b_addr);
  if (elect_one()) {                                         // one election per K tile: elect.sync itself is ~25 cycles
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < ksteps) mma_tf32_ss(acc, da + 2 * k, db + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
  }
}
// K-step k of every K tile of the long contractions (K = RH) goes to its own accumulator (column offset k * nw): four
// independent chains, summed by the epilogue.
__device__ __forceinline__ void issue_ktile_split(uint32_t a_addr, uint32_t b_addr, uint32_t acc, uint32_t nw, uint32_t idesc,
                                                  int ksteps, bool first) {
  const uint64_t da = make_desc(a_addr), db = make_desc(b_addr);
  if (elect_one()) {
    if (ksteps == 4) {                                       // every tile but the last: straight-line
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_tf32_ss(acc + k * nw, da + 2 * k, db + 2 * k, idesc, first ? 0u : 1u);
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (k < ksteps) mma_tf32_ss(acc + k * nw, da + 2 * k, db + 2 * k, idesc, first ? 0u : 1u);
    }
  }
}
// One 32-column pass of a partial-product chunk.  Lane quarter 0 (warps 4, 8) holds hi*hi + hi*lo of row `lane`, lane quarter 1
// (warps 5, 9) the lo*hi term of the same row: the quarter-1 warp hands its 32 values over through shared memory ([column][row]:
// conflict free both ways, two buffers alternate so that one named barrier per pass is enough) and the quarter-0 warp stores
// the finished partial row -- 128 contiguous bytes per lane, no transpose: the epilogue of a 20-row batch is latency, not
// bandwidth.
__device__ __forceinline__ void partial_pass(uint32_t t_main, int cb, int eq, int pair, int& xph, float* xbuf, int lane,
                                             float* row_ptr, int c0, int zp, bool row_ok) {
  float* xb = xbuf + (pair * 2 + (xph & 1)) * 1024;
  ++xph;
  float v[32];
  if (eq == 1) {
    tmem_ld32_raw(t_main + (32u << 16) + cb, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) xb[i * 32 + lane] = v[i];
  } else {
    tmem_ld_groups(t_main, t_main + 128, cb, cb + 8, cb + 16, cb + 24, v);      // hi*hi + hi*lo
  }
  __syncwarp();
  // writer and reader warp meet at the same barrier instruction
  if (pair == 0) asm volatile("bar.sync 1, 64;" ::: "memory"); else asm volatile("bar.sync 2, 64;" ::: "memory");
  if (eq == 0) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += xb[i * 32 + lane];
    if (row_ok) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (c0 + 4 * i < zp)
          *reinterpret_cast<float4*>(row_ptr + c0 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
  }
}
__device__ __forceinline__ void init_common(Bars3* bars, int S, int warp) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bars->accf[b], 1); mbar_init(&bars->acce[b], 4); }
    mbar_init(&bars->wbar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

// ------------------------------------------------------------------------------------------------- forward
struct FwdArgs3 {
  const float* xp;            // [T*B, 4H]  zx Vx^T + bias + x (.) Dx
  const float* Dh; const float* h0; const float* c0;
  float* y; long long ys_t, ys_b;
  float *hT, *cT;
  float* gates;               // [T,B,4,H] or null (inference)
  float* cs;                  // [T,B,H] when saving, else a [2,B,H] ping-pong scratch
  float* z;                   // [T*B, zp] saved z (null in inference)
  float* hop;                 // [CS][hi | lo][32][32] tile-major: CTA s's units in columns 0..7 of its tile
  float* zop;                 // [nkz][hi | lo][32][32] tile-major z_t
  float* zpart;               // [CS, 32, zp]
  unsigned int* sync;         // group barrier counter (one 128-byte line), zeroed before the launch
  int T, B, H, RH;
  int CS, zp, S, G;           // S = activation ring stages of G K tiles each
};

template <bool SAVE>
__global__ void __launch_bounds__(kThreads, 1)
r3_fwd_kernel(const __grid_constant__ CUtensorMap m_hop, const __grid_constant__ CUtensorMap m_zop,
              const __grid_constant__ CUtensorMap m_p_hi, const __grid_constant__ CUtensorMap m_p_lo,
              const __grid_constant__ CUtensorMap m_w2_hi, const __grid_constant__ CUtensorMap m_w2_lo, const FwdArgs3 a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* const base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // the warp index through a broadcast: the compiler then knows the role branches are warp-uniform (uniform registers inside)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int S = a.S, CS = a.CS, G = a.G;
  const int stage_bytes = G * kAStage;
  const int RHr = (a.RH + 7) & ~7;
  const int nzc = (RHr + 127) / 128;                         // phase Z chunks (128 z columns each; <= 4: one k-step slot each)
  const int nkz = (a.RH + BK - 1) / BK;                      // phase G K tiles
  const int ngr = (nkz + G - 1) / G;                         // ... in groups of G per stage
  const int nacc = min(4, (a.RH + 7) / 8);                   // independent accumulation chains of phase G
  // shared memory: activation ring | phase Z weights (hi, lo) | phase G weights (nkz x (hi | lo)) | barriers | hand-over buffers
  uint8_t* const ring = base;
  uint8_t* const p_hi = base + S * stage_bytes;
  uint8_t* const p_lo = p_hi + kPTile;
  uint8_t* const w2s = p_lo + kPTile;
  Bars3* const bars = reinterpret_cast<Bars3*>(w2s + nkz * kAStage);
  float* const xbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kBarBytes);
  const int s_rank = (int)blockIdx.x;
  const int u0 = s_rank * UB;
  unsigned int epoch = 0;

  init_common(bars, S, warp);
  const uint32_t tmem_d = bars->tmem_slot;

  if (warp == 0 && lane == 0) {                              // the CTA's factor rows: loaded once, resident for all T steps
    mbar_arrive_expect_tx(&bars->wbar, 2 * kPTile + nkz * kAStage);
    tma_load_2d(p_hi, &m_p_hi, 0, s_rank * BM, &bars->wbar);
    tma_load_2d(p_lo, &m_p_lo, 0, s_rank * BM, &bars->wbar);
    for (int kt = 0; kt < nkz; ++kt) {
      tma_load_2d(w2s + kt * kAStage, &m_w2_hi, kt * BK, s_rank * 32, &bars->wbar);
      tma_load_2d(w2s + kt * kAStage + kATile, &m_w2_lo, kt * BK, s_rank * 32, &bars->wbar);
    }
  }

  const int eq = warp & 3, ehalf = (warp - 2) >> 2;          // epilogue: tensor-memory lane quarter, column half
  const bool epi = warp >= 2 && eq <= 1;                     // warps 4, 8: accumulator lanes 0..31 (hi rows); 5, 9: lanes 32..63 (lo rows)
  const int rl = lane & 3, c8 = lane >> 2;
  uint32_t n_stage = 0, n_chunk = 0;
  bool wready = false;
  int xph = 0;
  int trc = 0;
  (void)trc;

  for (int t = 0; t < a.T; ++t) {
    // ======================================= phase Z =======================================
    if (warp == 0) {
      R3_TRACE(1);
      const int s = n_stage % S, it = n_stage / S;
      if (it > 0) mbar_wait(&bars->empty[s], (it - 1) & 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->full[s], kAStage);
        tma_load_2d(ring + s * stage_bytes, &m_hop, 0, s_rank * 64, &bars->full[s]);
      }
      ++n_stage;
    } else if (warp == 1) {
      if (!wready) { mbar_wait(&bars->wbar, 0); wready = true; }
      const int s = n_stage % S, it = n_stage / S;
      const uint32_t a_addr = smem_u32(ring + s * stage_bytes), p_addr = smem_u32(p_hi);
      for (int zc = 0; zc < nzc; ++zc, ++n_chunk) {
        const int buf = n_chunk & 1, use = n_chunk >> 1;
        if (use > 0) mbar_wait(&bars->acce[buf], (use - 1) & 1);
        if (zc == 0) { mbar_wait(&bars->full[s], it & 1); R3_TRACE(10); }
        tc_fence_after();
        const int ncol = min(128, RHr - zc * 128);
        const uint32_t idesc = make_idesc(BM, 128 + ((ncol + 15) & ~15));      // [p_hi rows | p_lo rows]
        // chunk zc of the z columns sits in k-step slot zc of the packed tile (seq_r3.cu: pack3_fwd_kernel)
        mma_elect(tmem_d + buf * 256, make_desc(a_addr), make_desc(p_addr + zc * 32), idesc, 0u);
        commit_elect(&bars->accf[buf]);
      }
      commit_elect(&bars->empty[s]);
      R3_TRACE(11);
      ++n_stage;
    } else if (epi) {
      for (int zc = 0; zc < nzc; ++zc, ++n_chunk) {
        const int buf = n_chunk & 1, use = n_chunk >> 1;
        mbar_wait(&bars->accf[buf], use & 1);
        tc_fence_after();
        if (warp == 4) R3_TRACE(20);
        const int ncol = min(128, RHr - zc * 128);
        const uint32_t t_main = tmem_d + buf * 256;
#pragma unroll 1
        for (int pp = 0; pp < 2; ++pp) {
          const int cb = (ehalf * 2 + pp) * 32;
          if (cb >= ncol) break;
          partial_pass(t_main, cb, eq, ehalf, xph, xbuf, lane, a.zpart + ((size_t)s_rank * RB + lane) * a.zp, zc * 128 + cb, a.zp, lane < a.B);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->acce[buf]);
      }
    }
    // ======================================= exchange =======================================
    if (warp == 4) R3_TRACE(21);
    group_sync(a.sync, epoch, CS);
    if (warp == 4) R3_TRACE(30);
    if (warp >= 2) {
      reduce_partials(a.zpart, CS, CS, s_rank, a.B, a.zp, a.zp >> 2, [&](int r, int c, const float4 v) {
        if (SAVE) *reinterpret_cast<float4*>(a.z + ((size_t)t * a.B + r) * a.zp + c) = v;
        store_split4(a.zop, r, c, v);
      }, [&](int ev) { if (warp == 4) R3_TRACE(ev); });
      fence_proxy_async_all();
      if (warp == 4) R3_TRACE(31);
    }
    group_sync(a.sync, epoch, CS);
    if (warp == 4) R3_TRACE(32);
    // ======================================= phase G =======================================
    if (warp == 0) {
      for (int g = 0; g < ngr; ++g, ++n_stage) {
        const int s = n_stage % S, it = n_stage / S;
        if (it > 0) mbar_wait(&bars->empty[s], (it - 1) & 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->full[s], stage_bytes);
          tma_load_2d(ring + s * stage_bytes, &m_zop, 0, g * G * 64, &bars->full[s]);
        }
        R3_TRACE(200 + g);
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc(BM, 64);               // [32 hi rows | 32 lo rows] of the CTA's Bm slice
      const int buf = n_chunk & 1, use = n_chunk >> 1;
      if (use > 0) mbar_wait(&bars->acce[buf], (use - 1) & 1);
      tc_fence_after();
      for (int g = 0; g < ngr; ++g, ++n_stage) {
        const int s = n_stage % S, it = n_stage / S;
        mbar_wait(&bars->full[s], it & 1);
        tc_fence_after();
        R3_TRACE(100 + g);
        const int nt = min(G, nkz - g * G);
        for (int tl = 0; tl < nt; ++tl) {
          const int kt = g * G + tl;
          issue_ktile_split(smem_u32(ring + s * stage_bytes + tl * kAStage), smem_u32(w2s + kt * kAStage), tmem_d + buf * 256, 64, idesc,
                            tile_ksteps(a.RH, kt), kt == 0);
        }
        commit_elect(&bars->empty[s]);
      }
      commit_elect(&bars->accf[buf]);
      R3_TRACE(13);
      ++n_chunk;
    } else if (epi) {
      const int buf = n_chunk & 1, use = n_chunk >> 1;
      if (ehalf == 0) {
        // warp 4 (lane quarter 0): lane (c8, rl) owns unit j for the 8 rows rg*4 + rl; the chunk's 32 columns are [gate k][unit].
        // warp 5 (lane quarter 1) only hands the lo*hi term over.  Both meet at ONE named-barrier instruction.
        const bool q1 = eq == 1;
        float* const xb = xbuf + (xph & 1) * 1024;
        const float* hprev = t ? a.y + (size_t)(t - 1) * a.ys_t : a.h0;
        const long long hp_sb = t ? a.ys_b : a.H;
        const float* cprev = SAVE ? (t ? a.cs + (size_t)(t - 1) * a.B * a.H : a.c0) : (t ? a.cs + (size_t)((t - 1) & 1) * a.B * a.H : a.c0);
        float* cout = SAVE ? a.cs + (size_t)t * a.B * a.H : a.cs + (size_t)(t & 1) * a.B * a.H;
        float* y_t = a.y + (size_t)t * a.ys_t;
        const bool last = (t == a.T - 1);
        const int j = u0 + c8;
        const bool act = !q1 && j < a.H;
        float dh[4], hp[8], cp[8], xq[8][4];
        // operands that do not depend on the accumulator: requested before waiting for it
#pragma unroll
        for (int k = 0; k < 4; ++k) dh[k] = act ? __ldg(a.Dh + k * a.H + j) : 0.f;
#pragma unroll
        for (int rg = 0; rg < 8; ++rg) {
          const int m = rg * 4 + rl;
          const bool ok = act && m < a.B;
          hp[rg] = (ok && hprev) ? hprev[(size_t)m * hp_sb + j] : 0.f;
          cp[rg] = (ok && cprev) ? cprev[(size_t)m * a.H + j] : 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) xq[rg][k] = ok ? __ldg(a.xp + ((size_t)t * a.B + m) * 4 * a.H + (size_t)k * a.H + j) : 0.f;
        }
        if (warp == 4) R3_TRACE(22);
        mbar_wait(&bars->accf[buf], use & 1);
        tc_fence_after();
        if (warp == 4) R3_TRACE(23);
        const uint32_t t_main = tmem_d + buf * 256;
        float v[32];
        if (q1) {
          // lo*hi term (accumulator lanes 32..63) -> shared memory, [column][row]: conflict free for writer and reader
          tmem_ld32_raw(t_main + (32u << 16), v);
          for (int ac = 1; ac < nacc; ++ac) {
            float w[32];
            tmem_ld32_raw(t_main + (32u << 16) + ac * 64, w);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += w[i];
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) xb[i * 32 + lane] = v[i];
        } else {
          tmem_ld_groups(t_main, t_main + 32, 0, 8, 16, 24, v);   // hi*hi + hi*lo of this row, first k-step chain
          for (int ac = 1; ac < nacc; ++ac) {
            float w[32];
            tmem_ld_groups(t_main + ac * 64, t_main + ac * 64 + 32, 0, 8, 16, 24, w);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += w[i];
          }
        }
        __syncwarp();
        asm volatile("bar.sync 1, 64;" ::: "memory");
        if (!q1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += xb[i * 32 + lane]; // + lo*hi
          xpose8(v, lane);                                       // -> v[k*8 + rg] = (row rg*4 + rl, unit c8, gate k)
          if (act) {
            float* hop_t = a.hop + (size_t)s_rank * kTileFloats;
#pragma unroll
            for (int rg = 0; rg < 8; ++rg) {
              const int m = rg * 4 + rl;
              if (m < a.B) {
                float pre[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) pre[k] = v[k * 8 + rg] + xq[rg][k] + hp[rg] * dh[k];
                const float gi = sigmoidf_acc(pre[0]), gf = sigmoidf_acc(pre[1]);
                const float go = sigmoidf_acc(pre[2]), gn = tanhf_acc(pre[3]);
                const float c = fmaf(gf, cp[rg], gi * gn);
                const float h = go * tanhf_acc(c);
                y_t[(size_t)m * a.ys_b + j] = h;
                hop_t[m * 32 + c8] = h;                            // hi part = the value itself (the tensor core truncates)
                hop_t[1024 + m * 32 + c8] = split_lo(h, h);
                cout[(size_t)m * a.H + j] = c;
                if (SAVE) {
                  float* gp = a.gates + ((size_t)t * a.B + m) * 4 * a.H + j;
                  gp[0] = gi; gp[a.H] = gf; gp[2 * a.H] = go; gp[3 * a.H] = gn;
                }
                if (last) { a.hT[(size_t)m * a.H + j] = h; a.cT[(size_t)m * a.H + j] = c; }
              }
            }
          }
        }
      } else {
        mbar_wait(&bars->accf[buf], use & 1);
      }
      if (ehalf == 0) ++xph;                                   // warps 4 and 5 used one hand-over buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acce[buf]);
      ++n_chunk;
      fence_proxy_async_all();
      if (warp == 4) R3_TRACE(24);
    }
    __syncthreads();       // h_t operand columns of this CTA are complete before its next phase Z load
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(512));
  }
}

// ------------------------------------------------------------------------------------------------- backward
struct BwdArgs3 {
  const float* gates;         // [T,B,4,H]
  const float* cs;            // [T,B,H]
  const float* c0;            // [B,H] or null
  const float* dy; long long dys_t, dys_b;      // or null
  const float *dhT, *dcT;     // [B,H] or null
  const float* Dh;
  float *dh0, *dc0;           // [B,H] or null
  float* dpre;                // [T*B, 4, Hp]   exact copy for the time-parallel weight-gradient GEMMs
  float* dz_all;              // [T*B, zp]
  float* dpo;                 // [CS][hi | lo][32][32] tile-major tf32 operand copy of dPre_t: CTA s's (gate, unit) slice
  float* dzo;                 // [nkz][hi | lo][32][32] tile-major tf32 operand copy of dz_t
  float *dhrun, *dcrun;       // [B, Hp]
  float* part;                // [CS, 32, zp]
  unsigned int* sync;
  int T, B, H, RH;
  int Hp, CS, zp, S, G;
};

__device__ __forceinline__ void pw_finish3(const BwdArgs3& a, int tq, int m, int j, const PwIn& in, float dh, const float (&dhc)[4]) {
  const size_t rowq = (size_t)tq * a.B + m;
  const float tcv = tanhf_acc(in.ct);
  const float dc = fmaf(dh * in.go, 1.f - tcv * tcv, in.dcin);
  float d[4];
  d[0] = dc * in.gn * in.gi * (1.f - in.gi);
  d[1] = dc * in.cp * in.gf * (1.f - in.gf);
  d[2] = dh * tcv * in.go * (1.f - in.go);
  d[3] = dc * in.gi * (1.f - in.gn * in.gn);
  float* o = a.dpre + rowq * 4 * a.Hp + j;
  float* op = a.dpo + (size_t)(j >> 3) * kTileFloats + m * 32;
  float sdh = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    o[(size_t)k * a.Hp] = d[k];
    op[slice_col(k, j)] = d[k];                                 // the tensor core truncates: the value is its own hi part
    op[1024 + slice_col(k, j)] = split_lo(d[k], d[k]);
    sdh = fmaf(d[k], dhc[k], sdh);
  }
  a.dcrun[(size_t)m * a.Hp + j] = dc * in.gf;
  a.dhrun[(size_t)m * a.Hp + j] = sdh;
}

__global__ void __launch_bounds__(kThreads, 1)
r3_bwd_kernel(const __grid_constant__ CUtensorMap m_dpo, const __grid_constant__ CUtensorMap m_dzo,
              const __grid_constant__ CUtensorMap m_w2t_hi, const __grid_constant__ CUtensorMap m_w2t_lo,
              const __grid_constant__ CUtensorMap m_ap_hi, const __grid_constant__ CUtensorMap m_ap_lo, const BwdArgs3 a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* const base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // the warp index through a broadcast: the compiler then knows the role branches are warp-uniform (uniform registers inside)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int S = a.S, CS = a.CS, G = a.G;
  const int stage_bytes = G * kAStage;
  const int RHr = (a.RH + 7) & ~7;
  const int nch1 = (RHr + 127) / 128;                        // phase 1 chunks (128 dz columns)
  const int nkz = (a.RH + BK - 1) / BK;                      // phase 2 K tiles
  const int ngr = (nkz + G - 1) / G;
  const int nacc = min(4, (a.RH + 7) / 8);                   // independent accumulation chains of phase 2
  // shared memory: activation ring | phase 1 weights (nch1 x (hi | lo) 16 KB tiles) | phase 2 weights (nkz x (hi | lo) 2 KB) | barriers | hand-over
  uint8_t* const ring = base;
  uint8_t* const w2ts = base + S * stage_bytes;
  uint8_t* const aps = w2ts + nch1 * 2 * kPTile;
  Bars3* const bars = reinterpret_cast<Bars3*>(aps + nkz * 2 * kApTile);
  float* const xbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kBarBytes);
  const int s_rank = (int)blockIdx.x;
  const int u0 = s_rank * UB;
  unsigned int epoch = 0;

  init_common(bars, S, warp);
  const uint32_t tmem_d = bars->tmem_slot;

  if (warp == 0 && lane == 0) {
    mbar_arrive_expect_tx(&bars->wbar, nch1 * 2 * kPTile + nkz * 2 * kApTile);
    for (int c = 0; c < nch1; ++c) {
      tma_load_2d(w2ts + c * 2 * kPTile, &m_w2t_hi, s_rank * 32, c * BM, &bars->wbar);
      tma_load_2d(w2ts + c * 2 * kPTile + kPTile, &m_w2t_lo, s_rank * 32, c * BM, &bars->wbar);
    }
    for (int kt = 0; kt < nkz; ++kt) {
      tma_load_2d(aps + kt * 2 * kApTile, &m_ap_hi, kt * BK, u0, &bars->wbar);
      tma_load_2d(aps + kt * 2 * kApTile + kApTile, &m_ap_lo, kt * BK, u0, &bars->wbar);
    }
  }

  const int eq = warp & 3, ehalf = (warp - 2) >> 2;
  const bool epi = warp >= 2 && eq <= 1;
  const int rl = lane & 3, c8 = lane >> 2;
  const int j = u0 + c8;                                     // this lane's hidden unit in the gate-gradient algebra
  const bool act = j < a.H;
  float dhc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) dhc[k] = act ? __ldg(a.Dh + k * a.H + j) : 0.f;
  uint32_t n_stage = 0, n_chunk = 0;
  bool wready = false;
  int xph = 0;
  int trc = 0;
  (void)trc;

  // ---- seed: gate-gradient algebra of the last step with dh = dhT (+ dy), dc = dcT ----
  if (epi && ehalf == 0 && eq == 0 && act) {
    PwIn in[8];
#pragma unroll
    for (int rg = 0; rg < 8; ++rg)
      if (rg * 4 + rl < a.B) pw_load(a, a.T - 1, rg * 4 + rl, j, true, in[rg]);
#pragma unroll
    for (int rg = 0; rg < 8; ++rg)
      if (rg * 4 + rl < a.B) pw_finish3(a, a.T - 1, rg * 4 + rl, j, in[rg], in[rg].dyv + in[rg].dhs, dhc);
  }
  fence_proxy_async_all();
  __syncthreads();

  for (int t = a.T - 1; t >= 0; --t) {
    // ======================================= phase 1 =======================================
    if (warp == 0) {
      R3_TRACE(1);
      const int s = n_stage % S, it = n_stage / S;
      if (it > 0) mbar_wait(&bars->empty[s], (it - 1) & 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->full[s], kAStage);
        tma_load_2d(ring + s * stage_bytes, &m_dpo, 0, s_rank * 64, &bars->full[s]);
      }
      ++n_stage;
    } else if (warp == 1) {
      if (!wready) { mbar_wait(&bars->wbar, 0); wready = true; }
      const int s = n_stage % S, it = n_stage / S;
      const uint32_t a_addr = smem_u32(ring + s * stage_bytes);
      for (int c = 0; c < nch1; ++c, ++n_chunk) {
        const int buf = n_chunk & 1, use = n_chunk >> 1;
        if (use > 0) mbar_wait(&bars->acce[buf], (use - 1) & 1);
        if (c == 0) { mbar_wait(&bars->full[s], it & 1); R3_TRACE(10); }
        tc_fence_after();
        const int ncol = min(128, RHr - c * 128);
        const uint32_t idesc = make_idesc(BM, 128 + ((ncol + 15) & ~15));      // [w2t hi rows | w2t lo rows]
        issue_ktile(a_addr, smem_u32(w2ts + c * 2 * kPTile), tmem_d + buf * 256, idesc, 4, true);
        commit_elect(&bars->accf[buf]);
      }
      commit_elect(&bars->empty[s]);
      R3_TRACE(11);
      ++n_stage;
    } else if (epi) {
      for (int c = 0; c < nch1; ++c, ++n_chunk) {
        const int buf = n_chunk & 1, use = n_chunk >> 1;
        mbar_wait(&bars->accf[buf], use & 1);
        tc_fence_after();
        if (warp == 4) R3_TRACE(20);
        const int ncol = min(128, RHr - c * 128);
        const uint32_t t_main = tmem_d + buf * 256;
#pragma unroll 1
        for (int pp = 0; pp < 2; ++pp) {
          const int cb = (ehalf * 2 + pp) * 32;
          if (cb >= ncol) break;
          partial_pass(t_main, cb, eq, ehalf, xph, xbuf, lane, a.part + ((size_t)s_rank * RB + lane) * a.zp, c * 128 + cb, a.zp, lane < a.B);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->acce[buf]);
      }
    }
    // ======================================= exchange =======================================
    if (warp == 4) R3_TRACE(21);
    group_sync(a.sync, epoch, CS);
    if (warp == 4) R3_TRACE(30);
    if (warp >= 2) {
      reduce_partials(a.part, CS, CS, s_rank, a.B, a.zp, a.zp >> 2, [&](int r, int c, const float4 v) {
        *reinterpret_cast<float4*>(a.dz_all + ((size_t)t * a.B + r) * a.zp + c) = v;
        store_split4(a.dzo, r, c, v);
      }, [&](int ev) { if (warp == 4) R3_TRACE(ev); });
      fence_proxy_async_all();
      if (warp == 4) R3_TRACE(31);
    }
    group_sync(a.sync, epoch, CS);
    if (warp == 4) R3_TRACE(32);
    // ======================================= phase 2 =======================================
    if (warp == 0) {
      for (int g = 0; g < ngr; ++g, ++n_stage) {
        const int s = n_stage % S, it = n_stage / S;
        if (it > 0) mbar_wait(&bars->empty[s], (it - 1) & 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->full[s], stage_bytes);
          tma_load_2d(ring + s * stage_bytes, &m_dzo, 0, g * G * 64, &bars->full[s]);
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc(BM, 32);               // [16 hi rows | 16 lo rows] of the CTA's A slice
      const int buf = n_chunk & 1, use = n_chunk >> 1;
      if (use > 0) mbar_wait(&bars->acce[buf], (use - 1) & 1);
      tc_fence_after();
      for (int g = 0; g < ngr; ++g, ++n_stage) {
        const int s = n_stage % S, it = n_stage / S;
        mbar_wait(&bars->full[s], it & 1);
        tc_fence_after();
        R3_TRACE(100 + g);
        const int nt = min(G, nkz - g * G);
        for (int tl = 0; tl < nt; ++tl) {
          const int kt = g * G + tl;
          issue_ktile_split(smem_u32(ring + s * stage_bytes + tl * kAStage), smem_u32(aps + kt * 2 * kApTile), tmem_d + buf * 256, 32, idesc,
                            tile_ksteps(a.RH, kt), kt == 0);
        }
        commit_elect(&bars->empty[s]);
      }
      commit_elect(&bars->accf[buf]);
      R3_TRACE(13);
      ++n_chunk;
    } else if (epi) {
      const int buf = n_chunk & 1, use = n_chunk >> 1;
      if (ehalf == 0) {
        // warp 4: gate-gradient algebra of step t-1; warp 5 hands the lo*hi term over; one named-barrier instruction for both
        const bool q1 = eq == 1;
        float* const xb = xbuf + (xph & 1) * 1024;
        // the saved activations of step t-1 are requested before the accumulator is waited for
        PwIn in[8];
        if (!q1 && t > 0 && act) {
#pragma unroll
          for (int rg = 0; rg < 8; ++rg)
            if (rg * 4 + rl < a.B) pw_load(a, t - 1, rg * 4 + rl, j, false, in[rg]);
        }
        if (warp == 4) R3_TRACE(22);
        mbar_wait(&bars->accf[buf], use & 1);
        tc_fence_after();
        if (warp == 4) R3_TRACE(23);
        const uint32_t t_main = tmem_d + buf * 256;
        float v[8];
        if (q1) {
          tc::tmem_ld8_raw(t_main + (32u << 16), v);             // lo*hi term of the 8 units
          tc::tmem_ld_wait();
          for (int ac = 1; ac < nacc; ++ac) {
            float w[8];
            tc::tmem_ld8_raw(t_main + (32u << 16) + ac * 32, w);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += w[i];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) xb[i * 32 + lane] = v[i];
        } else {
          tmem_ld_group(t_main, t_main + 16, 0, v);             // hi*hi + hi*lo, first k-step chain
          for (int ac = 1; ac < nacc; ++ac) {
            float w[8];
            tmem_ld_group(t_main + ac * 32, t_main + ac * 32 + 16, 0, w);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += w[i];
          }
        }
        __syncwarp();
        asm volatile("bar.sync 1, 64;" ::: "memory");
        if (!q1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] += xb[i * 32 + lane];  // + lo*hi
          xpose8_group(v, lane);                                  // -> v[rg] = dh_{t-1}(row rg*4 + rl, unit c8) without the Dh term
          if (act) {
#pragma unroll
            for (int rg = 0; rg < 8; ++rg) {
              const int m = rg * 4 + rl;
              if (m < a.B) {
                if (t > 0) {
                  pw_finish3(a, t - 1, m, j, in[rg], in[rg].dyv + in[rg].dhs + v[rg], dhc);
                } else {
                  if (a.dh0) a.dh0[(size_t)m * a.H + j] = a.dhrun[(size_t)m * a.Hp + j] + v[rg];
                  if (a.dc0) a.dc0[(size_t)m * a.H + j] = a.dcrun[(size_t)m * a.Hp + j];
                }
              }
            }
          }
        }
      } else {
        mbar_wait(&bars->accf[buf], use & 1);
      }
      if (ehalf == 0) ++xph;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acce[buf]);
      ++n_chunk;
      fence_proxy_async_all();
      if (warp == 4) R3_TRACE(24);
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(512));
  }
}

}  // namespace r3
}  // namespace vmlmf
