// seq_r2.cuh -- regime R2: the PERSISTENT large-H / high-rank recurrence on the 5th-generation tensor cores.
//
// Replaces the launch-per-timestep generic regime for every shape beyond the register-resident kernels (LM layer
// H=650 r=300, V/models/vmlmf_lm.py:272-280 with lstm_step :222-269; ranks > 16; the H >= 1024 sweep,
// V/models/vmlmf.py:308-310 with the cell body :78-125).  One launch runs all T timesteps.
//
// Decomposition.  A GROUP of CS co-resident CTAs (cooperative launch, one CTA per SM) owns one 128-row batch tile for
// all T steps; CTA s of the group owns the hidden units [s*HS, (s+1)*HS) (HS % 32 == 0).  Groups walk the batch tiles
// round-robin.  CS is chosen so that groups x CS covers the SMs: up to H/32 CTAs per tile (a thread-block cluster would
// cap it at 8; the exchange goes through L2 either way, so the group barrier is a release/acquire counter in global memory).
// Per timestep, per CTA (fp32-accurate 3xTF32 products, accumulators in tensor memory):
//   phase Z   partial z_s[128, RH]   = h_{t-1}[:, slice s] * A[slice s, :]            K = HS   (K-split over the group)
//   exchange  z = sum_s z_s  (partials through L2, fixed-order reduce, rows split over the group's CTAs)
//   phase G   pre[128, 4 x 32 units] = [z | zx_t] * [Bm | Vx]^T  per 32-unit chunk      K = RH + RX  (x side fused:
//             XP[T*B,4H] is never materialised), epilogue = + x (.) Dx + h_{t-1} (.) Dh + bias, gates, c/h update,
//             saved activations, and h_t written back as the next step's tensor-core operand (tf32 hi / lo parts).
// All tensor-core operands are TMA tiles (128-byte swizzle) of L2-resident buffers: the factor matrices are packed
// once per call (K-major, zero padded, pre-split into tf32 hi and lo parts), the activations h / z are rewritten by
// the epilogues every step ([B, Hp] / [B, zp] scratch, a few hundred KB per tile: they never leave L2).
// Warp roles: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+ TMEM allocation), warps 2-9 = epilogue (two per
// tensor-memory lane quarter; accumulator rows reach row-major global memory through a register transpose, xpose8).
// Tensor memory: 2 accumulator buffers x (main | cross-term) x 128 columns = all 512 columns, so the MMAs of chunk
// c+1 overlap the epilogue of chunk c.  hi*hi products go to `main`, the two cross products to `cross` (the tensor
// core's accumulate truncates: keeping the small terms apart keeps the number of roundings of the large sum at K/8).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace vmlmf {
namespace r2 {

using tc::elect_one;
using tc::fence_barrier_init;
using tc::make_desc;
using tc::make_idesc;
using tc::mbar_arrive;
using tc::mbar_arrive_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::mma_commit;
using tc::mma_tf32_ss;
using tc::smem_u32;
using tc::tc_fence_after;
using tc::tc_fence_before;
using tc::tma_load_2d;
using tc::tma_load_3d;
using tc::warp_index;

constexpr int BM = 128;                      // batch rows per tile (= TMEM lanes)
constexpr int BK = 32;                       // fp32 per K tile (128 bytes = one swizzle row)
constexpr int kStages = 3;
constexpr int kTile = BM * BK * 4;           // 16 KB: one [128 x 32] fp32 operand tile
constexpr int kStageBytes = 4 * kTile;       // A hi | A lo | B hi | B lo
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment*/ + 256 /*barriers*/;
constexpr int kXposeBytes = 32 * 32 * 4;     // backward: per-epilogue-warp [32 rows][32 units] transpose tile (XOR swizzled)
constexpr int kEpiWarps = 8;                 // two per tensor-memory lane quarter
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kSmemBytesBwd = kSmemBytes + kEpiWarps * kXposeBytes;
constexpr int kMaxGroup = 64;                // CTAs per batch tile

// ------------------------------------------------------------------------------------------------- PTX
// Barrier among the CS CTAs of one tile group (all co-resident: cooperative launch).  `ctr` is the group's monotonic
// arrival counter (zeroed before the launch); the k-th barrier completes when it reaches (k+1) * CS.  Release / acquire at
// GPU scope: everything the group's threads wrote before the barrier is visible to all of them after it.
__device__ __forceinline__ void group_sync(unsigned int* ctr, unsigned int& epoch, int CS) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const unsigned int target = epoch * (unsigned int)CS;
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}
// generic-proxy writes (st.global / st.shared) -> visible to the async proxy (TMA reads) once a barrier follows
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// tf32 hi / lo parts of an fp32 operand.  The tensor core reads the top 19 bits of a 32-bit tf32 container and ignores the low
// 13 mantissa bits, so the raw value is its own (truncated) hi part; lo = v - trunc(v) is exact in fp32 and < 2^-10 |v|
// (measured: same end-to-end error as round-to-nearest splitting, tests/test_gpu_parity.py; -DVMLMF_SPLIT_ROUND restores it).
#ifdef VMLMF_SPLIT_ROUND
__device__ __forceinline__ float split_hi(float v) { return tf32_rna(v); }
__device__ __forceinline__ float split_lo(float v, float hi) { return tf32_rna(v - hi); }
#else
__device__ __forceinline__ float split_hi(float v) { return v; }
__device__ __forceinline__ float split_lo(float v, float) { return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
#endif

// Debug-only cycle trace (-DVMLMF_R2_TRACE, tools/trace_r2.py): lane 0 of each role of CTA 0 appends (event, clock64) pairs
// for a few timesteps.  Compiled out of the shipped library.
#ifdef VMLMF_R2_TRACE
static __device__ long long g_r2_trace[8192];
static __device__ int g_r2_trace_n;
#define R2_TRACE(ev)                                                                      \
  do {                                                                                    \
    if (blockIdx.x == 0 && lane == 0 && t >= 8 && t < 11) {                               \
      const int i__ = atomicAdd(&g_r2_trace_n, 1);                                        \
      if (i__ < 4096) { g_r2_trace[2 * i__] = (ev) + 1000 * t; g_r2_trace[2 * i__ + 1] = clock64(); } \
    }                                                                                     \
  } while (0)
#else
#define R2_TRACE(ev) do {} while (0)
#endif

// ------------------------------------------------------------------------------------------------- shared pieces
struct Bars {
  uint64_t full[kStages];      // TMA bytes of the stage landed
  uint64_t empty[kStages];     // MMAs that read the stage completed (tcgen05.commit)
  uint64_t accf[2];            // accumulator buffer complete
  uint64_t acce[2];            // accumulator buffer drained by the four epilogue warps
  uint32_t tmem_slot;
};

struct Smem {
  uint8_t* stages;
  Bars* bars;
};
__device__ __forceinline__ Smem carve(uint8_t* raw) {
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  Smem s;
  s.stages = base;
  s.bars = reinterpret_cast<Bars*>(base + kStages * kStageBytes);
  return s;
}

// number of 8-wide tf32 k-steps of K tile `kt` when only `valid` of the K columns are non-zero
__device__ __forceinline__ int tile_ksteps(int valid, int kt) {
  const int left = valid - kt * BK;
  return left >= BK ? 4 : (left + 7) / 8;
}

// the 3 x ksteps MMAs of one K tile: cross terms first (small), then hi*hi.  Called by all lanes of the MMA warp (warp-uniform
// operands live in uniform registers); one elected lane issues.  A k-step advances the descriptors' start address by 32 bytes.
__device__ __forceinline__ void issue_tile(uint32_t stage_addr, uint32_t acc_main, uint32_t acc_cross, uint32_t idesc,
                                           int ksteps, bool first) {
  const uint64_t a_hi = make_desc(stage_addr), a_lo = make_desc(stage_addr + kTile);
  const uint64_t b_hi = make_desc(stage_addr + 2 * kTile), b_lo = make_desc(stage_addr + 3 * kTile);
  if (elect_one()) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < ksteps) {
        const uint32_t acc = (first && k == 0) ? 0u : 1u;
        mma_tf32_ss(acc_cross, a_lo + 2 * k, b_hi + 2 * k, idesc, acc);
        mma_tf32_ss(acc_cross, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
        mma_tf32_ss(acc_main, a_hi + 2 * k, b_hi + 2 * k, idesc, acc);
      }
    }
  }
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
  if (elect_one()) mma_commit(bar);
}

// ---- epilogue data movement without shared memory ----
// tcgen05.ld.32x32b hands thread `lane` of a warp the accumulator ROW (batch row) lane; global tensors are row-major
// with the hidden unit / z column contiguous.  A three-stage butterfly exchanges the low three bits of the register
// index (8 consecutive columns) with lane bits 4..2: afterwards lane (c8 = lane >> 2, rl = lane & 3) holds, in register
// (g, rg), the element (row rg*4 + rl, column g*8 + c8) -- one warp instruction then touches 4 rows x 8 consecutive
// columns = four fully used 32-byte sectors.  48 shuffles per 32 values; no staging buffer, no barrier between warps.
__device__ __forceinline__ void xpose8(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const int c = 1 << s, lb = 4 << s;
    const bool up = (lane & lb) != 0;
#pragma unroll
    for (int r0 = 0; r0 < 32; ++r0) {
      if (r0 & c) continue;
      const int r1 = r0 | c;
      const float send = up ? v[r0] : v[r1];
      const float recv = __shfl_xor_sync(0xffffffffu, send, lb);
      if (up) v[r0] = recv; else v[r1] = recv;
    }
  }
}
// Variant for outputs whose 32 columns are contiguous in memory (z, dz, partials): exchange register bits 4..2 (column / 4)
// with lane bits 4..2.  Afterwards lane (c4 = lane >> 2, rl = lane & 3) holds v[rg*4 + i] = (row rg*4 + rl, column c4*4 + i):
// one float4 per row, and a warp instruction writes 4 rows x 128 contiguous bytes.
__device__ __forceinline__ void xpose_vec4(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const int c = 4 << s, lb = 4 << s;
    const bool up = (lane & lb) != 0;
#pragma unroll
    for (int r0 = 0; r0 < 32; ++r0) {
      if (r0 & c) continue;
      const int r1 = r0 | c;
      const float send = up ? v[r0] : v[r1];
      const float recv = __shfl_xor_sync(0xffffffffu, send, lb);
      if (up) v[r0] = recv; else v[r1] = recv;
    }
  }
}
// the same exchange on one 8-column group: v[i] (row lane, column i) -> lane (c8, rl) holds v[rg] = (row rg*4 + rl, column c8)
__device__ __forceinline__ void xpose8_group(float (&v)[8], int lane) {
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const int c = 1 << s, lb = 4 << s;
    const bool up = (lane & lb) != 0;
#pragma unroll
    for (int r0 = 0; r0 < 8; ++r0) {
      if (r0 & c) continue;
      const int r1 = r0 | c;
      const float send = up ? v[r0] : v[r1];
      const float recv = __shfl_xor_sync(0xffffffffu, send, lb);
      if (up) v[r0] = recv; else v[r1] = recv;
    }
  }
}
__device__ __forceinline__ void tmem_ld_group(uint32_t t_main, uint32_t t_cross, int c0, float (&v)[8]) {
  float b[8];
  tc::tmem_ld8_raw(t_main + c0, v);
  tc::tmem_ld8_raw(t_cross + c0, b);
  tc::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] += b[i];
}
// v[g*8 + i] = main + cross of this thread's row at columns col(g) + i, g = 0..3 (four 8-column groups)
__device__ __forceinline__ void tmem_ld_groups(uint32_t t_main, uint32_t t_cross, int c0, int c1, int c2, int c3, float (&v)[32]) {
  float a[4][8], b[4][8];
  tc::tmem_ld8_raw(t_main + c0, a[0]); tc::tmem_ld8_raw(t_main + c1, a[1]);
  tc::tmem_ld8_raw(t_main + c2, a[2]); tc::tmem_ld8_raw(t_main + c3, a[3]);
  tc::tmem_ld8_raw(t_cross + c0, b[0]); tc::tmem_ld8_raw(t_cross + c1, b[1]);
  tc::tmem_ld8_raw(t_cross + c2, b[2]); tc::tmem_ld8_raw(t_cross + c3, b[3]);
  tc::tmem_ld_wait();
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int i = 0; i < 8; ++i) v[g * 8 + i] = a[g][i] + b[g][i];
}

// ------------------------------------------------------------------------------------------------- forward
struct FwdArgs {
  // canonical vectors / inputs
  const float* x; long long xs_t, xs_b;
  const float* Dx; const float* Dh; const float* bias;
  const float* h0; const float* c0;
  // outputs
  float* y; long long ys_t, ys_b;
  float *hT, *cT;
  float* gates;               // [T,B,4,H] or null (inference)
  float* cs;                  // [T,B,H] when saving, else a [2,B,H] ping-pong scratch
  float* z;                   // [T*B, zp] saved z (null in inference)
  // operand scratch
  float *hop_hi, *hop_lo;     // [B, Hp]
  float *zop_hi, *zop_lo;     // [B, zp]
  float* zpart;               // [groups, CS, 128, zp] partial z (CS > 1)
  unsigned int* sync;         // [groups, 32] group barrier counters, zeroed before the launch
  int T, B, I, H, RX, RH;
  int Hp, HS, CS, zp, KZP;    // KZP = round_up(RH, 32): K offset of the x side inside the packed gate factor
  int save;
};

template <bool SAVE>
__global__ void __launch_bounds__(kThreads, 1)
r2_fwd_kernel(const __grid_constant__ CUtensorMap m_hop_hi, const __grid_constant__ CUtensorMap m_hop_lo,
              const __grid_constant__ CUtensorMap m_at_hi, const __grid_constant__ CUtensorMap m_at_lo,
              const __grid_constant__ CUtensorMap m_zop_hi, const __grid_constant__ CUtensorMap m_zop_lo,
              const __grid_constant__ CUtensorMap m_zx_hi, const __grid_constant__ CUtensorMap m_zx_lo,
              const __grid_constant__ CUtensorMap m_w2_hi, const __grid_constant__ CUtensorMap m_w2_lo, const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const Smem sm = carve(smem_raw);
  Bars* bars = sm.bars;
  const int warp = warp_index(), lane = threadIdx.x & 31;      // broadcast: role branches are warp-uniform for the compiler
  const int CS = a.CS;
  const int s_rank = (int)blockIdx.x % CS, cid = (int)blockIdx.x / CS, ncl = (int)gridDim.x / CS;
  const int ntiles = (a.B + BM - 1) / BM;
  unsigned int* const sync_ctr = a.sync + cid * 32;          // one 128-byte line per group
  unsigned int epoch = 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bars->accf[b], 1); mbar_init(&bars->acce[b], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bars->tmem_slot;

  // this CTA's slice of the hidden units and the shapes of its two phases
  const int u0 = s_rank * a.HS;
  const int uvalid = min(a.HS, a.H - u0);                    // > 0 by construction of (CS, HS)
  const int RHr = (a.RH + 7) & ~7;                           // z columns the tensor core produces
  const int nzc = (RHr + 127) / 128;                         // phase Z chunks (128 z columns each)
  const int nkh = (uvalid + BK - 1) / BK;                    // phase Z K tiles
  const int ngc = (uvalid + 31) / 32;                        // phase G chunks (32 units x 4 gates)
  const int nkz = (a.RH + BK - 1) / BK, nkx = (a.RX + BK - 1) / BK;

  // epilogue warps: tensor-memory lane quarter q (rows 32q..32q+31 of the tile), and which half of a chunk's columns
  const int eq = warp & 3, ehalf = (warp - 2) >> 2;
  const int rl = lane & 3, c8 = lane >> 2;

  uint32_t n_tile = 0;      // K tiles produced / consumed so far (each role keeps its own copy; all advance alike)
  uint32_t n_chunk = 0;     // accumulator chunks so far

  for (int tile = cid; tile < ntiles; tile += ncl) {
    const int row0 = tile * BM;
    for (int t = 0; t < a.T; ++t) {
      // ======================================= phase Z =======================================
      if (warp == 0) {
        fence_proxy_async_all();
        R2_TRACE(1);
        for (int zc = 0; zc < nzc; ++zc)
          for (int kt = 0; kt < nkh; ++kt, ++n_tile) {
            const int s = n_tile % kStages, it = n_tile / kStages;
            if (it > 0) mbar_wait(&bars->empty[s], (it - 1) & 1);
            uint8_t* st = sm.stages + s * kStageBytes;
            if (elect_one()) {
              mbar_arrive_expect_tx(&bars->full[s], kStageBytes);
              tma_load_2d(st, &m_hop_hi, u0 + kt * BK, row0, &bars->full[s]);
              tma_load_2d(st + kTile, &m_hop_lo, u0 + kt * BK, row0, &bars->full[s]);
              tma_load_2d(st + 2 * kTile, &m_at_hi, u0 + kt * BK, zc * 128, &bars->full[s]);
              tma_load_2d(st + 3 * kTile, &m_at_lo, u0 + kt * BK, zc * 128, &bars->full[s]);
            }
          }
        R2_TRACE(2);
      } else if (warp == 1) {
        for (int zc = 0; zc < nzc; ++zc, ++n_chunk) {
          const int buf = n_chunk & 1, use = n_chunk >> 1;
          if (use > 0) mbar_wait(&bars->acce[buf], (use - 1) & 1);
          tc_fence_after();
          const int ncol = min(128, RHr - zc * 128);
          const uint32_t idesc = make_idesc(BM, (ncol + 15) & ~15);
          const uint32_t acc_main = tmem_d + buf * 256, acc_cross = acc_main + 128;
          for (int kt = 0; kt < nkh; ++kt, ++n_tile) {
            const int s = n_tile % kStages, it = n_tile / kStages;
            mbar_wait(&bars->full[s], it & 1);
            tc_fence_after();
            if (kt == 0) R2_TRACE(10);
            issue_tile(smem_u32(sm.stages + s * kStageBytes), acc_main, acc_cross, idesc, tile_ksteps(uvalid, kt), kt == 0);
            commit_elect(&bars->empty[s]);
          }
          commit_elect(&bars->accf[buf]);
          R2_TRACE(11);
        }
      } else {
        for (int zc = 0; zc < nzc; ++zc, ++n_chunk) {
          const int buf = n_chunk & 1, use = n_chunk >> 1;
          mbar_wait(&bars->accf[buf], use & 1);
          tc_fence_after();
          if (warp == 2) R2_TRACE(20);
          const int ncol = min(128, RHr - zc * 128);
          const uint32_t t_main = tmem_d + ((uint32_t)(eq * 32) << 16) + buf * 256;
          // this warp's two 32-column passes of the chunk: partial z (CS > 1) or the final z (CS == 1); zp % 4 == 0, so a
          // lane's four columns are all inside or all outside the row
#pragma unroll 1
          for (int pp = 0; pp < 2; ++pp) {
            const int cb = (ehalf * 2 + pp) * 32;
            if (cb >= ncol) break;
            float v[32];
            tmem_ld_groups(t_main, t_main + 128, cb, cb + 8, cb + 16, cb + 24, v);
            xpose_vec4(v, lane);
            const int c = zc * 128 + cb + (lane >> 2) * 4;
            if (c < a.zp) {
#pragma unroll
              for (int rg = 0; rg < 8; ++rg) {
                const int r = eq * 32 + rg * 4 + rl, m = row0 + r;
                if (m < a.B) {
                  const float4 val = make_float4(v[rg * 4], v[rg * 4 + 1], v[rg * 4 + 2], v[rg * 4 + 3]);
                  if (CS > 1) {
                    *reinterpret_cast<float4*>(a.zpart + (((size_t)cid * CS + s_rank) * BM + r) * a.zp + c) = val;
                  } else {
                    if (SAVE) *reinterpret_cast<float4*>(a.z + ((size_t)t * a.B + m) * a.zp + c) = val;
                    float4 hi, lo;
                    hi.x = split_hi(val.x); hi.y = split_hi(val.y); hi.z = split_hi(val.z); hi.w = split_hi(val.w);
                    lo.x = split_lo(val.x, hi.x); lo.y = split_lo(val.y, hi.y); lo.z = split_lo(val.z, hi.z); lo.w = split_lo(val.w, hi.w);
                    *reinterpret_cast<float4*>(a.zop_hi + (size_t)m * a.zp + c) = hi;
                    *reinterpret_cast<float4*>(a.zop_lo + (size_t)m * a.zp + c) = lo;
                  }
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->acce[buf]);
        }
        fence_proxy_async_all();
        if (warp == 2) R2_TRACE(21);
      }
      // ======================================= exchange =======================================
      if (CS > 1) {
        group_sync(sync_ctr, epoch, CS);
        if (warp == 2) R2_TRACE(30);
        if (warp >= 2) {
          // fixed-order sum of the CS partials; the tile's valid rows are split over the group's CTAs
          const int rows_valid = min(BM, a.B - row0);
          const int rpc = (rows_valid + CS - 1) / CS;
          const int r_lo = s_rank * rpc, r_hi = min(rows_valid, r_lo + rpc);
          const int et = threadIdx.x - 64;                                   // 0 .. 32 * kEpiWarps - 1
          const int zp4 = a.zp >> 2;
          const int nel = (r_hi - r_lo) * zp4;                               // float4 elements of this CTA's rows
          const float4* pbase = reinterpret_cast<const float4*>(a.zpart + ((size_t)cid * CS * BM + r_lo) * a.zp);
          const size_t pstride = (size_t)BM * zp4;
          for (int e = et; e < nel; e += 32 * kEpiWarps) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int q0 = 0; q0 < CS; q0 += 8) {                              // eight 16-byte loads in flight; fixed summation order
              float4 pv[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) pv[q] = q0 + q < CS ? __ldcg(pbase + (size_t)(q0 + q) * pstride + e) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int q = 0; q < 8; ++q) { v.x += pv[q].x; v.y += pv[q].y; v.z += pv[q].z; v.w += pv[q].w; }   // absent ranks add +0
            }
            const int r = r_lo + e / zp4, c = (e - (e / zp4) * zp4) * 4;
            const int m = row0 + r;
            if (SAVE) *reinterpret_cast<float4*>(a.z + ((size_t)t * a.B + m) * a.zp + c) = v;
            float4 hi, lo;
            hi.x = split_hi(v.x); hi.y = split_hi(v.y); hi.z = split_hi(v.z); hi.w = split_hi(v.w);
            lo.x = split_lo(v.x, hi.x); lo.y = split_lo(v.y, hi.y); lo.z = split_lo(v.z, hi.z); lo.w = split_lo(v.w, hi.w);
            *reinterpret_cast<float4*>(a.zop_hi + (size_t)m * a.zp + c) = hi;
            *reinterpret_cast<float4*>(a.zop_lo + (size_t)m * a.zp + c) = lo;
          }
          fence_proxy_async_all();
          if (warp == 2) R2_TRACE(31);
        }
        group_sync(sync_ctr, epoch, CS);
        if (warp == 2) R2_TRACE(32);
      } else {
        __syncthreads();
      }
      // ======================================= phase G =======================================
      if (warp == 0) {
        fence_proxy_async_all();
        R2_TRACE(3);
        for (int gc = 0; gc < ngc; ++gc)
          for (int kt = 0; kt < nkz + nkx; ++kt, ++n_tile) {
            const int s = n_tile % kStages, it = n_tile / kStages;
            if (it > 0) mbar_wait(&bars->empty[s], (it - 1) & 1);
            uint8_t* st = sm.stages + s * kStageBytes;
            if (elect_one()) {
              mbar_arrive_expect_tx(&bars->full[s], kStageBytes);
              if (kt < nkz) {
                tma_load_2d(st, &m_zop_hi, kt * BK, row0, &bars->full[s]);
                tma_load_2d(st + kTile, &m_zop_lo, kt * BK, row0, &bars->full[s]);
                tma_load_3d(st + 2 * kTile, &m_w2_hi, kt * BK, u0 + gc * 32, 0, &bars->full[s]);
                tma_load_3d(st + 3 * kTile, &m_w2_lo, kt * BK, u0 + gc * 32, 0, &bars->full[s]);
              } else {
                const int kx = kt - nkz;
                tma_load_3d(st, &m_zx_hi, kx * BK, row0, t, &bars->full[s]);
                tma_load_3d(st + kTile, &m_zx_lo, kx * BK, row0, t, &bars->full[s]);
                tma_load_3d(st + 2 * kTile, &m_w2_hi, a.KZP + kx * BK, u0 + gc * 32, 0, &bars->full[s]);
                tma_load_3d(st + 3 * kTile, &m_w2_lo, a.KZP + kx * BK, u0 + gc * 32, 0, &bars->full[s]);
              }
            }
          }
        R2_TRACE(4);
      } else if (warp == 1) {
        const uint32_t idesc = make_idesc(BM, 128);
        for (int gc = 0; gc < ngc; ++gc, ++n_chunk) {
          const int buf = n_chunk & 1, use = n_chunk >> 1;
          if (use > 0) mbar_wait(&bars->acce[buf], (use - 1) & 1);
          tc_fence_after();
          const uint32_t acc_main = tmem_d + buf * 256, acc_cross = acc_main + 128;
          for (int kt = 0; kt < nkz + nkx; ++kt, ++n_tile) {
            const int s = n_tile % kStages, it = n_tile / kStages;
            mbar_wait(&bars->full[s], it & 1);
            tc_fence_after();
            const int ks = kt < nkz ? tile_ksteps(a.RH, kt) : tile_ksteps(a.RX, kt - nkz);
            if (kt == 0) R2_TRACE(12);
            issue_tile(smem_u32(sm.stages + s * kStageBytes), acc_main, acc_cross, idesc, ks, kt == 0);
            commit_elect(&bars->empty[s]);
          }
          commit_elect(&bars->accf[buf]);
          R2_TRACE(13);
        }
      } else {
        const float* hprev = t ? a.y + (size_t)(t - 1) * a.ys_t : a.h0;
        const long long hp_sb = t ? a.ys_b : a.H;
        const float* cprev = SAVE ? (t ? a.cs + (size_t)(t - 1) * a.B * a.H : a.c0) : (t ? a.cs + (size_t)((t - 1) & 1) * a.B * a.H : a.c0);
        float* cout = SAVE ? a.cs + (size_t)t * a.B * a.H : a.cs + (size_t)(t & 1) * a.B * a.H;
        float* y_t = a.y + (size_t)t * a.ys_t;
        const bool last = (t == a.T - 1);
        for (int gc = 0; gc < ngc; ++gc, ++n_chunk) {
          const int buf = n_chunk & 1, use = n_chunk >> 1;
          const uint32_t t_main = tmem_d + ((uint32_t)(eq * 32) << 16) + buf * 256;
          bool waited = false;
          // this warp's two 8-unit groups of the chunk; lane (c8, rl) owns unit j for the 8 rows rg*4 + rl of its quarter
#pragma unroll 1
          for (int pp = 0; pp < 2; ++pp) {
            const int ug = ehalf * 2 + pp;
            const int j = u0 + gc * 32 + ug * 8 + c8;
            const bool act = j < a.H;
            // operands that do not depend on the accumulator: requested before waiting for it
            float bs[4], dh[4], dxc[4], hp[8], cp[8], xv[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              bs[k] = act ? __ldg(a.bias + k * a.H + j) : 0.f;
              dh[k] = act ? __ldg(a.Dh + k * a.H + j) : 0.f;
              dxc[k] = (act && j < a.I) ? __ldg(a.Dx + k * a.I + j) : 0.f;
            }
#pragma unroll
            for (int rg = 0; rg < 8; ++rg) {
              const int m = row0 + eq * 32 + rg * 4 + rl;
              const bool ok = act && m < a.B;
              hp[rg] = (ok && hprev) ? hprev[(size_t)m * hp_sb + j] : 0.f;
              cp[rg] = (ok && cprev) ? cprev[(size_t)m * a.H + j] : 0.f;
              xv[rg] = (ok && j < a.I) ? __ldg(a.x + (long long)t * a.xs_t + (long long)m * a.xs_b + j) : 0.f;
            }
            if (!waited) {
              if (warp == 2) R2_TRACE(22);
              mbar_wait(&bars->accf[buf], use & 1);
              tc_fence_after();
              waited = true;
              if (warp == 2) R2_TRACE(23);
            }
            float v[32];                                        // [gate k][unit] -> after the transpose [gate k][row group]
            tmem_ld_groups(t_main, t_main + 128, ug * 8, 32 + ug * 8, 64 + ug * 8, 96 + ug * 8, v);
            xpose8(v, lane);
            if (act) {
#pragma unroll
              for (int rg = 0; rg < 8; ++rg) {
                const int m = row0 + eq * 32 + rg * 4 + rl;
                if (m < a.B) {
                  float pre[4];
#pragma unroll
                  for (int k = 0; k < 4; ++k) pre[k] = v[k * 8 + rg] + bs[k] + xv[rg] * dxc[k] + hp[rg] * dh[k];
                  const float gi = sigmoidf_acc(pre[0]), gf = sigmoidf_acc(pre[1]);
                  const float go = sigmoidf_acc(pre[2]), gn = tanhf_acc(pre[3]);
                  const float c = fmaf(gf, cp[rg], gi * gn);
                  const float h = go * tanhf_acc(c);
                  y_t[(size_t)m * a.ys_b + j] = h;
                  const float hi = split_hi(h);
                  a.hop_hi[(size_t)m * a.Hp + j] = hi;
                  a.hop_lo[(size_t)m * a.Hp + j] = split_lo(h, hi);
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
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->acce[buf]);
          if (warp == 2) R2_TRACE(24);
        }
        fence_proxy_async_all();
      }
      __syncthreads();       // h_t operand tiles are complete before the next step's TMA reads them (same CTA only)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(512));
  }
}


// ------------------------------------------------------------------------------------------------- backward
// Reverse-time recurrence of the same decomposition (autograd replay of V/models/vmlmf.py:95-125 / vmlmf_lm.py:222-269):
//   phase 1   partial dzc_s[128, KP] = dPre_t[:, 4 gates x slice s] * [Bm | Vx][slice s rows, :]     K = 4*HS (K-split)
//   exchange  dzc = sum of the partials -> dz_t (kept for dA = Hprev^T dZ), dzx_t (kept for dUx, dX), tf32 operand copy
//   phase 2   dh_{t-1}[128, slice s] = dz_t * A[slice s, :]^T + sum_k dPre_t,k (.) Dh_k               K = RH
//             epilogue = the gate-gradient algebra of step t-1 on the fresh dh_{t-1}: dPre_{t-1} (exact copy for the
//             time-parallel weight-gradient GEMMs + tf32 hi / lo operand copy), dc_{t-2}, the Dh term of dh_{t-2}.
// The weight gradients themselves (dBm, dVx, dA, dUx, dX, column sums) are contractions over all T*B rows: time-parallel
// tcgen05 GEMMs after this kernel (generic_bwd_tp).
struct BwdArgs {
  const float* gates;         // [T,B,4,H]
  const float* cs;            // [T,B,H]
  const float* c0;            // [B,H] or null
  const float* dy; long long dys_t, dys_b;      // or null
  const float *dhT, *dcT;     // [B,H] or null
  const float* Dh;
  float *dh0, *dc0;           // [B,H] or null
  float* dpre;                // [T*B, 4, Hp]
  float* dz_all;              // [T*B, zp]
  float* dzx_all;             // [T*B, zxp]
  float* dpo_lo;              // [B, 4, Hp]   tf32 lo part of dPre_t (the hi part is dPre itself: the tensor core truncates)
  float *dzo_hi, *dzo_lo;     // [B, zp]      tf32 operand copy of dz_t
  float *dhrun, *dcrun;       // [B, Hp]
  float* part;                // [groups, NP, 128, KPp]   NP = CS * KSPLIT partials
  unsigned int* sync;         // [groups, 32] group barrier counters, zeroed before the launch
  int T, B, H, RX, RH;
  int Hp, HS, CS, zp, zxp, KZP, KPp, KSPLIT;
};

// gate-gradient algebra of one (row m, unit j) at step tq, in two halves so that a warp can have the loads of eight
// cells in flight before the first store (the scratch buffers it writes may alias what it reads as far as the compiler
// can tell, so loads are hoisted by hand).  Saved activations go through the read-only path.
struct PwIn { float gi, gf, go, gn, ct, cp, dyv, dhs, dcin; };
// loads only, no arithmetic on the loaded values (a use would make the warp wait for the load right here and defeat the
// prefetch): dyv = dy_t, dhs = dh seed or the Dh term kept in dhrun, dcin = dc seed or dcrun
template <class Args>
__device__ __forceinline__ void pw_load(const Args& a, int tq, int m, int j, bool seed, PwIn& in) {
  const size_t rowq = (size_t)tq * a.B + m;
  const float* g = a.gates + rowq * 4 * a.H + j;
  in.gi = __ldg(g); in.gf = __ldg(g + a.H); in.go = __ldg(g + 2 * a.H); in.gn = __ldg(g + 3 * a.H);
  in.ct = __ldg(a.cs + rowq * a.H + j);
  in.cp = tq > 0 ? __ldg(a.cs + (rowq - a.B) * a.H + j) : (a.c0 ? __ldg(a.c0 + (size_t)m * a.H + j) : 0.f);
  in.dyv = a.dy ? __ldg(a.dy + (long long)tq * a.dys_t + (long long)m * a.dys_b + j) : 0.f;
  if (seed) {
    in.dhs = a.dhT ? __ldg(a.dhT + (size_t)m * a.H + j) : 0.f;
    in.dcin = a.dcT ? __ldg(a.dcT + (size_t)m * a.H + j) : 0.f;
  } else {
    in.dhs = a.dhrun[(size_t)m * a.Hp + j];
    in.dcin = a.dcrun[(size_t)m * a.Hp + j];
  }
}
__device__ __forceinline__ void pw_finish(const BwdArgs& a, int tq, int m, int j, const PwIn& in, float dh, const float (&dhc)[4]) {
  const size_t rowq = (size_t)tq * a.B + m;
  const float tcv = tanhf_acc(in.ct);
  const float dc = fmaf(dh * in.go, 1.f - tcv * tcv, in.dcin);
  float d[4];
  d[0] = dc * in.gn * in.gi * (1.f - in.gi);
  d[1] = dc * in.cp * in.gf * (1.f - in.gf);
  d[2] = dh * tcv * in.go * (1.f - in.go);
  d[3] = dc * in.gi * (1.f - in.gn * in.gn);
  float* o = a.dpre + rowq * 4 * a.Hp + j;
  float* ol = a.dpo_lo + (size_t)m * 4 * a.Hp + j;
  float sdh = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    o[(size_t)k * a.Hp] = d[k];                               // exact copy for the time-parallel GEMMs AND the hi operand of phase 1
    ol[(size_t)k * a.Hp] = split_lo(d[k], d[k]);
    sdh = fmaf(d[k], dhc[k], sdh);
  }
  a.dcrun[(size_t)m * a.Hp + j] = dc * in.gf;
  a.dhrun[(size_t)m * a.Hp + j] = sdh;
}

template <int kVariant>      // a template only for linkage: the header is included by two translation units
__global__ void __launch_bounds__(kThreads, 1)
r2_bwd_kernel(const __grid_constant__ CUtensorMap m_dpre, const __grid_constant__ CUtensorMap m_dpo_lo,
              const __grid_constant__ CUtensorMap m_w2t_hi, const __grid_constant__ CUtensorMap m_w2t_lo,
              const __grid_constant__ CUtensorMap m_dzo_hi, const __grid_constant__ CUtensorMap m_dzo_lo,
              const __grid_constant__ CUtensorMap m_ap_hi, const __grid_constant__ CUtensorMap m_ap_lo, const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const Smem sm = carve(smem_raw);
  Bars* bars = sm.bars;
  const int warp = warp_index(), lane = threadIdx.x & 31;      // broadcast: role branches are warp-uniform for the compiler
  // per-epilogue-warp transpose tile behind the barriers: accumulator rows (thread = batch row) -> lane = hidden unit, so
  // that every global access of the gate-gradient algebra is one full 128-byte line (32 consecutive units of one row)
  float* const xt = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sm.bars) + 256) + (warp >= 2 ? (warp - 2) * (kXposeBytes / 4) : 0);
  const int CS = a.CS, KSPLIT = a.KSPLIT, NP = CS * KSPLIT;
  const int s_rank = (int)blockIdx.x % CS, cid = (int)blockIdx.x / CS, ncl = (int)gridDim.x / CS;
  const int ntiles = (a.B + BM - 1) / BM;
  unsigned int* const sync_ctr = a.sync + cid * 32;          // one 128-byte line per group
  unsigned int epoch = 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bars->accf[b], 1); mbar_init(&bars->acce[b], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bars->tmem_slot;

  const int u0 = s_rank * a.HS;
  const int uvalid = min(a.HS, a.H - u0);
  const int nkh = (uvalid + BK - 1) / BK;                    // K tiles per gate in phase 1
  const int KT = 4 * nkh;                                    // phase 1 K tiles (gate-major)
  const int kseg = (KT + KSPLIT - 1) / KSPLIT;               // K tiles per accumulation segment
  const int nch1 = (a.KPp + 127) / 128;                      // phase 1 chunks (128 columns of [dz | dzx])
  const int nch2 = (uvalid + 127) / 128;                     // phase 2 chunks (128 units)
  const int nkz = (a.RH + BK - 1) / BK;                      // phase 2 K tiles

  const int eq = warp & 3, ehalf = (warp - 2) >> 2;
  const int rl = lane & 3, c8 = lane >> 2;

  uint32_t n_tile = 0, n_chunk = 0;

  for (int tile = cid; tile < ntiles; tile += ncl) {
    const int row0 = tile * BM;
    // ---- seed: gate-gradient algebra of the last step with dh = dhT, dc = dcT ----
    if (warp >= 2) {
      const int tq = a.T - 1;
      for (int ub = ehalf * 32; ub < uvalid; ub += 64) {          // lane = unit: 128-byte lines
        const int j = u0 + ub + lane;
        if (j < a.H) {
          float dhc[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) dhc[k] = __ldg(a.Dh + k * a.H + j);
#pragma unroll 1
          for (int r0 = 0; r0 < 32; r0 += 4) {
            PwIn in[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int m = row0 + eq * 32 + r0 + u;
              if (m < a.B) pw_load(a, tq, m, j, true, in[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int m = row0 + eq * 32 + r0 + u;
              if (m < a.B) pw_finish(a, tq, m, j, in[u], in[u].dyv + in[u].dhs, dhc);
            }
          }
        }
      }
      fence_proxy_async_all();
    }
    __syncthreads();
    for (int t = a.T - 1; t >= 0; --t) {
      // ======================================= phase 1 =======================================
      if (warp == 0) {
        fence_proxy_async_all();
        for (int c = 0; c < nch1; ++c)
          for (int ka = 0; ka < KT; ++ka, ++n_tile) {
            const int k = ka / nkh, kt = ka - k * nkh;
            const int s = n_tile % kStages, it = n_tile / kStages;
            if (it > 0) mbar_wait(&bars->empty[s], (it - 1) & 1);
            uint8_t* st = sm.stages + s * kStageBytes;
            if (elect_one()) {
              mbar_arrive_expect_tx(&bars->full[s], kStageBytes);
              tma_load_4d(st, &m_dpre, u0 + kt * BK, k, row0, t, &bars->full[s]);
              tma_load_3d(st + kTile, &m_dpo_lo, u0 + kt * BK, k, row0, &bars->full[s]);
              tma_load_3d(st + 2 * kTile, &m_w2t_hi, u0 + kt * BK, k, c * 128, &bars->full[s]);
              tma_load_3d(st + 3 * kTile, &m_w2t_lo, u0 + kt * BK, k, c * 128, &bars->full[s]);
            }
          }
      } else if (warp == 1) {
        for (int c = 0; c < nch1; ++c) {
          const int ncol = min(128, a.KPp - c * 128);
          const uint32_t idesc = make_idesc(BM, ncol);
          for (int ks = 0; ks < KSPLIT; ++ks, ++n_chunk) {
            const int buf = n_chunk & 1, use = n_chunk >> 1;
            if (use > 0) mbar_wait(&bars->acce[buf], (use - 1) & 1);
            tc_fence_after();
            const uint32_t acc_main = tmem_d + buf * 256, acc_cross = acc_main + 128;
            const int ka0 = ks * kseg, ka1 = min(KT, ka0 + kseg);
            for (int ka = ka0; ka < ka1; ++ka, ++n_tile) {
              const int kt = ka % nkh;
              const int s = n_tile % kStages, it = n_tile / kStages;
              mbar_wait(&bars->full[s], it & 1);
              tc_fence_after();
              if (ka == ka0) R2_TRACE(10);
              issue_tile(smem_u32(sm.stages + s * kStageBytes), acc_main, acc_cross, idesc, tile_ksteps(uvalid, kt), ka == ka0);
              commit_elect(&bars->empty[s]);
            }
            commit_elect(&bars->accf[buf]);
            R2_TRACE(11);
          }
        }
      } else {
        for (int c = 0; c < nch1; ++c) {
          const int ncol = min(128, a.KPp - c * 128);
          for (int ks = 0; ks < KSPLIT; ++ks, ++n_chunk) {
            const int buf = n_chunk & 1, use = n_chunk >> 1;
            mbar_wait(&bars->accf[buf], use & 1);
            tc_fence_after();
            if (warp == 2) R2_TRACE(20);
            const uint32_t t_main = tmem_d + ((uint32_t)(eq * 32) << 16) + buf * 256;
#pragma unroll 1
            for (int pp = 0; pp < 2; ++pp) {
              const int cb = (ehalf * 2 + pp) * 32;
              if (cb >= ncol) break;
              float v[32];
              tmem_ld_groups(t_main, t_main + 128, cb, cb + 8, cb + 16, cb + 24, v);
              xpose_vec4(v, lane);
              const int n = c * 128 + cb + (lane >> 2) * 4;            // first of this lane's 4 columns of [dz | pad | dzx | pad]
#pragma unroll
              for (int rg = 0; rg < 8; ++rg) {
                const int r = eq * 32 + rg * 4 + rl, m = row0 + r;
                if (m < a.B) {
                  const float4 val = make_float4(v[rg * 4], v[rg * 4 + 1], v[rg * 4 + 2], v[rg * 4 + 3]);
                  if (NP > 1) {
                    *reinterpret_cast<float4*>(a.part + (((size_t)cid * NP + s_rank * KSPLIT + ks) * BM + r) * a.KPp + n) = val;
                  } else if (n < a.zp) {                               // zp, KZP, zxp are multiples of 4
                    *reinterpret_cast<float4*>(a.dz_all + ((size_t)t * a.B + m) * a.zp + n) = val;
                    float4 hi, lo;
                    hi.x = split_hi(val.x); hi.y = split_hi(val.y); hi.z = split_hi(val.z); hi.w = split_hi(val.w);
                    lo.x = split_lo(val.x, hi.x); lo.y = split_lo(val.y, hi.y); lo.z = split_lo(val.z, hi.z); lo.w = split_lo(val.w, hi.w);
                    *reinterpret_cast<float4*>(a.dzo_hi + (size_t)m * a.zp + n) = hi;
                    *reinterpret_cast<float4*>(a.dzo_lo + (size_t)m * a.zp + n) = lo;
                  } else if (n >= a.KZP && n < a.KZP + a.zxp) {
                    *reinterpret_cast<float4*>(a.dzx_all + ((size_t)t * a.B + m) * a.zxp + (n - a.KZP)) = val;
                  }
                }
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->acce[buf]);
          }
        }
        fence_proxy_async_all();
        if (warp == 2) R2_TRACE(21);
      }
      // ======================================= exchange =======================================
      if (NP > 1) {
        if (CS > 1) group_sync(sync_ctr, epoch, CS); else __syncthreads();
        if (warp == 2) R2_TRACE(30);
        if (warp >= 2) {
          const int rows_valid = min(BM, a.B - row0);
          const int rpc = (rows_valid + CS - 1) / CS;
          const int r_lo = s_rank * rpc, r_hi = min(rows_valid, r_lo + rpc);
          const int et = threadIdx.x - 64;
          const int zp4 = a.zp >> 2, w4 = (a.zp + a.zxp) >> 2;               // float4 columns: dz then dzx
          const int nel = (r_hi - r_lo) * w4;
          const float* pbase = a.part + (size_t)cid * NP * BM * a.KPp;
          const size_t pstride = (size_t)BM * a.KPp;
          for (int e = et; e < nel; e += 32 * kEpiWarps) {
            const int r = r_lo + e / w4, c4 = e - (e / w4) * w4;
            const int n = c4 < zp4 ? c4 * 4 : a.KZP + (c4 - zp4) * 4;
            const float* src = pbase + (size_t)r * a.KPp + n;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int p0 = 0; p0 < NP; p0 += 8) {                              // eight 16-byte loads in flight; fixed summation order
              float4 pv[8];
#pragma unroll
              for (int q = 0; q < 8; ++q)
                pv[q] = p0 + q < NP ? __ldcg(reinterpret_cast<const float4*>(src + (size_t)(p0 + q) * pstride)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int q = 0; q < 8; ++q) { v.x += pv[q].x; v.y += pv[q].y; v.z += pv[q].z; v.w += pv[q].w; }
            }
            const int m = row0 + r;
            if (c4 < zp4) {
              const int cc = c4 * 4;
              *reinterpret_cast<float4*>(a.dz_all + ((size_t)t * a.B + m) * a.zp + cc) = v;
              float4 hi, lo;
              hi.x = split_hi(v.x); hi.y = split_hi(v.y); hi.z = split_hi(v.z); hi.w = split_hi(v.w);
              lo.x = split_lo(v.x, hi.x); lo.y = split_lo(v.y, hi.y); lo.z = split_lo(v.z, hi.z); lo.w = split_lo(v.w, hi.w);
              *reinterpret_cast<float4*>(a.dzo_hi + (size_t)m * a.zp + cc) = hi;
              *reinterpret_cast<float4*>(a.dzo_lo + (size_t)m * a.zp + cc) = lo;
            } else {
              *reinterpret_cast<float4*>(a.dzx_all + ((size_t)t * a.B + m) * a.zxp + (c4 - zp4) * 4) = v;
            }
          }
          fence_proxy_async_all();
          if (warp == 2) R2_TRACE(31);
        }
        if (CS > 1) group_sync(sync_ctr, epoch, CS); else __syncthreads();
        if (warp == 2) R2_TRACE(32);
      } else {
        __syncthreads();
      }
      // ======================================= phase 2 =======================================
      if (warp == 0) {
        fence_proxy_async_all();
        for (int c = 0; c < nch2; ++c)
          for (int kt = 0; kt < nkz; ++kt, ++n_tile) {
            const int s = n_tile % kStages, it = n_tile / kStages;
            if (it > 0) mbar_wait(&bars->empty[s], (it - 1) & 1);
            uint8_t* st = sm.stages + s * kStageBytes;
            if (elect_one()) {
              mbar_arrive_expect_tx(&bars->full[s], kStageBytes);
              tma_load_2d(st, &m_dzo_hi, kt * BK, row0, &bars->full[s]);
              tma_load_2d(st + kTile, &m_dzo_lo, kt * BK, row0, &bars->full[s]);
              tma_load_2d(st + 2 * kTile, &m_ap_hi, kt * BK, u0 + c * 128, &bars->full[s]);
              tma_load_2d(st + 3 * kTile, &m_ap_lo, kt * BK, u0 + c * 128, &bars->full[s]);
            }
          }
      } else if (warp == 1) {
        const uint32_t idesc = make_idesc(BM, 128);
        for (int c = 0; c < nch2; ++c, ++n_chunk) {
          const int buf = n_chunk & 1, use = n_chunk >> 1;
          if (use > 0) mbar_wait(&bars->acce[buf], (use - 1) & 1);
          tc_fence_after();
          const uint32_t acc_main = tmem_d + buf * 256, acc_cross = acc_main + 128;
          for (int kt = 0; kt < nkz; ++kt, ++n_tile) {
            const int s = n_tile % kStages, it = n_tile / kStages;
            mbar_wait(&bars->full[s], it & 1);
            tc_fence_after();
            if (kt == 0) R2_TRACE(12);
            issue_tile(smem_u32(sm.stages + s * kStageBytes), acc_main, acc_cross, idesc, tile_ksteps(a.RH, kt), kt == 0);
            commit_elect(&bars->empty[s]);
          }
          commit_elect(&bars->accf[buf]);
          R2_TRACE(13);
        }
      } else {
        for (int c = 0; c < nch2; ++c, ++n_chunk) {
          const int buf = n_chunk & 1, use = n_chunk >> 1;
          // this warp's two 32-unit passes of the chunk.  The accumulator arrives with thread = batch row; a swizzled
          // shared tile private to the warp turns it into lane = hidden unit, and the rows are then walked four at a time
          // (runtime loop: the body holds the gate-gradient algebra of four cells and stays inside the instruction cache) with
          // the next four rows' saved activations already requested.
          if (warp == 2) R2_TRACE(22);
          mbar_wait(&bars->accf[buf], use & 1);
          tc_fence_after();
          if (warp == 2) R2_TRACE(23);
          const uint32_t t_main = tmem_d + ((uint32_t)(eq * 32) << 16) + buf * 256;
#pragma unroll 1
          for (int pp = 0; pp < 2; ++pp) {
            const int cb = (ehalf * 2 + pp) * 32;
            if (c * 128 + cb >= uvalid) break;
            {
              float v[32];
              tmem_ld_groups(t_main, t_main + 128, cb, cb + 8, cb + 16, cb + 24, v);
              __syncwarp();                                      // the previous pass's readers are done with the tile
#pragma unroll
              for (int cc = 0; cc < 32; ++cc) xt[lane * 32 + ((cc ^ lane) & 31)] = v[cc];
              __syncwarp();
            }
            const int j = u0 + c * 128 + cb + lane;
            const bool act = j < a.H;
            const int mbase = row0 + eq * 32;
            if (t > 0) {
              float dhc[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) dhc[k] = act ? __ldg(a.Dh + k * a.H + j) : 0.f;
              PwIn pa[4];
              if (act) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  if (mbase + u < a.B) pw_load(a, t - 1, mbase + u, j, false, pa[u]);
              }
#pragma unroll 1
              for (int r0 = 0; r0 < 32; r0 += 4) {
                if (mbase + r0 >= a.B) break;
                PwIn pn[4];
                if (act && r0 + 4 < 32) {
#pragma unroll
                  for (int u = 0; u < 4; ++u)
                    if (mbase + r0 + 4 + u < a.B) pw_load(a, t - 1, mbase + r0 + 4 + u, j, false, pn[u]);
                }
                if (act) {
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const int r = r0 + u, m = mbase + r;
                    if (m < a.B) pw_finish(a, t - 1, m, j, pa[u], pa[u].dyv + pa[u].dhs + xt[r * 32 + ((lane ^ r) & 31)], dhc);
                  }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) pa[u] = pn[u];
              }
            } else if (act) {
#pragma unroll 4
              for (int r = 0; r < 32; ++r) {
                const int m = mbase + r;
                if (m < a.B) {
                  if (a.dh0) a.dh0[(size_t)m * a.H + j] = a.dhrun[(size_t)m * a.Hp + j] + xt[r * 32 + ((lane ^ r) & 31)];
                  if (a.dc0) a.dc0[(size_t)m * a.H + j] = a.dcrun[(size_t)m * a.Hp + j];
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->acce[buf]);
          if (warp == 2) R2_TRACE(24);
        }
        fence_proxy_async_all();
        if (warp == 2) R2_TRACE(25);
      }
      __syncthreads();
      if (warp == 2) R2_TRACE(40);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(512));
  }
}

}  // namespace r2
}  // namespace vmlmf
