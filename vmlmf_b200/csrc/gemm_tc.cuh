// gemm_tc.cuh -- fp32-accurate GEMM on the 5th-generation tensor cores (sm_100a):
//     C[M,N] (+epilogue) = A[M,K] * B[N,K]^T          A, B fp32, K-major (row pitch in floats, % 4 == 0)
// TMA (cp.async.bulk.tensor, 128-byte swizzle) stages raw fp32 tiles in shared memory; four "split" warps turn
// every tile into a tf32 hi part (in place) and a lo part (second buffer, same swizzled offsets); one elected
// thread issues tcgen05.mma.kind::tf32 three times per 8-deep k-step (lo*hi, hi*lo, hi*hi -- 3xTF32 error
// compensation, fp32 accumulate in TMEM); the same four warps read the accumulator back with tcgen05.ld and run
// the epilogue functor.  Pipeline: kStages smem stages, mbarriers  full (TMA landed) -> split (hi/lo ready) ->
// empty (MMAs that read the stage have completed, via tcgen05.commit).
//
// This is the time-parallel x-side product of the VMLMF cells ((x U_x) V_x^T with bias and the vector-
// multiplication term fused in the epilogue: V/models/vmlmf.py:98,103-104,109; vmlmf_lm.py:246,256) and the
// per-timestep hidden-side products of the generic regime (V/models/vmlmf_lm.py:247; vmlmf.py:99).
//
// B can also be a rank-3 view [4][H][K] (the four gate blocks of a [4H,K] factor): the tile then holds
// 4 gates x 32 units, so one accumulator row carries i,f,o,n of 32 hidden units and the LSTM cell update runs in
// the epilogue (EpiGate32).
#pragma once
#include <cuda.h>
#include <type_traits>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"

namespace vmlmf {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;          // tile: 128 x 128, 32 fp32 (= 128 bytes, one swizzle row) deep
constexpr int kStages = 3;
constexpr int kTileBytes = BM * BK * 4;             // 16 KB per operand tile (BN == BM)
constexpr int kStageBytes = 4 * kTileBytes;         // A hi | A lo | B hi | B lo
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int kEpiPitch = BN + 4;                    // staged accumulator row pitch (floats): conflict-free both ways
constexpr int kThreads = 192;                       // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: split + epilogue
constexpr int kAcc = 4;                             // accumulators: hi*hi round-robin over 3, cross terms in the 4th
constexpr int kTmemCols = kAcc * BN;                // 4 x (128 lanes x 128 fp32 columns) = all 512 TMEM columns

// ------------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ float tf32r(float v) { return tf32_rna(v); }
// v minus its tf32 truncation (exact; |result| < 2^-10 |v|)
__device__ __forceinline__ float tf32_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One lane of a fully active warp.  The producer / MMA warps run their loops with all 32 lanes in warp-uniform control flow and
// elect right at the asynchronous instructions: inside an `if (lane == 0)` region the compiler keeps descriptors and addresses
// in per-thread registers and wraps every tcgen05.mma / TMA in an elect + R2UR.BROADCAST + loop sequence (~135 cycles per MMA
// measured, ~45-60 for the uniform form: tools/ubench_mma.cu).  The warp index must come from warp_index() for that.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ int warp_index() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
// shared-memory matrix descriptor: K-major tile, rows of 128 bytes, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);            // start address
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused with 128B swizzle, K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = BN
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 8 consecutive fp32 columns of one accumulator -> 8 registers per thread
__device__ __forceinline__ void tmem_ld8_raw(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// sum of the `nacc` accumulators that were written (column offset BN apart); waits for the loads
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int nhi, float (&v)[8], bool cross = true) {
  float a[8], b[8];
  tmem_ld8_raw(taddr, v);                 // hi*hi accumulator 0
  if (cross) tmem_ld8_raw(taddr + 3 * 128, a);       // cross terms (not written in the single-pass mode)
  tmem_ld_wait();
  if (cross) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += a[i];
  }
  if (nhi > 1) {
    tmem_ld8_raw(taddr + 128, a);
    if (nhi > 2) tmem_ld8_raw(taddr + 256, b);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += a[i];
    if (nhi > 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += b[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------- epilogues
// The accumulator tile is staged in shared memory and handed out one ROW per call to a whole warp:
// lane l receives the four values of columns n0 + l + 32*i (i = 0..3), so every global access is a contiguous
// 128-byte segment per warp instruction.
// Epilogues whose rows read more global operands than they write (XP, dX) declare `Col` (per-column constants, loaded once
// per CTA), `In` (per-row operands) and cols() / load() / finish(): the kernel then keeps the operands of four rows in
// flight before the first is consumed.  With the plain operator() form every row waited a full L2 / HBM round trip for its
// own operands: 80 - 90 us for a [700 x 650] tile grid, longer than its whole mainloop.
template <class E, class = void> struct epi_prefetch { static constexpr bool value = false; };
template <class E> struct epi_prefetch<E, std::void_t<typename E::Col>> { static constexpr bool value = true; };

struct EpiStoreTC {                      // C[m, n] = v (+ C[m, n] when beta)
  float* C; long long ldc; int beta;
  static constexpr bool kGate = false;
  __device__ void operator()(int m, int n0, int N, int lane, const float (&v)[4]) const {
    float* c = C + (size_t)m * ldc + n0 + lane;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (n0 + lane + 32 * i < N) c[32 * i] = beta ? c[32 * i] + v[i] : v[i];
  }
};
struct EpiBiasTC {                       // C[m, n] = v + bias[n] (+ C[m, n] when beta); bias may be null
  float* C; long long ldc; const float* bias; int beta;
  static constexpr bool kGate = false;
  __device__ void operator()(int m, int n0, int N, int lane, const float (&v)[4]) const {
    float* c = C + (size_t)m * ldc + n0 + lane;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + lane + 32 * i;
      if (n < N) {
        const float r = v[i] + (bias ? __ldg(bias + n) : 0.f);
        c[32 * i] = beta ? c[32 * i] + r : r;
      }
    }
  }
};
struct EpiPartialTC {                    // split-K partial: part[blockIdx.z][m][n] = v  (summed later in a fixed order)
  float* part; int M, N;
  static constexpr bool kGate = false;
  __device__ void operator()(int m, int n0, int, int lane, const float (&v)[4]) const {
    float* c = part + ((size_t)blockIdx.z * M + m) * N + n0 + lane;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (n0 + lane + 32 * i < N) c[32 * i] = v[i];
  }
};
struct EpiXPTC {                         // XP[m, kH+j] = v + bias[kH+j] + [j<I] x[m,j] Dx[k,j]
  float* xp; const float* bias; const float* x; long long xs_t, xs_b; int Bsz; const float* Dx; int H, I;
  static constexpr bool kGate = false;
  struct Col { float bias[4], dx[4]; int j[4]; };          // j < 0: no x term (or column outside the matrix)
  struct In { float x[4]; };
  __device__ void cols(int n0, int N, int lane, Col& c) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int nn = n0 + lane + 32 * i;
      c.bias[i] = 0.f; c.dx[i] = 0.f; c.j[i] = -1;
      if (nn < N) {
        const int k = nn / H, j = nn - k * H;
        c.bias[i] = __ldg(bias + nn);
        if (j < I) { c.j[i] = j; c.dx[i] = __ldg(Dx + k * I + j); }
      }
    }
  }
  __device__ void load(int m, int, int, int, const Col& c, In& in) const {
    const float* xr = x + (long long)(m / Bsz) * xs_t + (long long)(m % Bsz) * xs_b;
#pragma unroll
    for (int i = 0; i < 4; ++i) in.x[i] = c.j[i] >= 0 ? __ldg(xr + c.j[i]) : 0.f;
  }
  __device__ void finish(int m, int n0, int N, int lane, const float (&v)[4], const Col& c, const In& in) const {
    float* o = xp + (size_t)m * N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int nn = n0 + lane + 32 * i;
      if (nn < N) o[nn] = fmaf(in.x[i], c.dx[i], v[i] + c.bias[i]);
    }
  }
};
// Gate epilogue (B is the rank-3 view, tile = 4 gates x 32 units): lane l receives i,f,o,n of hidden unit j0 + l.
struct EpiGate32 {
  const float* xp_t;        // XP rows of this step   [B, 4H]
  const float* hprev; long long hp_sb;    // h_{t-1}[b] = hprev + b*hp_sb (null = zeros)
  const float* cprev;                     // [B,H] or null
  const float* Dh;
  float* y_t; long long y_sb;
  float* c_out;                           // [B,H]
  float* gates_t;                         // [B,4,H] or null
  float *hT, *cT;                         // written on the last step only (else null)
  float* hpad; int hp4;                   // padded copy of h_t, row pitch hp4 (% 4 == 0): next step's TMA operand
  int H;
  static constexpr bool kGate = true;
  struct In { float xp[4], hp, cp; };
  // global operands of (row m, unit j): issued for several rows before any of them is consumed
  __device__ void load(int m, int j, In& in) const {
    const float* xr = xp_t + (size_t)m * 4 * H + j;
#pragma unroll
    for (int k = 0; k < 4; ++k) in.xp[k] = __ldg(xr + (size_t)k * H);
    in.hp = hprev ? hprev[(size_t)m * hp_sb + j] : 0.f;
    in.cp = cprev ? cprev[(size_t)m * H + j] : 0.f;
  }
  __device__ void finish(int m, int j, const float (&g)[4], const In& in, const float (&dh)[4]) const {
    float pre[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) pre[k] = g[k] + in.xp[k] + in.hp * dh[k];
    const float gi = sigmoidf_acc(pre[0]), gf = sigmoidf_acc(pre[1]);
    const float go = sigmoidf_acc(pre[2]), gn = tanhf_acc(pre[3]);
    const float c = fmaf(gf, in.cp, gi * gn);
    const float h = go * tanhf_acc(c);
    y_t[(size_t)m * y_sb + j] = h;
    hpad[(size_t)m * hp4 + j] = h;
    c_out[(size_t)m * H + j] = c;
    if (gates_t) {
      float* gp = gates_t + (size_t)m * 4 * H + j;
      gp[0] = gi; gp[H] = gf; gp[2 * H] = go; gp[3 * H] = gn;
    }
    if (hT) { hT[(size_t)m * H + j] = h; cT[(size_t)m * H + j] = c; }
  }
};

// ------------------------------------------------------------------------------------------------- kernel
// grid = (ceil(N / BN) [or ceil(H / 32) for the gate view], ceil(M / BM)); one output tile per CTA.
template <class Epi>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                              const __grid_constant__ CUtensorMap mapB, int M, int N,
                                                              int K, int kb_per_split, Epi epi, int fast) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;                        // [kStages] TMA bytes landed
  uint64_t* split = bars + kStages;             // [kStages] hi/lo tiles written (4 warps arrive)
  uint64_t* empty = bars + 2 * kStages;         // [kStages] MMAs reading the stage completed
  uint64_t* accf = bars + 3 * kStages;          // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);

  const int warp = warp_index(), lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM;
  const int n_tile = blockIdx.x;
  // split-K: blockIdx.z owns k-blocks [kb0, kb0 + nkb)
  const int nkb_total = (K + BK - 1) / BK;
  const int kb0 = blockIdx.z * kb_per_split;
  const int nkb = (nkb_total - kb0) < kb_per_split ? (nkb_total - kb0) : kb_per_split;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&split[s], 4); mbar_init(&empty[s], 1); }
    mbar_init(accf, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kStages, it = kb / kStages;
      if (it > 0) mbar_wait(&empty[s], (it - 1) & 1);
      uint8_t* st = smem + s * kStageBytes;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[s], 2 * kTileBytes);
        tma_load_2d(st, &mapA, (kb0 + kb) * BK, m0, &full[s]);
        if (Epi::kGate) tma_load_3d(st + 2 * kTileBytes, &mapB, (kb0 + kb) * BK, n_tile * 32, 0, &full[s]);
        else tma_load_2d(st + 2 * kTileBytes, &mapB, (kb0 + kb) * BK, n_tile * BN, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = make_idesc(BM, BN);
    int rr = 0;                                          // kk % 3 without a division in the issue loop
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kStages, it = kb / kStages;
      mbar_wait(&split[s], it & 1);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * kStageBytes);
      // descriptors of the four operand tiles; a k-step (8 tf32 = 32 bytes inside the 128-byte swizzle row) adds 2
      const uint64_t da_hi = make_desc(a_addr), da_lo = make_desc(a_addr + kTileBytes);
      const uint64_t db_hi = make_desc(a_addr + 2 * kTileBytes), db_lo = make_desc(a_addr + 3 * kTileBytes);
      // The tensor core adds into the fp32 accumulator with truncation, so the error grows with the number of
      // accumulations into one accumulator (measured ~3e-8 relative per add).  The large hi*hi products are
      // spread round-robin over three accumulators, the small cross terms go to a fourth; the epilogue adds
      // the four in fp32.
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const int kk = kb * (BK / 8) + k;
          if (!fast) {                                     // fast mode (single-pass TF32): no error-compensation terms
            mma_tf32_ss(tmem_d + 3 * BN, da_lo + 2 * k, db_hi + 2 * k, idesc, kk ? 1u : 0u);
            mma_tf32_ss(tmem_d + 3 * BN, da_hi + 2 * k, db_lo + 2 * k, idesc, 1u);
          }
          const int r3 = (rr + k) % 3;                     // rr in 0..2, k a constant: folds to compares
          mma_tf32_ss(tmem_d + r3 * BN, da_hi + 2 * k, db_hi + 2 * k, idesc, kk >= 3 ? 1u : 0u);
        }
        mma_commit(&empty[s]);                           // stage reusable once these MMAs have read it
      }
      rr = (rr + BK / 8) % 3;
    }
    if (elect_one()) mma_commit(accf);                   // accumulator complete
  } else {
    // ================= split warps, then epilogue =================
    const int sw = warp - 2;                             // 0..3
    const int st_tid = sw * 32 + lane;                   // 0..127
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kStages, it = kb / kStages;
      mbar_wait(&full[s], it & 1);
      float4* a_hi = reinterpret_cast<float4*>(smem + s * kStageBytes);
      float4* a_lo = a_hi + kTileBytes / 16;
      float4* b_hi = a_hi + 2 * kTileBytes / 16;
      float4* b_lo = a_hi + 3 * kTileBytes / 16;
      // elementwise at identical offsets: the swizzle pattern of the TMA write is preserved
#pragma unroll 4
      for (int i = st_tid; i < (fast ? 0 : kTileBytes / 16); i += 128) {     // fast mode: the raw tile is the only operand
        const float4 va = a_hi[i], vb = b_hi[i];
#ifdef VMLMF_SPLIT_ROUND
        float4 ha, la, hb, lb;
        ha.x = tf32r(va.x); ha.y = tf32r(va.y); ha.z = tf32r(va.z); ha.w = tf32r(va.w);
        la.x = tf32r(va.x - ha.x); la.y = tf32r(va.y - ha.y); la.z = tf32r(va.z - ha.z); la.w = tf32r(va.w - ha.w);
        hb.x = tf32r(vb.x); hb.y = tf32r(vb.y); hb.z = tf32r(vb.z); hb.w = tf32r(vb.w);
        lb.x = tf32r(vb.x - hb.x); lb.y = tf32r(vb.y - hb.y); lb.z = tf32r(vb.z - hb.z); lb.w = tf32r(vb.w - hb.w);
        a_hi[i] = ha; a_lo[i] = la; b_hi[i] = hb; b_lo[i] = lb;
#else
        // The tensor core reads the top 19 bits of a tf32 operand and ignores the low 13 mantissa bits, so the raw fp32 tile
        // the TMA delivered IS its own (truncated) hi part: only the remainder is computed and written -- a third less
        // shared-memory traffic in the stage that bounds this kernel (same trick as the mma.sync kernels, seq_mma.cuh:split4).
        float4 la, lb;
        la.x = tf32_lo(va.x); la.y = tf32_lo(va.y); la.z = tf32_lo(va.z); la.w = tf32_lo(va.w);
        lb.x = tf32_lo(vb.x); lb.y = tf32_lo(vb.y); lb.z = tf32_lo(vb.z); lb.w = tf32_lo(vb.w);
        a_lo[i] = la; b_lo[i] = lb;
#endif
      }
      fence_proxy_async();                               // generic-proxy writes -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&split[s]);
    }
    // ---- epilogue: TMEM -> registers -> shared memory (row per thread), then one row per warp with lane <-> column
    mbar_wait(accf, 0);
    tc_fence_after();
    // the staging tile below reuses the pipeline stages these same four warps wrote (hi / lo split): ordered through the
    // mbarrier chain already, the CTA-level barrier makes that ordering explicit (and visible to racecheck)
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int lane_grp = warp & 3;                       // TMEM lane quarter this warp may access
    const uint32_t tbase = tmem_d + ((uint32_t)(lane_grp * 32) << 16);
    const int nks = nkb * (BK / 8);
    const int nhi = nks < 3 ? nks : 3;                   // hi*hi accumulators that received at least one product
    float* S = reinterpret_cast<float*>(smem);           // [128][kEpiPitch]: the pipeline stages are idle by now
    {
      float* srow = S + (size_t)(lane_grp * 32 + lane) * kEpiPitch;
#pragma unroll 1
      for (int c = 0; c < BN; c += 8) {
        float v[8];
        tmem_ld8(tbase + c, nhi, v, !fast);
        reinterpret_cast<float4*>(srow + c)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(srow + c)[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");       // the four epilogue warps only
    const int n0 = Epi::kGate ? n_tile * 32 : n_tile * BN;
    if constexpr (Epi::kGate) {
      // four rows in flight per warp: all global operands are requested before the first row is finished
      const int j = n0 + lane;
      if (j < epi.H) {
        float dh[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dh[k] = __ldg(epi.Dh + k * epi.H + j);
#pragma unroll 1
        for (int r = sw; r < BM; r += 16) {
          typename Epi::In in[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (m0 + r + 4 * u < M) epi.load(m0 + r + 4 * u, j, in[u]);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (m0 + r + 4 * u < M) {
              const float* srow = S + (size_t)(r + 4 * u) * kEpiPitch + lane;
              const float g[4] = {srow[0], srow[32], srow[64], srow[96]};
              epi.finish(m0 + r + 4 * u, j, g, in[u], dh);
            }
        }
      }
    } else if constexpr (epi_prefetch<Epi>::value) {
      typename Epi::Col col;
      epi.cols(n0, N, lane, col);
#pragma unroll 1
      for (int r = sw; r < BM; r += 16) {
        typename Epi::In in[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (m0 + r + 4 * u < M) epi.load(m0 + r + 4 * u, n0, N, lane, col, in[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (m0 + r + 4 * u < M) {
            const float* srow = S + (size_t)(r + 4 * u) * kEpiPitch + lane;
            const float v[4] = {srow[0], srow[32], srow[64], srow[96]};
            epi.finish(m0 + r + 4 * u, n0, N, lane, v, col, in[u]);
          }
      }
    } else {
      for (int r = sw; r < BM; r += 4) {
        const int m = m0 + r;
        if (m >= M) break;
        const float* srow = S + (size_t)r * kEpiPitch + lane;
        const float v[4] = {srow[0], srow[32], srow[64], srow[96]};
        epi(m, n0, N, lane, v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols));
  }
}

// ------------------------------------------------------------------------------------------------- TN kernel
// C[M,N] = At[K,M]^T Bt[K,N]: both operands stored with the CONTRACTION index as the row (row-major activations
// [T*B rows, features]), i.e. MN-major tensor-core operands.  This is the shape of every weight-gradient contraction
// of the backward (dBm = dPre^T Z, dVx = dPre^T ZX, ...: V/models/vmlmf.py:98-99 replayed by autograd): reading the
// activations in place removes the transposed copies a K-major kernel needs.
// Shared-memory tile of one operand: four blocks of [32 K rows][32 features = 128 bytes] (one TMA box each), blocks 4096
// bytes apart (leading byte offset).  MN-major 32-bit operands have exactly one legal swizzle: 128-byte rows swizzled in
// 32-byte chunks with a 4-row period (descriptor layout type 1, TMA mode SWIZZLE_128B_ATOM_32B), so the K groups are 4
// rows = 512 bytes apart (stride byte offset); one tf32 MMA (K = 8) consumes two of them, the k-step advance is 1024 bytes.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(4096 >> 4) << 16;                  // leading byte offset: next 32-feature block along M / N
  d |= (uint64_t)(512 >> 4) << 32;                   // stride byte offset: next group of 4 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                            // SWIZZLE_128B_BASE32B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N) { return make_idesc(M, N) | (1u << 15) | (1u << 16); }

template <class Epi>
__global__ void __launch_bounds__(kThreads, 1) gemm_tn_kernel(const __grid_constant__ CUtensorMap mapA,
                                                              const __grid_constant__ CUtensorMap mapB,
                                                              const __grid_constant__ CUtensorMap mapB2, int nsplitB, int M, int N,
                                                              int K, int kb_per_split, Epi epi, int fast, int a_batch) {
  // a_batch > 0: A is a rank-3 [T][a_batch][M] view (time-major OR batch-first activations: the two outer strides are free) and
  // contraction index k = t * a_batch + b; a_batch % 32 == 0, so a 32-row K tile never straddles a timestep.
  // Two B matrices side by side along N: columns [0, nsplitB) come from mapB, columns [nsplitB, N) from mapB2 (nsplitB % 32 == 0;
  // TMA zero-fills past each matrix's own width).  One pass over A then serves two products that share it (dBm and dVx both
  // contract dPre: at cfg5 that operand is 4.3 GB).  Single-matrix callers pass nsplitB >= N.
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;
  uint64_t* split = bars + kStages;
  uint64_t* empty = bars + 2 * kStages;
  uint64_t* accf = bars + 3 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);

  const int warp = warp_index(), lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int nkb_total = (K + BK - 1) / BK;
  const int kb0 = blockIdx.z * kb_per_split;
  const int nkb = (nkb_total - kb0) < kb_per_split ? (nkb_total - kb0) : kb_per_split;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&split[s], 4); mbar_init(&empty[s], 1); }
    mbar_init(accf, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kStages, it = kb / kStages;
      if (it > 0) mbar_wait(&empty[s], (it - 1) & 1);
      uint8_t* st = smem + s * kStageBytes;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[s], 2 * kTileBytes);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (a_batch > 0) {
            const int k0 = (kb0 + kb) * BK;
            tma_load_3d(st + i * 4096, &mapA, m0 + 32 * i, k0 % a_batch, k0 / a_batch, &full[s]);
          } else {
            tma_load_2d(st + i * 4096, &mapA, m0 + 32 * i, (kb0 + kb) * BK, &full[s]);
          }
          const int nb = n0 + 32 * i;
          if (nb < nsplitB) tma_load_2d(st + 2 * kTileBytes + i * 4096, &mapB, nb, (kb0 + kb) * BK, &full[s]);
          else tma_load_2d(st + 2 * kTileBytes + i * 4096, &mapB2, nb - nsplitB, (kb0 + kb) * BK, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_mn(BM, BN);
    int rr = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kStages, it = kb / kStages;
      mbar_wait(&split[s], it & 1);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * kStageBytes);
      // a k-step is the next group of 8 K rows: 1024 bytes = 64 descriptor units
      const uint64_t da_hi = make_desc_mn(a_addr), da_lo = make_desc_mn(a_addr + kTileBytes);
      const uint64_t db_hi = make_desc_mn(a_addr + 2 * kTileBytes), db_lo = make_desc_mn(a_addr + 3 * kTileBytes);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const int kk = kb * (BK / 8) + k;
          if (!fast) {
            mma_tf32_ss(tmem_d + 3 * BN, da_lo + 64 * k, db_hi + 64 * k, idesc, kk ? 1u : 0u);
            mma_tf32_ss(tmem_d + 3 * BN, da_hi + 64 * k, db_lo + 64 * k, idesc, 1u);
          }
          const int r3 = (rr + k) % 3;
          mma_tf32_ss(tmem_d + r3 * BN, da_hi + 64 * k, db_hi + 64 * k, idesc, kk >= 3 ? 1u : 0u);
        }
        mma_commit(&empty[s]);
      }
      rr = (rr + BK / 8) % 3;
    }
    if (elect_one()) mma_commit(accf);
  } else {
    const int sw = warp - 2;
    const int st_tid = sw * 32 + lane;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kStages, it = kb / kStages;
      mbar_wait(&full[s], it & 1);
      float4* a_hi = reinterpret_cast<float4*>(smem + s * kStageBytes);
      float4* a_lo = a_hi + kTileBytes / 16;
      float4* b_hi = a_hi + 2 * kTileBytes / 16;
      float4* b_lo = a_hi + 3 * kTileBytes / 16;
#pragma unroll 4
      for (int i = st_tid; i < (fast ? 0 : kTileBytes / 16); i += 128) {     // fast mode: the raw tile is the only operand
        const float4 va = a_hi[i], vb = b_hi[i];
#ifdef VMLMF_SPLIT_ROUND
        float4 ha, la, hb, lb;
        ha.x = tf32r(va.x); ha.y = tf32r(va.y); ha.z = tf32r(va.z); ha.w = tf32r(va.w);
        la.x = tf32r(va.x - ha.x); la.y = tf32r(va.y - ha.y); la.z = tf32r(va.z - ha.z); la.w = tf32r(va.w - ha.w);
        hb.x = tf32r(vb.x); hb.y = tf32r(vb.y); hb.z = tf32r(vb.z); hb.w = tf32r(vb.w);
        lb.x = tf32r(vb.x - hb.x); lb.y = tf32r(vb.y - hb.y); lb.z = tf32r(vb.z - hb.z); lb.w = tf32r(vb.w - hb.w);
        a_hi[i] = ha; a_lo[i] = la; b_hi[i] = hb; b_lo[i] = lb;
#else
        // The tensor core reads the top 19 bits of a tf32 operand and ignores the low 13 mantissa bits, so the raw fp32 tile
        // the TMA delivered IS its own (truncated) hi part: only the remainder is computed and written -- a third less
        // shared-memory traffic in the stage that bounds this kernel (same trick as the mma.sync kernels, seq_mma.cuh:split4).
        float4 la, lb;
        la.x = tf32_lo(va.x); la.y = tf32_lo(va.y); la.z = tf32_lo(va.z); la.w = tf32_lo(va.w);
        lb.x = tf32_lo(vb.x); lb.y = tf32_lo(vb.y); lb.z = tf32_lo(vb.z); lb.w = tf32_lo(vb.w);
        a_lo[i] = la; b_lo[i] = lb;
#endif
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&split[s]);
    }
    mbar_wait(accf, 0);
    tc_fence_after();
    asm volatile("bar.sync 1, 128;" ::: "memory");       // see gemm_tc_kernel: staging reuses the stages these warps wrote
    const int lane_grp = warp & 3;
    const uint32_t tbase = tmem_d + ((uint32_t)(lane_grp * 32) << 16);
    const int nks = nkb * (BK / 8);
    const int nhi = nks < 3 ? nks : 3;
    float* S = reinterpret_cast<float*>(smem);
    {
      float* srow = S + (size_t)(lane_grp * 32 + lane) * kEpiPitch;
#pragma unroll 1
      for (int c = 0; c < BN; c += 8) {
        float v[8];
        tmem_ld8(tbase + c, nhi, v, !fast);
        reinterpret_cast<float4*>(srow + c)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(srow + c)[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int r = sw; r < BM; r += 4) {
      const int m = m0 + r;
      if (m >= M) break;
      const float* srow = S + (size_t)r * kEpiPitch + lane;
      const float v[4] = {srow[0], srow[32], srow[64], srow[96]};
      epi(m, n0, N, lane, v);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols));
  }
}

// ------------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
constexpr int kTcNoFit = -1001;

// rows x K fp32 matrix, row pitch ld floats: box = 32 (K) x 128 rows
inline int make_map_2d(CUtensorMap* map, const float* p, long long rows, long long K, long long ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return kTcNoFit;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {BK, BM};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : kTcNoFit;
}
// [4][H][K] view of a [4H, K] factor (row pitch ld): box = 32 (K) x 32 units x 4 gates
inline int make_map_gate3d(CUtensorMap* map, const float* p, long long H, long long K, long long ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return kTcNoFit;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)H, 4};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)H * ld * 4};
  cuuint32_t box[3] = {BK, 32, 4};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : kTcNoFit;
}
inline bool tc_operand_ok(const float* p, long long ld) {
  return ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && (ld % 4 == 0) && ld > 0;
}

// C = A[M,K] B[N,K]^T with a plain epilogue.  Returns kTcNoFit when an operand does not meet the TMA constraints.
// k-blocks per split so that the grid has about one wave of CTAs (never more than `max_splits` splits)
inline int tc_splits(int M, int N, int K, int max_splits) {
  const long long tiles = (long long)ceil_div(M, BM) * ceil_div(N, BN);
  const int nkb = ceil_div(K, BK);
  long long s = (num_sms() + tiles - 1) / tiles;
  if (s > nkb / 4) s = nkb / 4;                      // keep at least 4 k-blocks per split
  if (s > max_splits) s = max_splits;
  if (s < 1) s = 1;
  return (int)s;
}

// VMLMF_FAST_TF32=1: OPTIONAL single-pass TF32 mode of the tcgen05 GEMMs (one MMA per k-step instead of the three of the
// fp32-accurate 3xTF32 scheme, no hi/lo split stage).  Products then carry ~2^-10 relative error per operand: results
// agree with the fp32 reference to ~1e-3 instead of 1e-5 (tests/test_gpu_tail.py states the bound and checks that
// classification decisions do not change).  Off by default; read per call, so vmlmf_b200.set_fast_tf32() can toggle it.
inline int fast_tf32() {
  const char* e = getenv("VMLMF_FAST_TF32");
  return (e && e[0] == '1') ? 1 : 0;
}

template <class Epi>
inline int gemm_tc(const float* A, long long lda, const float* Bm, long long ldb, int M, int N, int K, Epi epi,
                   cudaStream_t st, int splits = 1) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (!tc_operand_ok(A, lda) || !tc_operand_ok(Bm, ldb)) return kTcNoFit;
  CUtensorMap ma, mb;
  int rc = make_map_2d(&ma, A, M, K, lda);
  if (rc) return rc;
  rc = Epi::kGate ? make_map_gate3d(&mb, Bm, N / 4, K, ldb) : make_map_2d(&mb, Bm, N, K, ldb);
  if (rc) return rc;
  auto kern = gemm_tc_kernel<Epi>;
  static PerDevice attr_pd;                            // one flag per (epilogue instantiation, device)
  int& attr = attr_pd.cur();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = 1;
  }
  const int ntn = Epi::kGate ? ceil_div(N / 4, 32) : ceil_div(N, BN);
  const int nkb = ceil_div(K, BK);
  const int kbs = ceil_div(nkb, splits);
  dim3 grid(ntn, ceil_div(M, BM), ceil_div(nkb, kbs));
  kern<<<grid, kThreads, kSmemBytes, st>>>(ma, mb, M, N, K, kbs, epi, Epi::kGate ? 0 : fast_tf32());
  return (int)cudaGetLastError();
}


// [rows, cols] fp32 matrix, row pitch ld floats, viewed as an MN-major operand: box = 32 columns x 32 rows
inline int make_map_tn(CUtensorMap* map, const float* p, long long rows, long long cols, long long ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return kTcNoFit;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, BK};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : kTcNoFit;
}
// C[M,N] = At[K rows, M cols]^T Bt[K rows, N cols]   (row pitches lda, ldb floats; both 16-byte aligned rows)
template <class Epi>
inline int gemm_tn(const float* At, long long lda, const float* Bt, long long ldb, int M, int N, long long K, Epi epi,
                   cudaStream_t st, int splits = 1) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (!tc_operand_ok(At, lda) || !tc_operand_ok(Bt, ldb) || K > 0x7fffffffLL) return kTcNoFit;
  CUtensorMap ma, mb;
  int rc = make_map_tn(&ma, At, K, M, lda);
  if (rc) return rc;
  rc = make_map_tn(&mb, Bt, K, N, ldb);
  if (rc) return rc;
  auto kern = gemm_tn_kernel<Epi>;
  static PerDevice attr_pd;
  int& attr = attr_pd.cur();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = 1;
  }
  const int nkb = ceil_div((int)K, BK);
  const int kbs = ceil_div(nkb, splits);
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM), ceil_div(nkb, kbs));
  kern<<<grid, kThreads, kSmemBytes, st>>>(ma, mb, mb, 0x7fffffff, M, N, (int)K, kbs, epi, fast_tf32(), 0);
  return (int)cudaGetLastError();
}
// C[M, N] = A^T Bt where A is the rank-3 view a[t][b][m] = p[t * s_t + b * s_b + m] (t < Tn, b < Bsz, Bsz % 32 == 0) contracted over
// k = t * Bsz + b, and Bt is [Tn * Bsz, N] row-major: lets dA = Hprev^T dZ read y in place whatever its layout
template <class Epi>
inline int gemm_tn_a3(const float* p, long long s_b, long long s_t, int Bsz, int Tn, const float* Bt, long long ldb, int M, int N,
                      Epi epi, cudaStream_t st, int splits) {
  const long long K = (long long)Tn * Bsz;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  EncodeTiledFn fn = encode_fn();
  if (!fn || (Bsz & 31) || (s_b & 3) || (s_t & 3) || (reinterpret_cast<uintptr_t>(p) & 15) || !tc_operand_ok(Bt, ldb) || K > 0x7fffffffLL)
    return kTcNoFit;
  CUtensorMap ma, mb;
  {
    cuuint64_t dims[3] = {(cuuint64_t)M, (cuuint64_t)Bsz, (cuuint64_t)Tn};
    cuuint64_t strides[2] = {(cuuint64_t)s_b * 4, (cuuint64_t)s_t * 4};
    cuuint32_t box[3] = {32, BK, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return kTcNoFit;
  }
  int rc = make_map_tn(&mb, Bt, K, N, ldb);
  if (rc) return rc;
  auto kern = gemm_tn_kernel<Epi>;
  static PerDevice attr_pd;
  int& attr = attr_pd.cur();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = 1;
  }
  const int nkb = ceil_div((int)K, BK);
  const int kbs = ceil_div(nkb, splits);
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM), ceil_div(nkb, kbs));
  kern<<<grid, kThreads, kSmemBytes, st>>>(ma, mb, mb, 0x7fffffff, M, N, (int)K, kbs, epi, fast_tf32(), Bsz);
  return (int)cudaGetLastError();
}
// C[M, N1p + N2] = At^T [B1 | B2] with B1's columns padded to N1p = round_up(N1, 32): two products sharing the A operand
template <class Epi>
inline int gemm_tn2(const float* At, long long lda, const float* B1, long long ldb1, int N1, const float* B2, long long ldb2,
                    int N2, int M, long long K, Epi epi, cudaStream_t st, int splits) {
  if (M <= 0 || N1 <= 0 || N2 <= 0 || K <= 0) return 0;
  if (!tc_operand_ok(At, lda) || !tc_operand_ok(B1, ldb1) || !tc_operand_ok(B2, ldb2) || K > 0x7fffffffLL) return kTcNoFit;
  CUtensorMap ma, mb1, mb2;
  int rc = make_map_tn(&ma, At, K, M, lda);
  if (rc) return rc;
  rc = make_map_tn(&mb1, B1, K, N1, ldb1);
  if (rc) return rc;
  rc = make_map_tn(&mb2, B2, K, N2, ldb2);
  if (rc) return rc;
  auto kern = gemm_tn_kernel<Epi>;
  static PerDevice attr_pd;
  int& attr = attr_pd.cur();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = 1;
  }
  const int N1p = (N1 + 31) / 32 * 32, N = N1p + N2;
  const int nkb = ceil_div((int)K, BK);
  const int kbs = ceil_div(nkb, splits);
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM), ceil_div(nkb, kbs));
  kern<<<grid, kThreads, kSmemBytes, st>>>(ma, mb1, mb2, N1p, M, N, (int)K, kbs, epi, fast_tf32(), 0);
  return (int)cudaGetLastError();
}

}  // namespace tc
}  // namespace vmlmf
