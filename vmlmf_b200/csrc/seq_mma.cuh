// seq_mma.cuh -- regime R1, forward: persistent recurrence with the rank contractions on warp-level
// tensor-core MMAs (mma.sync m16n8k8, 3xTF32 error-compensated => fp32-accurate) and everything else
// (vector-multiplication terms, gate non-linearities, c/h update) in registers on the MMA accumulator
// fragments.
//
// Why mma.sync and not tcgen05 here: per step the contraction is  pre[16..128 seq, 4H] = [z|zx|1][.,16] W^T
// with K = RH+RX+1 <= 32 and  z[., RH<=16] = h[., H] A;  the outputs feed a strictly sequential element-
// wise update whose state (c, h) must stay in registers between steps.  tcgen05 accumulates in TMEM
// (4H = 1024 fp32 columns do not fit the 512-column TMEM at H=256) and needs its A operand in shared
// memory or TMEM, i.e. a TMEM->RF->SMEM round trip of h every timestep; with mma.sync the accumulator
// fragment of the gate GEMM *is* the A fragment of the next step's z GEMM (same lane, same registers),
// so h never leaves the register file.  The time-parallel GEMMs (regime G) are the tcgen05 kernels.
//
// Tiling: one CTA owns 16 sequences for all T steps (persistent over batch tiles); warp w owns two 8-unit halves,
// P = 0: units [8w, 8w+8) and P = 1: units [8NW + 8w, 8NW + 8w + 8) -- halves of one warp are NOT adjacent, so
// the units with an input-side term (j < I, the low indices) are spread over as many warps as possible instead of
// loading the first I/16 warps with all of the x-side work.  Lane (g = lane/4, q = lane%4) owns the two units
// half(P) + 2q + {0,1} of sequences g and g+8: 8 (sequence, unit) pairs whose c and h_{t-1} live in registers.
// Per step:
//   A fragments   rows [z | zx | 1 | 0] of the 16 sequences: z = sum over warps of last step's partial
//                 products (smem), zx from the time-parallel x projection, "1" carries the bias.
//   gate GEMM     per P and gate: 8 columns = gate k of units 32w+8P..+7, K = 8*KS   (3 MMAs per k-step)
//                 B fragments (Bm | Vx | bias, pre-scaled by -log2 e so ex2 needs no multiply) are split
//                 into tf32 hi/lo once per CTA and parked in lane-private shared memory.
//   epilogue      + Dh*h_{t-1} + Dx*x_t, sigmoid/tanh via ex2/rcp, c/h update, stores (float2 per lane).
//   z GEMM        h_t (accumulator layout == A-fragment layout under the unit permutation
//                 k-slot q <-> unit 8P+2q, q+4 <-> 8P+2q+1) times this warp's 32 rows of A -> partial z.
// One __syncthreads per step (partials are double buffered).
//
// Replaces (reference, "V/" = rnn_compression_factorization_vmlmf/src/): V/models/vmlmf.py:308-310 with the
// cell body :78-125; vmlmf_group.py:85-155; vmlmf_lm.py:272-280.
#pragma once
#include "seq_r1.cuh"

namespace vmlmf {

struct SeqFwdMmaArgs {
  SeqFwdArgs s;
  int zp, zxp;      // row pitch of saved z / of zx (floats)
};

// d += a(16x8, row) * b(8x8, col), tf32 inputs (fp32 bit patterns), fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], float b0, float b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])),
                 "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
// 3xTF32: small cross terms first
__device__ __forceinline__ void mma_3x(float (&d)[4], const float (&ahi)[4], const float (&alo)[4], float b0hi,
                                       float b1hi, float b0lo, float b1lo) {
  mma_tf32(d, alo, b0hi, b1hi);
  mma_tf32(d, ahi, b0lo, b1lo);
  mma_tf32(d, ahi, b0hi, b1hi);
}
// For sums that run over many timesteps / tiles: the tensor core adds into its fp32 accumulator with truncation (~3e-8
// relative per add, always toward zero), so a long-lived MMA accumulator drifts linearly with the number of adds.  Here the
// three products go into a fresh accumulator and join the running sum with FADDs (round to nearest).
__device__ __forceinline__ void mma_3x_rn(float (&sum)[4], const float (&ahi)[4], const float (&alo)[4], float b0hi,
                                          float b1hi, float b0lo, float b1lo) {
  float t[4] = {0.f, 0.f, 0.f, 0.f};
  mma_3x(t, ahi, alo, b0hi, b1hi, b0lo, b1lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) sum[i] += t[i];
}
__device__ __forceinline__ void split4(const float (&v)[4], float (&hi)[4], float (&lo)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#ifdef VMLMF_SPLIT_ROUND
    hi[i] = tf32_rna(v[i]);
    lo[i] = v[i] - hi[i];            // exact; the tensor core drops its low 13 bits (|error| <= 2^-21 |v|)
#else
    // The tensor core ignores the low 13 mantissa bits of a tf32 operand, so the raw value IS its own truncated hi
    // part: only the remainder costs instructions (one LOP, one FADD; exact).  |lo| < 2^-10 |v| instead of 2^-11 with
    // round-to-nearest, so the dropped lo*lo term is <= 2^-20 of a product (measured end to end in the parity tests).
    hi[i] = v[i];
    lo[i] = v[i] - __uint_as_float(__float_as_uint(v[i]) & 0xffffe000u);
#endif
  }
}

// asynchronous prefetch of a contiguous global range into L2 (bytes: multiple of 16, p: 16-byte aligned)
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

constexpr float kNegLog2e = -1.4426950408889634f;
constexpr float kNeg2Log2e = -2.8853900817779268f;

// Fragment-major layout of the tensors that only the warp-MMA kernels exchange (saved gates and c of the forward,
// dPre of the backward).  One block per (timestep t, 16-sequence tile): blk = t * ntiles + tile.  Inside a
// block, quantity Q of NQ, warp w (16 hidden units), P (8-unit half), hf (sequence half) hold 32 lanes x 2 floats:
// lane (g,q) = units 16w + 8P + 2q + {0,1} of sequence 16*tile + g + 8*hf  -- exactly the MMA accumulator
// fragment, so every warp-level load/store is one contiguous 256-byte segment.
__host__ __device__ inline size_t frag_addr(size_t blk, int NQ, int Q, int NW, int w, int P, int hf, int lane) {
  return ((((blk * NQ + Q) * NW + w) * 4 + P * 2 + hf) * 64) + lane * 2;
}
__host__ __device__ inline size_t frag_floats(int T, int B, int H, int NQ) {       // buffer size incl. padding
  return (size_t)T * ((B + 15) / 16) * NQ * ((H + 15) / 16) * 256;
}

__host__ __device__ constexpr int mma_pp(int NZ) { return 8 * NZ + 4; }    // pitch of a z-partial row
__host__ __device__ constexpr int mma_sp(int KS) { return 8 * KS + 4; }    // pitch of an A row

inline size_t seq_fwd_mma_smem_bytes(int NW, int KS, int NZ) {
  size_t fl = (size_t)NW * 8 * KS * 32 * 4             // B fragments (float4 per lane)
              + (size_t)NW * 16 * mma_pp(NZ)           // z partials of the warps
              + (size_t)16 * mma_sp(KS)                // A rows [z | zx | 1 | 0] of the 16 sequences
              + 2 * 4 * (size_t)NW * 16;               // Dh, Dx (scaled), [4][NW*16] each
  return fl * sizeof(float);
}

// KS: k-steps of the gate GEMM (8*KS >= RH+RX+1); NZ: n-tiles of the z GEMM (8*NZ >= RH)
// blockDim = 32 * NW, NW = ceil(H/16) <= 16: warp w owns units [16w, 16w+16).
template <int KS, int NZ, bool SAVE>
__global__ void __launch_bounds__(512, 1) seq_fwd_mma_kernel(const SeqFwdMmaArgs aa) {
  constexpr int PP = mma_pp(NZ), SP = mma_sp(KS);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int H = aa.s.H, I = aa.s.I, B = aa.s.B, T = aa.s.T, RH = aa.s.RH, RX = aa.s.RX;
  const int HP = NW * 16;
  const int ubase = warp * 8;                            // first unit of this warp's half P = 0
  const int PS = 8 * NW;                                 // half P = 1 holds units PS + 8 warp .. (see the tiling note above)
  const int zp = aa.zp, zxp = aa.zxp;

  extern __shared__ __align__(16) float smem[];
  float4* Bf = reinterpret_cast<float4*>(smem);          // [NW][2 P][4 k][KS][32]
  float* Pz = smem + (size_t)NW * 8 * KS * 32 * 4;       // [NW][16][PP]
  float* Ar = Pz + (size_t)NW * 16 * PP;                 // [16][SP]
  float* DhS = Ar + 16 * SP;                             // [4][HP]
  float* DxS = DhS + 4 * HP;                             // [4][HP]

  // ---------------- per-CTA prologue: weights -> fragments ----------------
  // concatenated, pre-scaled row of gate k / unit j:  [Bm(RH) | Vx(RX) | bias | 0..]
  auto wcat = [&](int k, int j, int slot) -> float {
    if (j >= H) return 0.f;
    const float sc = (k == 3) ? kNeg2Log2e : kNegLog2e;
    const size_t row = (size_t)k * H + j;
    float v = 0.f;
    if (slot < RH) v = __ldg(aa.s.Bm + row * RH + slot);
    else if (slot < RH + RX) v = __ldg(aa.s.Vx + row * RX + (slot - RH));
    else if (slot == RH + RX) v = __ldg(aa.s.bias + row);
    return v * sc;
  };
  const float4* myB = Bf + (size_t)warp * 8 * KS * 32 + lane;
#pragma unroll 1
  for (int pk = 0; pk < 8; ++pk) {
    const int P = pk >> 2, k = pk & 3;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const float b0 = wcat(k, ubase + PS * P + g, 8 * s + q), b1 = wcat(k, ubase + PS * P + g, 8 * s + q + 4);
      const float b0h = tf32_rna(b0), b1h = tf32_rna(b1);
      const_cast<float4*>(myB)[(pk * KS + s) * 32] = make_float4(b0h, b1h, tf32_rna(b0 - b0h), tf32_rna(b1 - b1h));
    }
  }
  for (int i = tid; i < 4 * HP; i += blockDim.x) {
    const int k = i / HP, j = i - k * HP;
    const float sc = (k == 3) ? kNeg2Log2e : kNegLog2e;
    DhS[i] = (j < H) ? __ldg(aa.s.Dh + k * H + j) * sc : 0.f;
    DxS[i] = (j < I) ? __ldg(aa.s.Dx + k * I + j) * sc : 0.f;
  }
  // this warp's rows of A as B fragments of the z GEMM (unit permutation: slot q -> 8P+2q, q+4 -> 8P+2q+1)
  float Azh[2][NZ][2], Azl[2][NZ][2];
#pragma unroll
  for (int P = 0; P < 2; ++P)
#pragma unroll
    for (int nz = 0; nz < NZ; ++nz)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = ubase + PS * P + 2 * q + e, r = 8 * nz + g;
        const float v = (j < H && r < RH) ? __ldg(aa.s.A + (size_t)j * RH + r) : 0.f;
        Azh[P][nz][e] = tf32_rna(v);
        Azl[P][nz][e] = tf32_rna(v - Azh[P][nz][e]);
      }
  // A rows: zero, 1 in the bias slot; z and zx slots are rewritten every step
  for (int i = tid; i < 16 * SP; i += blockDim.x) Ar[i] = ((i % SP) == RH + RX) ? 1.f : 0.f;

  const bool xwarp = ubase < I;                          // this warp has units with an x term
  const int ntiles = ceil_div(B, 16);
  const int j0 = ubase + 2 * q;                          // this lane's units: j0 + PS*P + e
  const int nthreads = blockDim.x;
  const bool wy = aa.s.y != nullptr;                    // y == nullptr: the caller consumes only (hT, cT)

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0 = tile * 16;
    const int sq[2] = {b0 + g, b0 + g + 8};
    const bool ok[2] = {sq[0] < B, sq[1] < B};
    // per-lane row pointers (advanced by one timestep at the end of every step)
    float* yrow[2]; const float* xrow[2];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      yrow[hf] = aa.s.y + (size_t)sq[hf] * aa.s.ys_b + j0;
      xrow[hf] = aa.s.x + (size_t)sq[hf] * aa.s.xs_b + j0;
    }
    // fragment-major saved state (frag_addr): block = (t, tile); this lane's slot of quantity 0, P = 0, hf = 0
    float* gfrag = SAVE ? aa.s.gates + frag_addr((size_t)tile, 4, 0, NW, warp, 0, 0, lane) : nullptr;
    float* cfrag = SAVE ? aa.s.cs + frag_addr((size_t)tile, 1, 0, NW, warp, 0, 0, lane) : nullptr;
    const size_t gstep = (size_t)ntiles * 4 * NW * 256, cstep = (size_t)ntiles * NW * 256;
    float c[2][2][2], hp[2][2][2], xn[2][2][2];          // [P][e][hf]
#pragma unroll
    for (int P = 0; P < 2; ++P)
#pragma unroll
      for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int j = j0 + PS * P + e;
          const bool ld = ok[hf] && j < H;
          hp[P][e][hf] = (ld && aa.s.h0) ? aa.s.h0[(size_t)sq[hf] * H + j] : 0.f;
          c[P][e][hf] = (ld && aa.s.c0) ? aa.s.c0[(size_t)sq[hf] * H + j] : 0.f;
          xn[P][e][hf] = (ok[hf] && j < I) ? xrow[hf][PS * P + e] : 0.f;                 // t = 0
        }
    // zx staging (warp NW-1): lane -> (sequence lane/2, half of each 8-column group)
    const int zsb = lane >> 1, zhh = lane & 1;
    float zxr[KS][4];
    auto fetch_zx = [&](int t) {
#pragma unroll
      for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = 8 * s + 4 * zhh + i;
          zxr[s][i] = (r < RX && (b0 + zsb) < B) ? aa.s.zx[((size_t)t * B + b0 + zsb) * zxp + r] : 0.f;
        }
    };
    auto stage_zx = [&]() {
      float* row = Ar + zsb * SP + RH;
#pragma unroll
      for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = 8 * s + 4 * zhh + i;
          if (r < RX) row[r] = zxr[s][i];
        }
    };
    auto z_partial_store = [&](const float (&zacc)[NZ][4]) {
      float* pw = Pz + (size_t)warp * 16 * PP;
#pragma unroll
      for (int nz = 0; nz < NZ; ++nz) {
        *reinterpret_cast<float2*>(pw + g * PP + 8 * nz + 2 * q) = make_float2(zacc[nz][0], zacc[nz][1]);
        *reinterpret_cast<float2*>(pw + (g + 8) * PP + 8 * nz + 2 * q) = make_float2(zacc[nz][2], zacc[nz][3]);
      }
    };
    // fixed-order sum of the warps' partials -> z slots of the A rows (and the saved z of step tz)
    auto z_reduce = [&](int tz) {
      for (int idx = tid; idx < 16 * 8 * NZ * 4; idx += nthreads) {
        const int el = idx >> 2, part = idx & 3, seq = el / (8 * NZ), slot = el - seq * (8 * NZ);
        float s = 0.f;
        for (int w = part; w < NW; w += 4) s += Pz[((size_t)w * 16 + seq) * PP + slot];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0) {
          if (slot < RH) Ar[seq * SP + slot] = s;
          if (SAVE && slot < zp && (b0 + seq) < B)
            aa.s.z[((size_t)tz * B + b0 + seq) * zp + slot] = slot < RH ? s : 0.f;
        }
      }
    };
    auto z_mma = [&](float (&zacc)[NZ][4], int P, const float (&hv)[4]) {
      float hh[4], hl[4];
      split4(hv, hh, hl);
#pragma unroll
      for (int nz = 0; nz < NZ; ++nz)
        mma_3x(zacc[nz], hh, hl, Azh[P][nz][0], Azh[P][nz][1], Azl[P][nz][0], Azl[P][nz][1]);
    };

    // ---- z_0 = h0 A, zx_0 ----
    {
      float zacc[NZ][4];
#pragma unroll
      for (int nz = 0; nz < NZ; ++nz)
#pragma unroll
        for (int i = 0; i < 4; ++i) zacc[nz][i] = 0.f;
      if (aa.s.h0) {
#pragma unroll
        for (int P = 0; P < 2; ++P) {
          const float hv[4] = {hp[P][0][0], hp[P][0][1], hp[P][1][0], hp[P][1][1]};
          z_mma(zacc, P, hv);
        }
      }
      __syncthreads();                                   // prologue visible / previous tile finished with Ar, Pz
      z_partial_store(zacc);
      if (warp == NW - 1) { fetch_zx(0); stage_zx(); }
      __syncthreads();
      z_reduce(0);
      __syncthreads();
    }

    for (int t = 0; t < T; ++t) {
      const bool more = t + 1 < T;
      // ---- A fragments of the gate GEMM ----
      float ahi[KS][4], alo[KS][4];
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        float v[4];
        v[0] = Ar[g * SP + 8 * s + q];
        v[1] = Ar[(g + 8) * SP + 8 * s + q];
        v[2] = Ar[g * SP + 8 * s + q + 4];
        v[3] = Ar[(g + 8) * SP + 8 * s + q + 4];
        split4(v, ahi[s], alo[s]);
      }
      // ---- prefetch step t+1 inputs ----
      float xv[2][2][2];
#pragma unroll
      for (int P = 0; P < 2; ++P)
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) xv[P][e][hf] = xn[P][e][hf];
      if (more) {
        if (warp == NW - 1) fetch_zx(t + 1);
        if (xwarp) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            xrow[hf] += aa.s.xs_t;
#pragma unroll
            for (int P = 0; P < 2; ++P)
#pragma unroll
              for (int e = 0; e < 2; ++e)
                xn[P][e][hf] = (ok[hf] && (j0 + PS * P + e) < I) ? xrow[hf][PS * P + e] : 0.f;
          }
        }
      }

      float zacc[NZ][4];
#pragma unroll
      for (int nz = 0; nz < NZ; ++nz)
#pragma unroll
        for (int i = 0; i < 4; ++i) zacc[nz][i] = 0.f;

#pragma unroll
      for (int P = 0; P < 2; ++P) {
        // ---- gate GEMM for units ubase+PS*P .. +7: acc[k][i], i = 2*hf + e ----
        float acc[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[k][i] = 0.f;
#pragma unroll
        for (int s = 0; s < KS; ++s) {                  // 4 independent accumulators per MMA round
          float4 b[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) b[k] = myB[((P * 4 + k) * KS + s) * 32];
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_tf32(acc[k], alo[s], b[k].x, b[k].y);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_tf32(acc[k], ahi[s], b[k].z, b[k].w);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_tf32(acc[k], ahi[s], b[k].x, b[k].y);
        }
        // ---- epilogue on the accumulator fragments ----
        float2 dh[4], dx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          dh[k] = *reinterpret_cast<const float2*>(DhS + k * HP + j0 + PS * P);
          dx[k] = *reinterpret_cast<const float2*>(DxS + k * HP + j0 + PS * P);
        }
        float hnew[2][2];                                // [e][hf]
        float gsave[4][2][2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int i = 2 * hf + e;
            const float hprev = hp[P][e][hf], xx = xv[P][e][hf];
            float pre[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float dhk = e ? dh[k].y : dh[k].x, dxk = e ? dx[k].y : dx[k].x;
              pre[k] = fmaf(dxk, xx, fmaf(dhk, hprev, acc[k][i]));     // = -log2e * pre-activation (x2 for n)
            }
            const float gi = rcp_approx(1.f + ex2_approx(pre[0]));
            const float gf = rcp_approx(1.f + ex2_approx(pre[1]));
            const float go = rcp_approx(1.f + ex2_approx(pre[2]));
            const float gn = fmaf(2.f, rcp_approx(1.f + ex2_approx(pre[3])), -1.f);
            const float cn = fmaf(gf, c[P][e][hf], gi * gn);
            const float tc = fmaf(2.f, rcp_approx(1.f + ex2_approx(cn * kNeg2Log2e)), -1.f);
            const float hn = go * tc;
            c[P][e][hf] = cn;
            hp[P][e][hf] = hn;
            hnew[e][hf] = hn;
            if (SAVE) { gsave[0][e][hf] = gi; gsave[1][e][hf] = gf; gsave[2][e][hf] = go; gsave[3][e][hf] = gn; }
          }
        // ---- stores: y in the caller's layout (8 rows x 32 B per instruction); saved gates / c in the
        //      fragment-major layout (each warp instruction writes 256 contiguous bytes) ----
        if (wy && ok[0] && (j0 + PS * P) < H) *reinterpret_cast<float2*>(yrow[0] + PS * P) = make_float2(hnew[0][0], hnew[1][0]);
        if (wy && ok[1] && (j0 + PS * P) < H) *reinterpret_cast<float2*>(yrow[1] + PS * P) = make_float2(hnew[0][1], hnew[1][1]);
        if (SAVE) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            *reinterpret_cast<float2*>(cfrag + (P * 2 + hf) * 64) = make_float2(c[P][0][hf], c[P][1][hf]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              *reinterpret_cast<float2*>(gfrag + (size_t)k * NW * 256 + (P * 2 + hf) * 64) =
                  make_float2(gsave[k][0][hf], gsave[k][1][hf]);
          }
        }
        // ---- z GEMM k-step P: A fragment = (h[g][u0], h[g+8][u0], h[g][u1], h[g+8][u1]) ----
        if (more) {
          const float hv[4] = {hnew[0][0], hnew[0][1], hnew[1][0], hnew[1][1]};
          z_mma(zacc, P, hv);
        }
      }
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        yrow[hf] += aa.s.ys_t;
      }
      if (SAVE) { gfrag += gstep; cfrag += cstep; }
      if (more) {
        z_partial_store(zacc);
        __syncthreads();                                 // partials complete; every warp is done reading Ar
        if (warp == NW - 1) stage_zx();
        z_reduce(t + 1);
        __syncthreads();
      }
    }
    // ---- final state ----
#pragma unroll
    for (int P = 0; P < 2; ++P)
#pragma unroll
      for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int j = j0 + PS * P + e;
          if (ok[hf] && j < H) {
            aa.s.hT[(size_t)sq[hf] * H + j] = hp[P][e][hf];
            aa.s.cT[(size_t)sq[hf] * H + j] = c[P][e][hf];
          }
        }
  }
}

// host launcher: returns kMmaNoFit when the shape is outside this kernel (caller falls back to the SIMT R1 kernel)
constexpr int kMmaNoFit = -1000;
int launch_fwd_mma(const SeqFwdMmaArgs& a, bool save, cudaStream_t st);
bool fwd_mma_fits(int I, int H, int RX, int RH);

}  // namespace vmlmf
