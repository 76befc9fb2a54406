// seq_bwd_mma.cuh -- regime R1, backward.  Two kernels:
//
//  K3a  seq_bwd_mma_kernel   reverse-time recurrence.  Same tiling as the forward (seq_mma.cuh): one CTA =
//       16 sequences, warp w = hidden units [16w,16w+16), lane (g,q) = units 16w+4q..+3 of sequences g, g+8.
//       Per step: dPre (gate-gradient algebra on registers), dzc = dPre [Bm|Vx]  (mma.sync 3xTF32, the
//       dPre registers are the A fragments; partial sums of the warps are added in a fixed order through
//       shared memory), dh_{t-1} = dz A^T + sum_k dPre_k Dh_k (second MMA, accumulator layout == state
//       layout).  Writes dPre[T*B,4,H] and dzc[T*B,8KS] (= [dz | dzx]) for K3b, and dh0 / dc0.
//  K3b  grad_rows_kernel     time-parallel: streams dPre once and accumulates every parameter gradient with
//       K = rows contractions on mma.sync (dBm|dVx|dbias = [z|zx|1]^T dPre,  dA = Hprev^T dz,
//       dUx = X^T dzx), the vector-multiplication gradients dDh, dDx on the same loaded fragments, and
//       dX = dzx Ux^T + sum_k dPre_k Dx_k.  Each CTA owns a contiguous range of rows and writes one partial;
//       reduce_partials_kernel adds them in a fixed order (bit-reproducible).
//
// No [H,4H] matrix or its gradient is formed.  Replaces the autograd replay of V/models/vmlmf.py:78-125
// inside the loop :308-310 (SURVEY Appendix A.3 is the algebra).
#pragma once
#include "seq_mma.cuh"

namespace vmlmf {

struct SeqBwdMmaArgs {
  const float *gates, *cs, *c0;
  const float* dy; long long dys_t, dys_b;
  const float *dhT, *dcT;
  const float *Vx, *A, *Bm, *Dh;
  float *dpre, *dzc, *dh0, *dc0;
  int T, B, H, RX, RH;
};

__host__ __device__ constexpr int bwd_pitch(int KS) { return 8 * KS + 4; }

inline size_t seq_bwd_mma_smem_bytes(int NW, int KS) {
  size_t fl = (size_t)NW * 8 * KS * 32 * 4             // B fragments of [Bm|Vx] (float4 {hi,hi,lo,lo} per lane)
              + (size_t)NW * 16 * bwd_pitch(KS)        // dzc partials of the warps
              + (size_t)16 * bwd_pitch(KS)             // reduced dzc rows
              + 4 * (size_t)NW * 16;                   // Dh [4][HP]
  return fl * sizeof(float);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float pick4(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// unit (within the warp's 16) of column c of n-tile / k-step P:  lane q then owns 4q..4q+3 across P = 0,1
__host__ __device__ constexpr int unit_of(int P, int c) { return 4 * (c >> 1) + 2 * P + (c & 1); }

template <int KS, int NZ>
__global__ void __launch_bounds__(512, 1) seq_bwd_mma_kernel(const SeqBwdMmaArgs a) {
  constexpr int PP = bwd_pitch(KS);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int H = a.H, B = a.B, T = a.T, RH = a.RH, RX = a.RX;
  const int HP = NW * 16;
  const int ubase = warp * 8;                            // half P = 0: units 8w..; half P = 1: units PS + 8w.. (as the forward)
  const int PS = 8 * NW;
  const int j0 = ubase + 2 * q;                          // this lane's units: j0 + PS*P + e  (same as the forward)

  extern __shared__ __align__(16) float smem[];
  float4* Bf = reinterpret_cast<float4*>(smem);          // [NW][2 P][4 k][KS][32]
  float* Pz = smem + (size_t)NW * 8 * KS * 32 * 4;       // [NW][16][PP]
  float* Dz = Pz + (size_t)NW * 16 * PP;                 // [16][PP]
  float* DhS = Dz + 16 * PP;                             // [4][HP]

  // ---- prologue: [Bm | Vx] rows of this warp's columns as B fragments of the dzc GEMM ----
  auto wc = [&](int k, int j, int slot) -> float {
    if (j >= H) return 0.f;
    const size_t row = (size_t)k * H + j;
    if (slot < RH) return __ldg(a.Bm + row * RH + slot);
    if (slot < RH + RX) return __ldg(a.Vx + row * RX + (slot - RH));
    return 0.f;
  };
  float4* myB = Bf + (size_t)warp * 8 * KS * 32 + lane;
#pragma unroll 1
  for (int pk = 0; pk < 8; ++pk) {
    const int P = pk >> 2, k = pk & 3;
#pragma unroll
    for (int s = 0; s < KS; ++s) {                       // k-slot q <-> unit 8P+2q, q+4 <-> 8P+2q+1; n = g <-> slot 8s+g
      const float b0 = wc(k, j0 + PS * P, 8 * s + g), b1 = wc(k, j0 + PS * P + 1, 8 * s + g);
      const float b0h = tf32_rna(b0), b1h = tf32_rna(b1);
      myB[(pk * KS + s) * 32] = make_float4(b0h, b1h, tf32_rna(b0 - b0h), tf32_rna(b1 - b1h));
    }
  }
  for (int i = tid; i < 4 * HP; i += blockDim.x) {
    const int k = i / HP, j = i - k * HP;
    DhS[i] = (j < H) ? __ldg(a.Dh + k * H + j) : 0.f;
  }
  // A^T as B fragments of the dh GEMM: k-slot = rank 8s+q (+4), n = g <-> unit 8P+g
  float Ath[2][NZ][2], Atl[2][NZ][2];
#pragma unroll
  for (int P = 0; P < 2; ++P)
#pragma unroll
    for (int s = 0; s < NZ; ++s)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = ubase + PS * P + g, r = 8 * s + q + 4 * e;
        const float v = (j < H && r < RH) ? __ldg(a.A + (size_t)j * RH + r) : 0.f;
        Ath[P][s][e] = tf32_rna(v);
        Atl[P][s][e] = tf32_rna(v - Ath[P][s][e]);
      }

  const int ntiles = ceil_div(B, 16);
  const int nthreads = blockDim.x;
  const size_t gstep = (size_t)ntiles * 4 * NW * 256, cstep = (size_t)ntiles * NW * 256, qstride = (size_t)NW * 256;
  const bool dy_vec = a.dy && ((reinterpret_cast<uintptr_t>(a.dy) & 7) == 0) && !(a.dys_t & 1) && !(a.dys_b & 1);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0 = tile * 16;
    const int sq[2] = {b0 + g, b0 + g + 8};
    bool ok[2][2];                                       // [hf][P]: sequence and unit pair exist
#pragma unroll
    for (int hf = 0; hf < 2; ++hf)
#pragma unroll
      for (int P = 0; P < 2; ++P) ok[hf][P] = sq[hf] < B && (j0 + PS * P) < H;
    float dhn[2][2][2], dcn[2][2][2];                    // [P][e][hf]
#pragma unroll
    for (int P = 0; P < 2; ++P)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float2 v = make_float2(0.f, 0.f), w = v;
        if (ok[hf][P] && a.dhT) v = __ldg(reinterpret_cast<const float2*>(a.dhT + (size_t)sq[hf] * H + j0 + PS * P));
        if (ok[hf][P] && a.dcT) w = __ldg(reinterpret_cast<const float2*>(a.dcT + (size_t)sq[hf] * H + j0 + PS * P));
        dhn[P][0][hf] = v.x; dhn[P][1][hf] = v.y;
        dcn[P][0][hf] = w.x; dcn[P][1][hf] = w.y;
      }
    // fragment-major pointers of the LAST timestep (walked backwards)
    const float* gfrag = a.gates + frag_addr((size_t)(T - 1) * ntiles + tile, 4, 0, NW, warp, 0, 0, lane);
    const float* cfrag = a.cs + frag_addr((size_t)(T - 1) * ntiles + tile, 1, 0, NW, warp, 0, 0, lane);
    float* dfrag = a.dpre + frag_addr((size_t)(T - 1) * ntiles + tile, 4, 0, NW, warp, 0, 0, lane);
    // L2 prefetch of the saved blocks two steps ahead (one thread, one bulk request per contiguous block)
    auto prefetch_step = [&](int tp) {
      if (tid != 0 || tp < 0) return;
      l2_prefetch_bulk(a.gates + frag_addr((size_t)tp * ntiles + tile, 4, 0, NW, 0, 0, 0, 0), (uint32_t)(4 * NW * 256 * sizeof(float)));
      if (tp > 0) l2_prefetch_bulk(a.cs + frag_addr((size_t)(tp - 1) * ntiles + tile, 1, 0, NW, 0, 0, 0, 0), (uint32_t)(NW * 256 * sizeof(float)));
    };
    if (tid == 0) l2_prefetch_bulk(a.cs + frag_addr((size_t)(T - 1) * ntiles + tile, 1, 0, NW, 0, 0, 0, 0), (uint32_t)(NW * 256 * sizeof(float)));
    prefetch_step(T - 1);
    prefetch_step(T - 2);
    __syncthreads();                                     // prologue visible; previous tile finished with Pz / Dz

    for (int t = T - 1; t >= 0; --t) {
      prefetch_step(t - 2);
      float dz[KS][4];
#pragma unroll
      for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) dz[s][i] = 0.f;
      float sd[2][2][2];                                 // [P][e][hf]  sum_k dPre_k Dh_k
#pragma unroll
      for (int P = 0; P < 2; ++P) {
        float dpre[4][2][2];                             // [k][e][hf]
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int o = (P * 2 + hf) * 64;
          float2 G[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) G[k] = __ldg(reinterpret_cast<const float2*>(gfrag + (size_t)k * qstride + o));
          const float2 ct2 = __ldg(reinterpret_cast<const float2*>(cfrag + o));
          float2 cp2 = make_float2(0.f, 0.f), dy2 = cp2;
          if (t > 0) cp2 = __ldg(reinterpret_cast<const float2*>(cfrag - cstep + o));
          else if (ok[hf][P] && a.c0) cp2 = __ldg(reinterpret_cast<const float2*>(a.c0 + (size_t)sq[hf] * H + j0 + PS * P));
          if (ok[hf][P] && a.dy) {
            const float* dp = a.dy + (size_t)t * a.dys_t + (size_t)sq[hf] * a.dys_b + j0 + PS * P;
            if (dy_vec) dy2 = __ldg(reinterpret_cast<const float2*>(dp));
            else { dy2.x = __ldg(dp); dy2.y = __ldg(dp + 1); }
          }
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float gi = e ? G[0].y : G[0].x, gf = e ? G[1].y : G[1].x, go = e ? G[2].y : G[2].x, gn = e ? G[3].y : G[3].x;
            const float ct = e ? ct2.y : ct2.x, cp = e ? cp2.y : cp2.x, dyv = e ? dy2.y : dy2.x;
            const float dh = dhn[P][e][hf] + dyv;
            const float tc = fmaf(2.f, rcp_approx(1.f + ex2_approx(ct * kNeg2Log2e)), -1.f);
            const float dc = fmaf(dh * go, fmaf(-tc, tc, 1.f), dcn[P][e][hf]);
            dpre[0][e][hf] = dc * gn * gi * (1.f - gi);
            dpre[1][e][hf] = dc * cp * gf * (1.f - gf);
            dpre[2][e][hf] = dh * tc * go * (1.f - go);
            dpre[3][e][hf] = dc * gi * fmaf(-gn, gn, 1.f);
            dcn[P][e][hf] = dc * gf;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<float2*>(dfrag + (size_t)k * qstride + o) = make_float2(dpre[k][0][hf], dpre[k][1][hf]);
        }
        // vector-multiplication part of dh_{t-1} and the partial dzc GEMM of this 8-unit half
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) sd[P][e][hf] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 d2 = *reinterpret_cast<const float2*>(DhS + k * HP + j0 + PS * P);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            sd[P][0][hf] = fmaf(dpre[k][0][hf], d2.x, sd[P][0][hf]);
            sd[P][1][hf] = fmaf(dpre[k][1][hf], d2.y, sd[P][1][hf]);
          }
          const float av[4] = {dpre[k][0][0], dpre[k][0][1], dpre[k][1][0], dpre[k][1][1]};
          float ah[4], al[4];
          split4(av, ah, al);
#pragma unroll
          for (int s = 0; s < KS; ++s) {
            const float4 b = myB[((P * 4 + k) * KS + s) * 32];
            mma_3x(dz[s], ah, al, b.x, b.y, b.z, b.w);
          }
        }
      }
      {
        float* pw = Pz + (size_t)warp * 16 * PP;
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          *reinterpret_cast<float2*>(pw + g * PP + 8 * s + 2 * q) = make_float2(dz[s][0], dz[s][1]);
          *reinterpret_cast<float2*>(pw + (g + 8) * PP + 8 * s + 2 * q) = make_float2(dz[s][2], dz[s][3]);
        }
      }
      gfrag -= gstep; cfrag -= cstep; dfrag -= gstep;
      __syncthreads();
      // ---- fixed-order sum over warps -> Dz rows (and the global dzc rows K3b reads) ----
      for (int idx = tid; idx < 16 * 8 * KS * 4; idx += nthreads) {
        const int el = idx >> 2, part = idx & 3, seq = el / (8 * KS), slot = el - seq * (8 * KS);
        float s = 0.f;
        for (int w = part; w < NW; w += 4) s += Pz[((size_t)w * 16 + seq) * PP + slot];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0) {
          Dz[seq * PP + slot] = s;
          if ((b0 + seq) < B) a.dzc[((size_t)t * B + b0 + seq) * (8 * KS) + slot] = s;
        }
      }
      __syncthreads();
      // ---- dh_{t-1} = dz A^T + sum_k dPre_k Dh_k ----
      float acc[2][4];
#pragma unroll
      for (int P = 0; P < 2; ++P) {
        acc[P][0] = sd[P][0][0]; acc[P][1] = sd[P][1][0]; acc[P][2] = sd[P][0][1]; acc[P][3] = sd[P][1][1];
      }
#pragma unroll
      for (int s = 0; s < NZ; ++s) {
        const float av[4] = {Dz[g * PP + 8 * s + q], Dz[(g + 8) * PP + 8 * s + q], Dz[g * PP + 8 * s + q + 4],
                             Dz[(g + 8) * PP + 8 * s + q + 4]};
        float ah[4], al[4];
        split4(av, ah, al);
#pragma unroll
        for (int P = 0; P < 2; ++P) mma_3x(acc[P], ah, al, Ath[P][s][0], Ath[P][s][1], Atl[P][s][0], Atl[P][s][1]);
      }
#pragma unroll
      for (int P = 0; P < 2; ++P) {
        dhn[P][0][0] = acc[P][0]; dhn[P][1][0] = acc[P][1]; dhn[P][0][1] = acc[P][2]; dhn[P][1][1] = acc[P][3];
      }
    }
#pragma unroll
    for (int P = 0; P < 2; ++P)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
        if (ok[hf][P]) {
          if (a.dh0) *reinterpret_cast<float2*>(a.dh0 + (size_t)sq[hf] * H + j0 + PS * P) = make_float2(dhn[P][0][hf], dhn[P][1][hf]);
          if (a.dc0) *reinterpret_cast<float2*>(a.dc0 + (size_t)sq[hf] * H + j0 + PS * P) = make_float2(dcn[P][0][hf], dcn[P][1][hf]);
        }
  }
}

// ------------------------------------------------------------------------------------------------- //
// K3b
// ------------------------------------------------------------------------------------------------- //
struct GradRowsArgs {
  const float* dpre;                       // fragment-major [T*ntiles][4][NW][2][2][32][2]
  const float* dzc;                        // [T*B, 8*KS]   [dz (RH) | dzx (RX) | 0]
  const float *z, *zx; int zp, zxp;        // saved z / zx rows
  const float* y; long long ys_t, ys_b;    // h_t for every t (h_{t-1} of row (t,b) is y[t-1,b], or h0 at t = 0)
  const float* h0;
  const float* x; long long xs_t, xs_b;
  const float *Ux, *Dx;
  float* dx; long long dxs_t, dxs_b;       // may be null
  float* partial;                          // [gridDim.x, GradLayout.total]
  int T, B, I, H, RX, RH;
  int blocks_per_cta;                      // (timestep, tile) blocks per CTA
};

inline size_t grad_rows_smem_bytes(int KS, int I, int RX) {
  (void)KS;
  return ((size_t)I * RX + 4 * (size_t)I) * sizeof(float);      // Ux, Dx
}

// One block = the 16 sequences of one tile at one timestep (the unit K3a writes dPre in).  blockDim = 32*NW,
// NW = ceil(H/16): warp w owns hidden units [16w,16w+16) for all four gates = 64 dPre columns, read as 32
// two-unit pieces (gate k, half P, pair qq): lane (g,q) loads piece (k = J, P = g>>2, qq = g&3) of rows q, q+4
// (+8, +12) -- one contiguous 128-byte segment per (k, P) in the fragment-major layout.
// MT = ceil(8*KS / 16) m-tiles of the [z|zx|1]^T dPre product.
template <int KS, int NT_MAX>
__global__ void __launch_bounds__(NT_MAX, 1) grad_rows_kernel(const GradRowsArgs a) {
  constexpr int MT = (8 * KS + 15) / 16;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int H = a.H, I = a.I, B = a.B, RH = a.RH, RX = a.RX;
  const int ntiles = ceil_div(B, 16);
  const long long nblocks = (long long)a.T * ntiles;
  const int Pg = g >> 2, qq = g & 3;
  const int PS = 8 * NW;                                 // unit offset of half P = 1 (same mapping as the recurrence kernels)
  const int ju = warp * 8 + PS * Pg + 2 * qq;            // this lane's unit pair ju, ju+1 (of every gate)
  const bool uin = ju < H;                               // H % 4 == 0
  const bool xcols = warp * 8 < I;                      // this warp's units overlap the input width

  extern __shared__ __align__(16) float smem[];
  float* UxS = smem;                                     // [I][RX]
  float* DxS = UxS + I * RX;                             // [4][I]
  for (int i = tid; i < I * RX; i += blockDim.x) UxS[i] = __ldg(a.Ux + i);
  for (int i = tid; i < 4 * I; i += blockDim.x) DxS[i] = __ldg(a.Dx + i);

  float accW[MT][8][4];                                  // [z|zx|1]^T dPre: n-tile (k, X): n = g <-> unit 16w + 8(g>>2) + 2(g&3) + X of gate k
  float accG[KS][4];                                     // Hprev^T dzc: m = g <-> unit pair element X=0, m = g+8 <-> X=1
  float accU[KS][4];                                     // X^T dzc
  float gDh[4][2], gDx[4][2];                            // [k][X] column sums of dPre*hprev, dPre*x
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) accW[m][n][i] = 0.f;
#pragma unroll
  for (int s = 0; s < KS; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i) accG[s][i] = accU[s][i] = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) gDh[k][0] = gDh[k][1] = gDx[k][0] = gDx[k][1] = 0.f;

  const long long blk_begin = (long long)blockIdx.x * a.blocks_per_cta;
  long long blk_end = blk_begin + a.blocks_per_cta;
  if (blk_end > nblocks) blk_end = nblocks;
  const size_t qstride = (size_t)NW * 256;
  __syncthreads();                                       // UxS / DxS visible; no block-level sync after this point

  // L2 prefetch two blocks ahead: dPre block (one bulk request) and the h_{t-1} rows
  auto prefetch_blk = [&](long long pb) {
    if (pb >= blk_end) return;
    if (tid == 0) l2_prefetch_bulk(a.dpre + frag_addr((size_t)pb, 4, 0, NW, 0, 0, 0, 0), (uint32_t)(4 * NW * 256 * sizeof(float)));
    const int tp = (int)(pb / ntiles), bp = (int)(pb % ntiles) * 16;
    if (tid >= 32 && tid < 48 && tp > 0 && (bp + tid - 32) < B)
      l2_prefetch_bulk(a.y + (size_t)(tp - 1) * a.ys_t + (size_t)(bp + tid - 32) * a.ys_b, (uint32_t)(H * sizeof(float)));
  };
  prefetch_blk(blk_begin);
  prefetch_blk(blk_begin + 1);

  // value of slot `slot` of the [z|zx|1] row r
  auto arow = [&](size_t r, int slot) -> float {
    if (slot < RH) return __ldg(a.z + r * a.zp + slot);
    if (slot < RH + RX) return __ldg(a.zx + r * a.zxp + (slot - RH));
    return slot == RH + RX ? 1.f : 0.f;
  };

  // Every warp walks the blocks on its own (no shared staging, no barriers): the per-row operands that all
  // warps need ([z|zx|1], dzc, x) are a few hundred bytes per block and come from L1/L2.
  for (long long blk = blk_begin; blk < blk_end; ++blk) {
    prefetch_blk(blk + 2);
    const int t = (int)(blk / ntiles), tile = (int)(blk % ntiles);
    const int b0 = tile * 16;
    const int nvalid = (B - b0) < 16 ? (B - b0) : 16;
    const float* dblk = a.dpre + frag_addr((size_t)blk, 4, 0, NW, warp, 0, 0, 0);
    // per-block bases (row rr of the block = sequence b0 + rr at timestep t); everything below adds small constants
    const size_t rbase = (size_t)t * B + b0;
    const float* hbase = nullptr;                        // h_{t-1} row of sequence b0, this lane's unit pair
    long long hstride = 0;
    if (uin) {
      if (t > 0) { hbase = a.y + (size_t)(t - 1) * a.ys_t + (size_t)b0 * a.ys_b + ju; hstride = a.ys_b; }
      else if (a.h0) { hbase = a.h0 + (size_t)b0 * H + ju; hstride = H; }
    }
    const float* xbase = a.x + (size_t)t * a.xs_t + (size_t)b0 * a.xs_b + ju;
    const float* dlane = dblk + (Pg * 2) * 64 + (q * 4 + qq) * 2;          // + k*qstride + ks*64 (+32 for row q+4)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {                     // k-step = 8 sequences (hf = ks): lane holds rows q and q+4 of it
      const int rr0 = 8 * ks + q, rr1 = rr0 + 4;
      const bool v0 = rr0 < nvalid, v1 = rr1 < nvalid;
      const size_t r0 = rbase + rr0, r1 = r0 + 4;
      // dPre pieces: (gate k, P = g>>2, hf = ks, source lane = row-in-half * 4 + qq)
      float2 d0[4], d1[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        d0[k] = __ldg(reinterpret_cast<const float2*>(dlane + (size_t)k * qstride + ks * 64));
        d1[k] = __ldg(reinterpret_cast<const float2*>(dlane + (size_t)k * qstride + ks * 64 + 32));
      }
      // h_{t-1} pairs of this lane's units (same for the four gates)
      float2 h0v = make_float2(0.f, 0.f), h1v = h0v;
      if (hbase) {
        if (v0) h0v = __ldg(reinterpret_cast<const float2*>(hbase + (long long)rr0 * hstride));
        if (v1) h1v = __ldg(reinterpret_cast<const float2*>(hbase + (long long)rr1 * hstride));
      }
      float2 x0v = make_float2(0.f, 0.f), x1v = x0v;
      if (xcols) {
        const float* xp0 = xbase + (long long)rr0 * a.xs_b;
        const float* xp1 = xbase + (long long)rr1 * a.xs_b;
        if (ju < I) { if (v0) x0v.x = __ldg(xp0); if (v1) x1v.x = __ldg(xp1); }
        if (ju + 1 < I) { if (v0) x0v.y = __ldg(xp0 + 1); if (v1) x1v.y = __ldg(xp1 + 1); }
      }
      // A operands: [z|zx|1]^T (m = slot g / g+8, k = row) and dzc (k = row, n = slot 8s+g)
      float av[MT][4], dzv[KS][2];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const int s0 = 16 * m + g, s1 = s0 + 8;
        av[m][0] = v0 ? arow(r0, s0) : 0.f; av[m][1] = v0 ? arow(r0, s1) : 0.f;
        av[m][2] = v1 ? arow(r1, s0) : 0.f; av[m][3] = v1 ? arow(r1, s1) : 0.f;
      }
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        dzv[s][0] = v0 ? __ldg(a.dzc + r0 * (8 * KS) + 8 * s + g) : 0.f;
        dzv[s][1] = v1 ? __ldg(a.dzc + r1 * (8 * KS) + 8 * s + g) : 0.f;
      }
      // ---- vector-multiplication gradients ----
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        gDh[k][0] = fmaf(d0[k].x, h0v.x, fmaf(d1[k].x, h1v.x, gDh[k][0]));
        gDh[k][1] = fmaf(d0[k].y, h0v.y, fmaf(d1[k].y, h1v.y, gDh[k][1]));
        if (xcols) {
          gDx[k][0] = fmaf(d0[k].x, x0v.x, fmaf(d1[k].x, x1v.x, gDx[k][0]));
          gDx[k][1] = fmaf(d0[k].y, x0v.y, fmaf(d1[k].y, x1v.y, gDx[k][1]));
        }
      }
      // ---- [z|zx|1]^T dPre ----
      float ah[MT][4], al[MT][4];
#pragma unroll
      for (int m = 0; m < MT; ++m) split4(av[m], ah[m], al[m]);
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int X = 0; X < 2; ++X) {
          const float b0 = X ? d0[k].y : d0[k].x, b1 = X ? d1[k].y : d1[k].x;
          const float b0h = tf32_rna(b0), b1h = tf32_rna(b1);
          const float b0l = b0 - b0h, b1l = b1 - b1h;
#pragma unroll
          for (int m = 0; m < MT; ++m) mma_3x_rn(accW[m][2 * k + X], ah[m], al[m], b0h, b1h, b0l, b1l);
        }
      // ---- Hprev^T dzc and X^T dzc: A = (m = g: pair element 0, m = g+8: element 1; k = row), B = dzc ----
      {
        const float hv[4] = {h0v.x, h0v.y, h1v.x, h1v.y};
        float hh[4], hl[4], xh[4], xl[4];
        split4(hv, hh, hl);
        if (xcols) {
          const float xv[4] = {x0v.x, x0v.y, x1v.x, x1v.y};
          split4(xv, xh, xl);
        }
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const float b0h = tf32_rna(dzv[s][0]), b1h = tf32_rna(dzv[s][1]);
          const float b0l = dzv[s][0] - b0h, b1l = dzv[s][1] - b1h;
          mma_3x_rn(accG[s], hh, hl, b0h, b1h, b0l, b1l);
          if (xcols) mma_3x_rn(accU[s], xh, xl, b0h, b1h, b0l, b1l);
        }
      }
    }
    // ---- dX rows of this block (warps whose units lie below I): dzx Ux^T + sum_k dPre_k Dx_k ----
    if (a.dx && xcols) {
      // lane -> (row = lane>>1, unit = 16*warp + 8*(lane&1) + i), i < 8
      const int rr = lane >> 1;
      if (rr < nvalid) {
        const size_t r = (size_t)t * B + b0 + rr;
        const int hf = rr >> 3, gs = rr & 7, Pu = lane & 1;
        const int jb = warp * 8 + PS * Pu;
        float sx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) sx[i] = 0.f;
        for (int rx = 0; rx < RX; ++rx) {
          const float dzx = __ldg(a.dzc + r * (8 * KS) + RH + rx);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (jb + i < I) sx[i] = fmaf(dzx, UxS[(jb + i) * RX + rx], sx[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = jb + i;
          if (j >= I) continue;
          float s = sx[i];
          const float* dp = dblk + (Pu * 2 + hf) * 64 + (gs * 4 + (i >> 1)) * 2 + (i & 1);
#pragma unroll
          for (int k = 0; k < 4; ++k) s = fmaf(__ldg(dp + (size_t)k * qstride), DxS[k * I + j], s);
          a.dx[(size_t)t * a.dxs_t + (size_t)(b0 + rr) * a.dxs_b + j] = s;
        }
      }
    }
  }

  // ---- this CTA's partial ----
  const GradLayout L(I, H, RX, RH);
  float* P = a.partial + (size_t)blockIdx.x * L.total;
  // accW[m][2k+X] = C fragment: (slot 16m+g, n = 2q / 2q+1), (slot 16m+g+8, same); n <-> unit 16w + 8(n>>2) + 2(n&3) + X
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int X = 0; X < 2; ++X)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int slot = 16 * m + g + 8 * (i >> 1);
          const int n = 2 * q + (i & 1);
          const int j = warp * 8 + PS * (n >> 2) + 2 * (n & 3) + X;
          if (j >= H) continue;
          const size_t row = (size_t)k * H + j;
          const float v = accW[m][2 * k + X][i];
          if (slot < RH) P[L.oBm + row * RH + slot] = v;
          else if (slot < RH + RX) P[L.oVx + row * RX + (slot - RH)] = v;
          else if (slot == RH + RX) P[L.oBias + row] = v;
        }
  // accG / accU: C fragment (m = g [+8], n = slot 8s + 2q [+1]); m = g <-> unit ju(g) + 0, m = g+8 <-> ju(g) + 1
#pragma unroll
  for (int s = 0; s < KS; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = ju + (i >> 1), slot = 8 * s + 2 * q + (i & 1);
      if (j < H && slot < RH) P[L.oA + (size_t)j * RH + slot] = accG[s][i];
      if (j < I && slot >= RH && slot < RH + RX) P[L.oUx + (size_t)j * RX + (slot - RH)] = accU[s][i];
    }
  // dDh / dDx: sum the four q lanes (rows); lane q == 0 stores its unit pair of every gate
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int X = 0; X < 2; ++X) {
      float v = gDh[k][X], w = gDx[k][X];
      v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
      w += __shfl_xor_sync(0xffffffffu, w, 1); w += __shfl_xor_sync(0xffffffffu, w, 2);
      const int j = ju + X;
      if (q == 0 && j < H) P[L.oDh + k * H + j] = v;
      if (q == 0 && j < I) P[L.oDx + k * I + j] = w;
    }
}

int launch_bwd_mma(const SeqBwdMmaArgs& a, const GradRowsArgs& gr, const GradOut& out, void* workspace, int* n_parts,
                   cudaStream_t st);
// workspace floats needed by the MMA backward (dPre + dzc + per-CTA partials), 0 when the shape is not covered
long long bwd_mma_workspace_floats(int T, int B, int I, int H, int RX, int RH);
bool bwd_mma_fits(int I, int H, int RX, int RH);

}  // namespace vmlmf
