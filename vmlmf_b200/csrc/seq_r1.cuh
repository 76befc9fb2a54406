// seq_r1.cuh -- regime R1 of the VMLMF recurrence: persistent "unit-owner" kernels.
//
// One CTA owns a tile of BT sequences for all T timesteps; thread j owns hidden unit j and
// keeps that unit's slice of every factor in REGISTERS for the whole kernel:
//   Bm[kH+j, 0:RH], Vx[kH+j, 0:RX] (k = i,f,o,n), A[j, 0:RH], Dh[k,j], Dx[k,j], bias[kH+j].
// Per step the only cross-thread traffic is the rank-sized vector z = h_{t-1} A  (RH floats
// per sequence): each thread forms its partial h[b,j]*A[j,:], a transposing warp butterfly
// (warp_multi_reduce) sums 32 units, one __syncthreads + a fixed-order sum over warps
// finishes it.  Everything else -- the two rank contractions, the diagonal (vector-
// multiplication) terms, bias, the four gate non-linearities, the c/h update -- is
// thread-local FMAs on registers.  The backward kernel runs the same tiling in reverse
// time and additionally keeps the unit's slice of every factor GRADIENT in registers,
// summed over all sequences/timesteps the CTA sees; CTAs write their partials to a
// workspace and a second tiny kernel adds them in a fixed order (bit-reproducible).
//
// Replaces (reference, "V/" = rnn_compression_factorization_vmlmf/src/):
//   forward : V/models/vmlmf.py:308-310 x :78-125, vmlmf_group.py:85-155, vmlmf_lm.py:272-280
//   backward: the autograd replay of those ops (SURVEY Appendix A.3 is the math).
#pragma once
#include "common.cuh"

namespace vmlmf {

struct SeqFwdArgs {
  const float* x; long long xs_t, xs_b;
  const float* zx;                                   // [T*B, RXP]
  const float *Vx, *Dx, *A, *Bm, *Dh, *bias;
  const float *h0, *c0;
  float* y; long long ys_t, ys_b;
  float *hT, *cT;
  float *gates, *cs, *z;                             // saved (training) or null
  int T, B, I, H, RX, RH;
};

struct SeqBwdArgs {
  const float* x; long long xs_t, xs_b;
  const float* zx;
  const float *Ux, *Vx, *Dx, *A, *Bm, *Dh;
  const float *h0, *c0;
  const float* y; long long ys_t, ys_b;
  const float *gates, *cs, *z;
  const float* dy; long long dys_t, dys_b;
  const float *dhT, *dcT;
  float* dx; long long dxs_t, dxs_b;
  float *dh0, *dc0;
  float* partial;                                    // [gridDim.x, P] factor-gradient partials
  int T, B, I, H, RX, RH;
};

// flat layout of one CTA's factor-gradient partial (and of the reduced result)
struct GradLayout {
  int oUx, oVx, oDx, oA, oBm, oDh, oBias, total;
  __host__ __device__ GradLayout(int I, int H, int RX, int RH) {
    oUx = 0;
    oVx = oUx + I * RX;
    oDx = oVx + 4 * H * RX;
    oA = oDx + 4 * I;
    oBm = oA + H * RH;
    oDh = oBm + 4 * H * RH;
    oBias = oDh + 4 * H;
    total = oBias + 4 * H;
  }
};

template <int N>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&out)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      const float4 v = reinterpret_cast<const float4*>(p)[q];
      out[4 * q] = v.x; out[4 * q + 1] = v.y; out[4 * q + 2] = v.z; out[4 * q + 3] = v.w;
    }
  } else if constexpr (N % 2 == 0) {
#pragma unroll
    for (int q = 0; q < N / 2; ++q) {
      const float2 v = reinterpret_cast<const float2*>(p)[q];
      out[2 * q] = v.x; out[2 * q + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int q = 0; q < N; ++q) out[q] = p[q];
  }
}

// ------------------------------------------------------------------------------------- //
// forward
// ------------------------------------------------------------------------------------- //
template <int RH_T, int RX_T, int BT, bool SAVE, int NT_MAX, int MIN_CTAS>
__global__ void __launch_bounds__(NT_MAX, MIN_CTAS) seq_fwd_r1_kernel(const SeqFwdArgs a) {
  constexpr int RHP = next_pow2(RH_T);        // pitch of a z row (smem, and saved z)
  constexpr int RXP = round_up(RX_T, 4);      // pitch of a zx row
  constexpr int NV = BT * RHP;                // values reduced across the CTA per step
  constexpr int N = NV < 32 ? NV : 32;        // values per butterfly round
  constexpr int ROUNDS = NV / N;
  constexpr int SPR = N / RHP;                // sequences per round
  constexpr int NZX = BT * RXP;
  constexpr int ZXL = ceil_div(NZX, 32);
  static_assert(RHP <= 32, "rank too large for R1");

  const int j = threadIdx.x, lane = j & 31, warp = j >> 5, NW = blockDim.x >> 5;
  const int H = a.H, I = a.I, B = a.B, T = a.T;
  const bool live = j < H, hasx = j < I;

  extern __shared__ __align__(16) float smem[];
  float* part = smem;                                    // [2][NW][NV]
  float* zw = smem + 2 * NW * NV + warp * NV;            // this warp's copy of z   [BT][RHP]
  float* zxw = smem + 3 * NW * NV + warp * NZX;          // this warp's copy of zx  [BT][RXP]

  // ---- this unit's factor slices -> registers (held for the whole kernel) ----
  float wB[4][RH_T], wV[4][RX_T], wA[RH_T], wDh[4], wDx[4], wb[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int r = 0; r < RH_T; ++r)
      wB[k][r] = (live && r < a.RH) ? __ldg(a.Bm + (size_t)(k * H + j) * a.RH + r) : 0.f;
#pragma unroll
    for (int r = 0; r < RX_T; ++r)
      wV[k][r] = (live && r < a.RX) ? __ldg(a.Vx + (size_t)(k * H + j) * a.RX + r) : 0.f;
    wDh[k] = live ? __ldg(a.Dh + k * H + j) : 0.f;
    wDx[k] = hasx ? __ldg(a.Dx + k * I + j) : 0.f;
    wb[k] = live ? __ldg(a.bias + k * H + j) : 0.f;
  }
#pragma unroll
  for (int r = 0; r < RH_T; ++r) wA[r] = (live && r < a.RH) ? __ldg(a.A + (size_t)j * a.RH + r) : 0.f;

  const int vidx = warp_multi_reduce_index<N>(lane);
  const int ntiles = ceil_div(B, BT);
  unsigned it = 0;                                       // running reduce counter (double buffer)

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0 = tile * BT;
    float h[BT], c[BT], xn[BT], zxn[ZXL];
    bool ok[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      ok[b] = (b0 + b) < B;
      const bool ld = ok[b] && live;
      h[b] = (ld && a.h0) ? a.h0[(size_t)(b0 + b) * H + j] : 0.f;
      c[b] = (ld && a.c0) ? a.c0[(size_t)(b0 + b) * H + j] : 0.f;
      xn[b] = (ok[b] && hasx) ? a.x[(size_t)(b0 + b) * a.xs_b + j] : 0.f;      // t = 0
    }
#pragma unroll
    for (int q = 0; q < ZXL; ++q) {
      const int idx = q * 32 + lane, b = idx / RXP, r = idx % RXP;
      zxn[q] = (idx < NZX && (b0 + b) < B) ? a.zx[(size_t)(b0 + b) * RXP + r] : 0.f;
    }

    // z for step 0 and the rolling per-step reduction share one code path
    auto reduce_z = [&](void) {
      float* pbuf = part + (it & 1u) * NW * NV + warp * NV;
#pragma unroll
      for (int rd = 0; rd < ROUNDS; ++rd) {
        float pv[N];
#pragma unroll
        for (int s = 0; s < SPR; ++s)
#pragma unroll
          for (int r = 0; r < RHP; ++r) pv[s * RHP + r] = (r < RH_T) ? h[rd * SPR + s] * wA[r < RH_T ? r : 0] : 0.f;
        const float tot = warp_multi_reduce<N>(pv, lane);
        if (lane < N) pbuf[rd * N + vidx] = tot;
      }
      __syncthreads();
      const float* pall = part + (it & 1u) * NW * NV;
      for (int idx = lane; idx < NV; idx += 32) {
        float s = 0.f;
        for (int w = 0; w < NW; ++w) s += pall[w * NV + idx];
        zw[idx] = s;
      }
#pragma unroll
      for (int q = 0; q < ZXL; ++q) {
        const int idx = q * 32 + lane;
        if (idx < NZX) zxw[idx] = zxn[q];
      }
      __syncwarp();
      ++it;
    };
    reduce_z();

    for (int t = 0; t < T; ++t) {
      float xv[BT];
#pragma unroll
      for (int b = 0; b < BT; ++b) xv[b] = xn[b];
      if (t + 1 < T) {                                   // prefetch step t+1 (independent of the recurrence)
#pragma unroll
        for (int b = 0; b < BT; ++b)
          xn[b] = (ok[b] && hasx) ? a.x[(size_t)(t + 1) * a.xs_t + (size_t)(b0 + b) * a.xs_b + j] : 0.f;
#pragma unroll
        for (int q = 0; q < ZXL; ++q) {
          const int idx = q * 32 + lane, b = idx / RXP, r = idx % RXP;
          zxn[q] = (idx < NZX && (b0 + b) < B) ? a.zx[((size_t)(t + 1) * B + b0 + b) * RXP + r] : 0.f;
        }
      }
      if (SAVE && warp == 0) {                           // z_t = h_{t-1} A, needed by backward
        for (int idx = lane; idx < NV; idx += 32) {
          const int b = idx / RHP;
          if ((b0 + b) < B) a.z[((size_t)t * B + b0 + b) * RHP + (idx % RHP)] = zw[idx];
        }
      }
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        float zr[RHP], zxr[RXP], pre[4];
        load_row<RHP>(zw + b * RHP, zr);
        load_row<RXP>(zxw + b * RXP, zxr);
#pragma unroll
        for (int k = 0; k < 4; ++k) pre[k] = fmaf(wDx[k], xv[b], fmaf(wDh[k], h[b], wb[k]));
#pragma unroll
        for (int r = 0; r < RH_T; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) pre[k] = fmaf(zr[r], wB[k][r], pre[k]);
#pragma unroll
        for (int r = 0; r < RX_T; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) pre[k] = fmaf(zxr[r], wV[k][r], pre[k]);
        const float gi = sigmoidf_acc(pre[0]), gf = sigmoidf_acc(pre[1]);
        const float go = sigmoidf_acc(pre[2]), gn = tanhf_acc(pre[3]);
        c[b] = fmaf(gf, c[b], gi * gn);
        h[b] = go * tanhf_acc(c[b]);
        if (ok[b] && live) {
          if (a.y) a.y[(size_t)t * a.ys_t + (size_t)(b0 + b) * a.ys_b + j] = h[b];      // y == nullptr: last-step-only caller
          if (SAVE) {
            float* g = a.gates + ((size_t)t * B + b0 + b) * 4 * H + j;
            g[0] = gi; g[H] = gf; g[2 * H] = go; g[3 * H] = gn;
            a.cs[((size_t)t * B + b0 + b) * H + j] = c[b];
          }
        }
      }
      if (t + 1 < T) reduce_z();
    }
#pragma unroll
    for (int b = 0; b < BT; ++b)
      if (ok[b] && live) {
        a.hT[(size_t)(b0 + b) * H + j] = h[b];
        a.cT[(size_t)(b0 + b) * H + j] = c[b];
      }
  }
}

// ------------------------------------------------------------------------------------- //
// backward through time
// ------------------------------------------------------------------------------------- //
template <int RH_T, int RX_T, int BT, int NT_MAX, int MIN_CTAS>
__global__ void __launch_bounds__(NT_MAX, MIN_CTAS) seq_bwd_r1_kernel(const SeqBwdArgs a) {
  constexpr int RHP = next_pow2(RH_T);        // pitch of saved z rows
  constexpr int RXP = round_up(RX_T, 4);      // pitch of zx rows
  constexpr int VPS = next_pow2(RH_T + RX_T); // reduced values per sequence: [dz | dzx | pad]
  constexpr int NV = BT * VPS;
  constexpr int N = NV < 32 ? NV : 32;
  constexpr int ROUNDS = NV / N;
  constexpr int SPR = N / VPS;
  constexpr int NIN = BT * (RHP + RXP);       // broadcast inputs per tile-step: z rows then zx rows
  constexpr int INL = ceil_div(NIN, 32);
  static_assert(VPS <= 32, "ranks too large for R1 backward");

  const int j = threadIdx.x, lane = j & 31, warp = j >> 5, NW = blockDim.x >> 5;
  const int H = a.H, I = a.I, B = a.B, T = a.T;
  const bool live = j < H, hasx = j < I;

  extern __shared__ __align__(16) float smem[];
  float* part = smem;                                     // [2][NW][NV]
  float* dzw = smem + 2 * NW * NV + warp * NV;            // this warp's copy of [dz|dzx] per sequence
  float* inw = smem + 3 * NW * NV + warp * NIN;           // this warp's copy of z / zx rows

  float wB[4][RH_T], wV[4][RX_T], wA[RH_T], wU[RX_T], wDh[4], wDx[4];
  float gB[4][RH_T], gV[4][RX_T], gA[RH_T], gU[RX_T], gDh[4], gDx[4], gb[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int r = 0; r < RH_T; ++r) {
      wB[k][r] = (live && r < a.RH) ? __ldg(a.Bm + (size_t)(k * H + j) * a.RH + r) : 0.f;
      gB[k][r] = 0.f;
    }
#pragma unroll
    for (int r = 0; r < RX_T; ++r) {
      wV[k][r] = (live && r < a.RX) ? __ldg(a.Vx + (size_t)(k * H + j) * a.RX + r) : 0.f;
      gV[k][r] = 0.f;
    }
    wDh[k] = live ? __ldg(a.Dh + k * H + j) : 0.f;
    wDx[k] = hasx ? __ldg(a.Dx + k * I + j) : 0.f;
    gDh[k] = gDx[k] = gb[k] = 0.f;
  }
#pragma unroll
  for (int r = 0; r < RH_T; ++r) {
    wA[r] = (live && r < a.RH) ? __ldg(a.A + (size_t)j * a.RH + r) : 0.f;
    gA[r] = 0.f;
  }
#pragma unroll
  for (int r = 0; r < RX_T; ++r) {
    wU[r] = (hasx && r < a.RX) ? __ldg(a.Ux + (size_t)j * a.RX + r) : 0.f;
    gU[r] = 0.f;
  }

  const int vidx = warp_multi_reduce_index<N>(lane);
  const int ntiles = ceil_div(B, BT);
  unsigned it = 0;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0 = tile * BT;
    bool ok[BT];
    float dhn[BT], dcn[BT], ct[BT];
    // prefetched operands of the step about to be processed
    float pg[BT][4], pcp[BT], php[BT], pdy[BT], px[BT], pin[INL];

    auto fetch = [&](int t) {
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const size_t bb = (size_t)(b0 + b);
        const bool ld = ok[b] && live;
        const float* g = a.gates + ((size_t)t * B + bb) * 4 * H + j;
#pragma unroll
        for (int k = 0; k < 4; ++k) pg[b][k] = ld ? g[k * H] : 0.f;
        if (t > 0) {
          pcp[b] = ld ? a.cs[((size_t)(t - 1) * B + bb) * H + j] : 0.f;
          php[b] = ld ? a.y[(size_t)(t - 1) * a.ys_t + bb * a.ys_b + j] : 0.f;
        } else {
          pcp[b] = (ld && a.c0) ? a.c0[bb * H + j] : 0.f;
          php[b] = (ld && a.h0) ? a.h0[bb * H + j] : 0.f;
        }
        pdy[b] = (ld && a.dy) ? a.dy[(size_t)t * a.dys_t + bb * a.dys_b + j] : 0.f;
        px[b] = (ok[b] && hasx) ? a.x[(size_t)t * a.xs_t + bb * a.xs_b + j] : 0.f;
      }
#pragma unroll
      for (int q = 0; q < INL; ++q) {
        const int idx = q * 32 + lane;
        float v = 0.f;
        if (idx < BT * RHP) {
          const int b = idx / RHP;
          if ((b0 + b) < B) v = a.z[((size_t)t * B + b0 + b) * RHP + idx % RHP];
        } else if (idx < NIN) {
          const int i2 = idx - BT * RHP, b = i2 / RXP;
          if ((b0 + b) < B) v = a.zx[((size_t)t * B + b0 + b) * RXP + i2 % RXP];
        }
        pin[q] = v;
      }
    };

#pragma unroll
    for (int b = 0; b < BT; ++b) {
      const size_t bb = (size_t)(b0 + b);
      ok[b] = (b0 + b) < B;
      const bool ld = ok[b] && live;
      dhn[b] = (ld && a.dhT) ? a.dhT[bb * H + j] : 0.f;
      dcn[b] = (ld && a.dcT) ? a.dcT[bb * H + j] : 0.f;
      ct[b] = ld ? a.cs[((size_t)(T - 1) * B + bb) * H + j] : 0.f;
    }
    fetch(T - 1);

    for (int t = T - 1; t >= 0; --t) {
      // ---- stage this step's operands, start fetching the next (earlier) step ----
      float g4[BT][4], cprev[BT], hprev[BT], dyv[BT], xv[BT];
#pragma unroll
      for (int b = 0; b < BT; ++b) {
#pragma unroll
        for (int k = 0; k < 4; ++k) g4[b][k] = pg[b][k];
        cprev[b] = pcp[b]; hprev[b] = php[b]; dyv[b] = pdy[b]; xv[b] = px[b];
      }
      __syncwarp();                                      // all lanes done reading inw of the previous step
#pragma unroll
      for (int q = 0; q < INL; ++q) {
        const int idx = q * 32 + lane;
        if (idx < NIN) inw[idx] = pin[q];
      }
      __syncwarp();
      if (t > 0) fetch(t - 1);

      // ---- phase A: gate gradients, factor-gradient accumulation, partial dz / dzx ----
      float dpre[BT][4], dhd[BT], dxd[BT];
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float gi = g4[b][0], gf = g4[b][1], go = g4[b][2], gn = g4[b][3];
        const float dh = dyv[b] + dhn[b];
        const float tc = tanhf_acc(ct[b]);
        const float dc = fmaf(dh * go, 1.f - tc * tc, dcn[b]);
        dpre[b][0] = dc * gn * gi * (1.f - gi);
        dpre[b][1] = dc * cprev[b] * gf * (1.f - gf);
        dpre[b][2] = dh * tc * go * (1.f - go);
        dpre[b][3] = dc * gi * (1.f - gn * gn);
        dcn[b] = dc * gf;
        ct[b] = cprev[b];
        float zr[RHP], zxr[RXP];
        load_row<RHP>(inw + b * RHP, zr);
        load_row<RXP>(inw + BT * RHP + b * RXP, zxr);
        float sd = 0.f, sx = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float d = dpre[b][k];
          gb[k] += d;
          gDh[k] = fmaf(d, hprev[b], gDh[k]);
          gDx[k] = fmaf(d, xv[b], gDx[k]);
          sd = fmaf(d, wDh[k], sd);
          sx = fmaf(d, wDx[k], sx);
#pragma unroll
          for (int r = 0; r < RH_T; ++r) gB[k][r] = fmaf(d, zr[r], gB[k][r]);
#pragma unroll
          for (int r = 0; r < RX_T; ++r) gV[k][r] = fmaf(d, zxr[r], gV[k][r]);
        }
        dhd[b] = sd; dxd[b] = sx;
      }
      {
        float* pbuf = part + (it & 1u) * NW * NV + warp * NV;
#pragma unroll
        for (int rd = 0; rd < ROUNDS; ++rd) {
          float pv[N];
#pragma unroll
          for (int s = 0; s < SPR; ++s) {
            const int b = rd * SPR + s;
#pragma unroll
            for (int v = 0; v < VPS; ++v) {
              float acc = 0.f;
              if (v < RH_T) {
#pragma unroll
                for (int k = 0; k < 4; ++k) acc = fmaf(dpre[b][k], wB[k][v < RH_T ? v : 0], acc);
              } else if (v < RH_T + RX_T) {
#pragma unroll
                for (int k = 0; k < 4; ++k) acc = fmaf(dpre[b][k], wV[k][(v >= RH_T && v < RH_T + RX_T) ? v - RH_T : 0], acc);
              }
              pv[s * VPS + v] = acc;
            }
          }
          const float tot = warp_multi_reduce<N>(pv, lane);
          if (lane < N) pbuf[rd * N + vidx] = tot;
        }
        __syncthreads();
        const float* pall = part + (it & 1u) * NW * NV;
        for (int idx = lane; idx < NV; idx += 32) {
          float s = 0.f;
          for (int w = 0; w < NW; ++w) s += pall[w * NV + idx];
          dzw[idx] = s;
        }
        __syncwarp();
        ++it;
      }
      // ---- phase B: dh_{t-1}, dx_t, gradients of A and Ux ----
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        float dv[VPS];
        load_row<VPS>(dzw + b * VPS, dv);
        float s = dhd[b];
#pragma unroll
        for (int r = 0; r < RH_T; ++r) {
          s = fmaf(dv[r], wA[r], s);
          gA[r] = fmaf(hprev[b], dv[r], gA[r]);
        }
        dhn[b] = s;
        float sx = dxd[b];
#pragma unroll
        for (int r = 0; r < RX_T; ++r) {
          sx = fmaf(dv[RH_T + r], wU[r], sx);
          gU[r] = fmaf(xv[b], dv[RH_T + r], gU[r]);
        }
        if (a.dx && ok[b] && hasx) a.dx[(size_t)t * a.dxs_t + (size_t)(b0 + b) * a.dxs_b + j] = sx;
      }
    }
#pragma unroll
    for (int b = 0; b < BT; ++b)
      if (ok[b] && live) {
        if (a.dh0) a.dh0[(size_t)(b0 + b) * H + j] = dhn[b];
        if (a.dc0) a.dc0[(size_t)(b0 + b) * H + j] = dcn[b];
      }
  }

  // ---- this CTA's factor-gradient partial ----
  const GradLayout L(I, H, a.RX, a.RH);
  float* P = a.partial + (size_t)blockIdx.x * L.total;
  if (live) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int r = 0; r < RH_T; ++r)
        if (r < a.RH) P[L.oBm + (size_t)(k * H + j) * a.RH + r] = gB[k][r];
#pragma unroll
      for (int r = 0; r < RX_T; ++r)
        if (r < a.RX) P[L.oVx + (size_t)(k * H + j) * a.RX + r] = gV[k][r];
      P[L.oDh + k * H + j] = gDh[k];
      P[L.oBias + k * H + j] = gb[k];
      if (hasx) P[L.oDx + k * I + j] = gDx[k];
    }
#pragma unroll
    for (int r = 0; r < RH_T; ++r)
      if (r < a.RH) P[L.oA + (size_t)j * a.RH + r] = gA[r];
    if (hasx) {
#pragma unroll
      for (int r = 0; r < RX_T; ++r)
        if (r < a.RX) P[L.oUx + (size_t)j * a.RX + r] = gU[r];
    }
  }
}

// Fixed-order sum of the per-CTA partials -> the seven canonical gradient tensors.
struct GradOut { float *dUx, *dVx, *dDx, *dA, *dBm, *dDh, *dbias; };

// Launch with kReduceElems outputs per 256-thread block: every output is summed by 8 threads over interleaved
// slices of the partial list (enough independent loads in flight to cover DRAM/L2 latency), then across the slices
// in slice order -- the association is fixed, so results reproduce bit for bit.
constexpr int kReduceElems = 32;
static __global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, int nparts,
                                                                     GradLayout L, GradOut o) {
  __shared__ float sl[8][kReduceElems + 1];
  const int e = threadIdx.x & (kReduceElems - 1), slice = threadIdx.x / kReduceElems;
  const int p = blockIdx.x * kReduceElems + e;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (p < L.total) {
    int c = slice;
    for (; c + 24 < nparts; c += 32) {
      s0 += partial[(size_t)c * L.total + p];
      s1 += partial[(size_t)(c + 8) * L.total + p];
      s2 += partial[(size_t)(c + 16) * L.total + p];
      s3 += partial[(size_t)(c + 24) * L.total + p];
    }
    for (; c < nparts; c += 8) s0 += partial[(size_t)c * L.total + p];
  }
  sl[slice][e] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (slice != 0 || p >= L.total) return;
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += sl[q][e];
  if (p < L.oVx) o.dUx[p - L.oUx] = s;
  else if (p < L.oDx) o.dVx[p - L.oVx] = s;
  else if (p < L.oA) o.dDx[p - L.oDx] = s;
  else if (p < L.oBm) o.dA[p - L.oA] = s;
  else if (p < L.oDh) o.dBm[p - L.oBm] = s;
  else if (p < L.oBias) o.dDh[p - L.oDh] = s;
  else o.dbias[p - L.oBias] = s;
}

}  // namespace vmlmf
