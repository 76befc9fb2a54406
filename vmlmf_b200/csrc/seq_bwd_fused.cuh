// seq_bwd_fused.cuh -- regime R1M backward with the weight-gradient accumulation fused into the reverse-time
// recurrence: dPre never leaves the SM.
//
// Same tiling as the forward (seq_mma.cuh): CTA = 16 sequences, warp w = the 8-unit halves [8w, 8w+8) (P = 0) and
// [8NW+8w, 8NW+8w+8) (P = 1), lane (g,q) = units half(P)+2q+{0,1} of sequences g, g+8.  The split design wrote dPre[T*B,4H] to HBM for a second kernel
// (0.8 GB out + 0.8 GB in at the bench workload: traffic twice the algorithmic bytes).  Here the gradient
// accumulators do not fit the register file next to the recurrence state (the kernel already runs at the
// 128-register cap of a 512-thread CTA), so they live in TENSOR MEMORY: every warp owns 80 TMEM columns x 32 lanes
// and moves accumulator fragments with tcgen05.ld / tcgen05.st around its mma.sync products:
//   columns  0..31  [z|zx|1]^T dPre   (dBm | dVx | dbias)   8 n-tiles (gate k, half P) x 4 registers
//   columns 32..47  dDh partial sums of the lane's 8 (unit, gate) pairs... per half P: 8 registers
//   columns 48..63  dDx partial sums (x-side warps)
//   columns 64..71  dA  = Hprev^T dz     columns 72..79  dUx = X^T dzx
// Per step and half P: dPre fragments -> per-warp shared tile (accumulator layout) -> read back transposed as the
// B operand (K = the 16 sequences) of the [z|zx|1]^T dPre product; A = the step's [z|zx|1] rows, brought in one
// step ahead with cp.async.  dX = dzx Ux^T + sum_k dPre_k Dx_k is finished in the same kernel (x-side warps), and so
// are the two products that need h_{t-1} and x TRANSPOSED against the reduced dzc (dA = Hprev^T dz, dUx = X^T dzx):
// their operands are the rows this warp just read (L1/L2 hits).  One partial per CTA; reduce_partials_kernel sums.
#pragma once
#include "gemm_tc.cuh"
#include "seq_bwd_mma.cuh"

namespace vmlmf {

struct SeqBwdFusedArgs {
  const float *gates, *cs, *c0;
  const float* dy; long long dys_t, dys_b;
  const float *dhT, *dcT;
  const float *Ux, *Vx, *Dx, *A, *Bm, *Dh;
  const float* x; long long xs_t, xs_b;
  const float* y; long long ys_t, ys_b;
  const float* h0;
  const float *z, *zx; int zp, zxp;
  float* dzc;                              // out: [T*B, zxp] rows of dzx = (reduced dzc)[RH .. RH+RX), for dux_rows_kernel
  int dz_bt;                               // row order of dzc: 1 = b*T + t (x is a contiguous [B,T,I] block), 0 = t*B + b
  float* dx; long long dxs_t, dxs_b;       // may be null
  float *dh0, *dc0;
  float* partial;                          // [gridDim.x, GradLayout.total]; this kernel writes dVx, dDx, dBm, dDh, dbias
  int T, B, I, H, RX, RH;
};

constexpr int kFusedTP = 40;               // pitch of the per-warp dPre tile (16 sequences x 32 columns of one half P)
constexpr int kFusedCols = 80;             // TMEM columns per warp: 32 dW, 16 dDh, 16 dDx, 8 dA, 8 dUx (+ 8 spare with KS = 1)
constexpr int kFusedAP = 17;               // pitch of a staged [z|zx|1] row (<= 16 slots: KS <= 2)

inline size_t seq_bwd_fused_smem_bytes(int NW, int KS, int I, int RX) {
  size_t fl = (size_t)NW * 8 * KS * 32 * 4             // B fragments of [Bm|Vx]
              + (size_t)NW * 16 * kFusedTP             // per-warp dPre tiles
              + (size_t)NW * 16 * bwd_pitch(KS)        // dzc partials
              + (size_t)16 * bwd_pitch(KS)             // reduced dzc rows
              + 2 * 4 * (size_t)NW * 16                // Dh, Dx [4][HP]
              + 2 * 16 * (size_t)kFusedAP              // [z|zx|1] rows, double buffered
              + (size_t)I * RX                         // Ux
              + 4;                                     // TMEM base address slot
  return fl * sizeof(float);
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
               : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
// 8 consecutive columns (two accumulator fragments) in one instruction
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[2][4]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i >> 2][i & 3] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[2][4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0][0])), "r"(__float_as_uint(v[0][1])), "r"(__float_as_uint(v[0][2])),
               "r"(__float_as_uint(v[0][3])), "r"(__float_as_uint(v[1][0])), "r"(__float_as_uint(v[1][1])),
               "r"(__float_as_uint(v[1][2])), "r"(__float_as_uint(v[1][3]))
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// KS <= 2 (one 16-slot m-tile of [z|zx|1]); NZ: n-tiles of the dh GEMM
template <int KS, int NZ>
__global__ void __launch_bounds__(512, 1) seq_bwd_fused_kernel(const SeqBwdFusedArgs a, const int tmem_cols) {
  constexpr int PP = bwd_pitch(KS), TP = kFusedTP, AP = kFusedAP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int H = a.H, I = a.I, B = a.B, T = a.T, RH = a.RH, RX = a.RX;
  const int HP = NW * 16;
  const int ubase = warp * 8;                            // first unit of this warp's half P = 0
  const int PS = 8 * (blockDim.x >> 5);                  // half P = 1 holds units PS + 8 warp .. (mapping of seq_mma.cuh)
  const int j0 = ubase + 2 * q;                          // this lane's units: j0 + PS*P + e
  const bool xwarp = ubase < I;                          // some unit of this warp has an input-side term ...
  const bool xhalf[2] = {ubase < I, ubase + PS < I};     // ... per half

  extern __shared__ __align__(16) float smem[];
  float4* Bf = reinterpret_cast<float4*>(smem);          // [NW][2 P][4 k][KS][32]
  float* Tt = smem + (size_t)NW * 8 * KS * 32 * 4;       // [NW][16][TP]
  float* Pz = Tt + (size_t)NW * 16 * TP;                 // [NW][16][PP]
  float* Dz = Pz + (size_t)NW * 16 * PP;                 // [16][PP]
  float* DhS = Dz + 16 * PP;                             // [4][HP]
  float* DxS = DhS + 4 * HP;                             // [4][HP] (zero beyond I)
  float* Ar = DxS + 4 * HP;                              // [2][16][AP]
  float* UxS = Ar + 2 * 16 * AP;                         // [I][RX]
  uint32_t* tslot = reinterpret_cast<uint32_t*>(UxS + I * RX);

  // ---- tensor memory: accumulator space, 64 columns per warp ----
  if (warp == 0) {
    const uint32_t ts = (uint32_t)__cvta_generic_to_shared(tslot);
    // the column count as an immediate (tmem_cols is a power of two in [32, 512])
    switch (tmem_cols) {
      case 32: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(ts)); break;
      case 64: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(ts)); break;
      case 128: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(ts)); break;
      case 256: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(ts)); break;
      default: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(ts)); break;
    }
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // ---- prologue (as K3a) ----
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
    for (int s = 0; s < KS; ++s) {
      const float b0 = wc(k, j0 + PS * P, 8 * s + g), b1 = wc(k, j0 + PS * P + 1, 8 * s + g);
      const float b0h = tf32_rna(b0), b1h = tf32_rna(b1);
      myB[(pk * KS + s) * 32] = make_float4(b0h, b1h, tf32_rna(b0 - b0h), tf32_rna(b1 - b1h));
    }
  }
  for (int i = tid; i < 4 * HP; i += blockDim.x) {
    const int k = i / HP, j = i - k * HP;
    DhS[i] = (j < H) ? __ldg(a.Dh + k * H + j) : 0.f;
    DxS[i] = (j < I) ? __ldg(a.Dx + k * I + j) : 0.f;
  }
  for (int i = tid; i < I * RX; i += blockDim.x) UxS[i] = __ldg(a.Ux + i);
  for (int i = tid; i < 2 * 16 * AP; i += blockDim.x) Ar[i] = ((i % AP) == RH + RX) ? 1.f : 0.f;
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
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = *tslot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * kFusedCols);
  {
    const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < kFusedCols; c += 4) tmem_st4(tbase + c, zero4);
    tmem_wait_st();
  }

  const int ntiles = ceil_div(B, 16);
  const int nthreads = blockDim.x;
  const size_t gstep = (size_t)ntiles * 4 * NW * 256, cstep = (size_t)ntiles * NW * 256, qstride = (size_t)NW * 256;
  const bool dy_vec = a.dy && ((reinterpret_cast<uintptr_t>(a.dy) & 7) == 0) && !(a.dys_t & 1) && !(a.dys_b & 1);
  float* myT = Tt + (size_t)warp * 16 * TP;

  // staging of the [z|zx] part of the [z|zx|1] rows: thread i < 16*(RH+RX) owns element (row i / (RH+RX), slot i % (RH+RX))
  // and walks its source pointer backwards in time (one 4-byte cp.async per step)
  // (CTAs with fewer threads than elements -- H < 128 -- let a thread take the further elements tid + k*nthreads
  //  through the slower generic path)
  const int st_rr = tid / (RH + RX), st_slot = tid - st_rr * (RH + RX);
  const bool st_own = tid < 16 * (RH + RX);
  const long long st_pitch = st_slot < RH ? a.zp : a.zxp;
  const float* st_base = st_slot < RH ? a.z + st_slot : a.zx + (st_slot - RH);
  const bool st_more = 16 * (RH + RX) > (int)blockDim.x;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0 = tile * 16;
    const int sq[2] = {b0 + g, b0 + g + 8};
    bool ok[2][2];                                       // [hf][P]
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
    const float* gfrag = a.gates + frag_addr((size_t)(T - 1) * ntiles + tile, 4, 0, NW, warp, 0, 0, lane);
    const float* cfrag = a.cs + frag_addr((size_t)(T - 1) * ntiles + tile, 1, 0, NW, warp, 0, 0, lane);
    auto prefetch_step = [&](int tp) {
      if (tid != 0 || tp < 0) return;
      l2_prefetch_bulk(a.gates + frag_addr((size_t)tp * ntiles + tile, 4, 0, NW, 0, 0, 0, 0), (uint32_t)(4 * NW * 256 * sizeof(float)));
      if (tp > 0) l2_prefetch_bulk(a.cs + frag_addr((size_t)(tp - 1) * ntiles + tile, 1, 0, NW, 0, 0, 0, 0), (uint32_t)(NW * 256 * sizeof(float)));
    };
    if (tid == 0) l2_prefetch_bulk(a.cs + frag_addr((size_t)(T - 1) * ntiles + tile, 1, 0, NW, 0, 0, 0, 0), (uint32_t)(NW * 256 * sizeof(float)));
    prefetch_step(T - 1);
    prefetch_step(T - 2);
    __syncthreads();                                     // previous tile finished with every shared buffer
    const bool st_valid = st_own && (b0 + st_rr) < B;
    const float* st_src = st_base + ((size_t)(T - 1) * B + b0 + st_rr) * st_pitch;      // row (T-1, b0 + st_rr)
    auto stage_rows = [&](int ts) {                      // rows of timestep ts -> buffer ts & 1
      if (ts < 0 || !st_own) return;
      float* dst = Ar + (size_t)(ts & 1) * 16 * AP + st_rr * AP + st_slot;
      if (st_valid) cp_async4(dst, st_src);
      else *dst = 0.f;
      st_src -= (size_t)B * st_pitch;
      if (st_more)
        for (int i = tid + nthreads; i < 16 * (RH + RX); i += nthreads) {
          const int rr = i / (RH + RX), slot = i - rr * (RH + RX);
          float* d2 = Ar + (size_t)(ts & 1) * 16 * AP + rr * AP + slot;
          if (b0 + rr < B) {
            const size_t r = (size_t)ts * B + b0 + rr;
            cp_async4(d2, slot < RH ? a.z + r * a.zp + slot : a.zx + r * a.zxp + (slot - RH));
          } else {
            *d2 = 0.f;
          }
        }
    };
    stage_rows(T - 1);
    cp_async_commit_wait_all();
    // per-lane row pointers of the caller-layout tensors at the LAST timestep (walked backwards)
    const float* yrow[2]; const float* xrow[2]; const float* dyrow[2]; float* dxrow[2];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      yrow[hf] = a.y + (size_t)(T - 2) * a.ys_t + (size_t)sq[hf] * a.ys_b + j0;          // h_{t-1} of step t = T-1
      xrow[hf] = a.x + (size_t)(T - 1) * a.xs_t + (size_t)sq[hf] * a.xs_b + j0;
      dyrow[hf] = a.dy ? a.dy + (size_t)(T - 1) * a.dys_t + (size_t)sq[hf] * a.dys_b + j0 : nullptr;
      dxrow[hf] = a.dx ? a.dx + (size_t)(T - 1) * a.dxs_t + (size_t)sq[hf] * a.dxs_b + j0 : nullptr;
    }
    // L2 prefetch of the h_{t-1} rows two steps ahead (16 lanes of warp 1, one row each)
    auto prefetch_y = [&](int tp) {
      if (warp != 1 || lane >= 16 || tp < 1 || (b0 + lane) >= B) return;
      l2_prefetch_bulk(a.y + (size_t)(tp - 1) * a.ys_t + (size_t)(b0 + lane) * a.ys_b, (uint32_t)(H * sizeof(float)));
    };
    prefetch_y(T - 1);
    prefetch_y(T - 2);
    __syncthreads();

    for (int t = T - 1; t >= 0; --t) {
      prefetch_step(t - 2);
      prefetch_y(t - 2);
      stage_rows(t - 1);                                 // rows of the NEXT (earlier) step, visible after this step's barriers
      const float* arow = Ar + (size_t)(t & 1) * 16 * AP;
      // A operand of the [z|zx|1]^T dPre product for this step: (m = slot g / g+8, k = sequence q / q+4) x two k-steps
      float wah[2][4], wal[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const float av[4] = {arow[(8 * ks + q) * AP + g], arow[(8 * ks + q) * AP + g + 8], arow[(8 * ks + q + 4) * AP + g],
                             arow[(8 * ks + q + 4) * AP + g + 8]};
        split4(av, wah[ks], wal[ks]);
      }
      float dz[KS][4];
#pragma unroll
      for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) dz[s][i] = 0.f;
      float sd[2][2][2];                                 // [P][e][hf]  sum_k dPre_k Dh_k
      float sx[2][2][2];                                 // [P][e][hf]  sum_k dPre_k Dx_k (x-side warps)
#pragma unroll
      for (int P = 0; P < 2; ++P) {
        float dpre[4][2][2];                             // [k][e][hf]
        float2 hp2[2], xv2[2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int o = (P * 2 + hf) * 64;
          float2 G[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) G[k] = __ldg(reinterpret_cast<const float2*>(gfrag + (size_t)k * qstride + o));
          const float2 ct2 = __ldg(reinterpret_cast<const float2*>(cfrag + o));
          float2 cp2 = make_float2(0.f, 0.f), dy2 = cp2;
          hp2[hf] = cp2; xv2[hf] = cp2;
          if (t > 0) cp2 = __ldg(reinterpret_cast<const float2*>(cfrag - cstep + o));
          else if (ok[hf][P] && a.c0) cp2 = __ldg(reinterpret_cast<const float2*>(a.c0 + (size_t)sq[hf] * H + j0 + PS * P));
          if (ok[hf][P]) {
            if (a.dy) {
              const float* dp = dyrow[hf] + PS * P;
              if (dy_vec) dy2 = __ldg(reinterpret_cast<const float2*>(dp));
              else { dy2.x = __ldg(dp); dy2.y = __ldg(dp + 1); }
            }
            if (t > 0) hp2[hf] = __ldg(reinterpret_cast<const float2*>(yrow[hf] + PS * P));
            else if (a.h0) hp2[hf] = __ldg(reinterpret_cast<const float2*>(a.h0 + (size_t)sq[hf] * H + j0 + PS * P));
          }
          if (xhalf[P] && sq[hf] < B) {
            const float* xp = xrow[hf] + PS * P;
            if (j0 + PS * P < I) xv2[hf].x = __ldg(xp);
            if (j0 + PS * P + 1 < I) xv2[hf].y = __ldg(xp + 1);
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
          // dPre of this half -> per-warp tile, accumulator layout: row = sequence, column = k*8 + unit-in-half
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<float2*>(myT + (g + 8 * hf) * TP + k * 8 + 2 * q) = make_float2(dpre[k][0][hf], dpre[k][1][hf]);
        }
        // ---- vector-multiplication terms: dh seed, dx seed, and the dDh / dDx sums (accumulators in tensor memory) ----
        {
          float gd[2][4];                                // [e][k] running sums of dPre_k * h_{t-1} over this lane's sequences
          float gx[2][4];                                // ... and of dPre_k * x_t (halves that hold input-side units)
          tmem_ld4(tbase + 32 + 8 * P, gd[0]);
          tmem_ld4(tbase + 36 + 8 * P, gd[1]);
          if (xhalf[P]) {                                // one wait covers both accumulator sets
            tmem_ld4(tbase + 48 + 8 * P, gx[0]);
            tmem_ld4(tbase + 52 + 8 * P, gx[1]);
          }
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) sd[P][e][hf] = sx[P][e][hf] = 0.f;
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 d2 = *reinterpret_cast<const float2*>(DhS + k * HP + j0 + PS * P);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              sd[P][0][hf] = fmaf(dpre[k][0][hf], d2.x, sd[P][0][hf]);
              sd[P][1][hf] = fmaf(dpre[k][1][hf], d2.y, sd[P][1][hf]);
              gd[0][k] = fmaf(dpre[k][0][hf], hp2[hf].x, gd[0][k]);
              gd[1][k] = fmaf(dpre[k][1][hf], hp2[hf].y, gd[1][k]);
            }
          }
          tmem_st4(tbase + 32 + 8 * P, gd[0]);
          tmem_st4(tbase + 36 + 8 * P, gd[1]);
          if (xhalf[P]) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                gx[0][k] = fmaf(dpre[k][0][hf], xv2[hf].x, gx[0][k]);
                gx[1][k] = fmaf(dpre[k][1][hf], xv2[hf].y, gx[1][k]);
              }
            tmem_st4(tbase + 48 + 8 * P, gx[0]);
            tmem_st4(tbase + 52 + 8 * P, gx[1]);
            if (a.dx) {                                  // dx seed, only when the caller wants dX
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 d2 = *reinterpret_cast<const float2*>(DxS + k * HP + j0 + PS * P);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                  sx[P][0][hf] = fmaf(dpre[k][0][hf], d2.x, sx[P][0][hf]);
                  sx[P][1][hf] = fmaf(dpre[k][1][hf], d2.y, sx[P][1][hf]);
                }
              }
            }
          }
        }
        // ---- partial dzc GEMM of this half (dPre registers are the A fragments) ----
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float av[4] = {dpre[k][0][0], dpre[k][0][1], dpre[k][1][0], dpre[k][1][1]};
          float ah[4], al[4];
          split4(av, ah, al);
#pragma unroll
          for (int s = 0; s < KS; ++s) {
            const float4 b = myB[((P * 4 + k) * KS + s) * 32];
            mma_3x(dz[s], ah, al, b.x, b.y, b.z, b.w);
          }
        }
        // ---- [z|zx|1]^T dPre of this half: B = dPre tile read back transposed (k = sequence, n = unit 8P+g of gate k) ----
        __syncwarp();
#pragma unroll
        for (int kk = 0; kk < 4; kk += 2) {              // two gates per TMEM round trip: two independent MMA chains
          // The tensor core adds into its fp32 accumulator with truncation (~3e-8 relative per add, one sign): a running sum
          // kept in the MMA accumulator over all timesteps and tiles of a CTA drifts linearly (1.2e-5 at 96 steps, measured
          // against fp64).  So each step's products go into a fresh accumulator and join the running sum with an FADD
          // (round to nearest); the running sum's TMEM load now also overlaps the MMAs.
          float acc[2][4], tmp[2][4], bh[2][4], bl[2][4];
          tmem_ld8(tbase + 4 * (P * 4 + kk), acc);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float* tc0 = myT + (kk + u) * 8 + g;
            const float b[4] = {tc0[q * TP], tc0[(q + 4) * TP], tc0[(q + 8) * TP], tc0[(q + 12) * TP]};
            split4(b, bh[u], bl[u]);
#pragma unroll
            for (int i = 0; i < 4; ++i) tmp[u][i] = 0.f;
          }
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int u = 0; u < 2; ++u) mma_tf32(tmp[u], wal[ks], bh[u][2 * ks], bh[u][2 * ks + 1]);
#pragma unroll
            for (int u = 0; u < 2; ++u) mma_tf32(tmp[u], wah[ks], bl[u][2 * ks], bl[u][2 * ks + 1]);
#pragma unroll
            for (int u = 0; u < 2; ++u) mma_tf32(tmp[u], wah[ks], bh[u][2 * ks], bh[u][2 * ks + 1]);
          }
          tmem_wait_ld();
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[u][i] += tmp[u][i];
          tmem_st8(tbase + 4 * (P * 4 + kk), acc);
        }
        __syncwarp();                                    // tile reads done before the next half overwrites it
      }
      {
        float* pw = Pz + (size_t)warp * 16 * PP;
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          *reinterpret_cast<float2*>(pw + g * PP + 8 * s + 2 * q) = make_float2(dz[s][0], dz[s][1]);
          *reinterpret_cast<float2*>(pw + (g + 8) * PP + 8 * s + 2 * q) = make_float2(dz[s][2], dz[s][3]);
        }
      }
      gfrag -= gstep; cfrag -= cstep;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        yrow[hf] -= a.ys_t; xrow[hf] -= a.xs_t;
        if (a.dy) dyrow[hf] -= a.dys_t;
      }
      cp_async_commit_wait_all();                        // this thread's staged element of step t-1 has landed
      __syncthreads();
      // ---- fixed-order sum over warps -> Dz rows ----
      for (int idx = tid; idx < 16 * 8 * KS * 4; idx += nthreads) {
        const int el = idx >> 2, part = idx & 3, seq = el / (8 * KS), slot = el - seq * (8 * KS);
        float s = 0.f;
        for (int w = part; w < NW; w += 4) s += Pz[((size_t)w * 16 + seq) * PP + slot];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0) {
          Dz[seq * PP + slot] = s;
        }
      }
      __syncthreads();
      // ---- dzx rows of this step -> HBM, in the row order of x (a.dz_bt: [B,T] else [T,B]), pad columns zero ----
      for (int idx = tid; idx < 16 * a.zxp; idx += nthreads) {
        const int seq = idx / a.zxp, r = idx - seq * a.zxp;
        if (b0 + seq < B) {
          const size_t row = a.dz_bt ? (size_t)(b0 + seq) * T + t : (size_t)t * B + b0 + seq;
          a.dzc[row * a.zxp + r] = r < RX ? Dz[seq * PP + RH + r] : 0.f;
        }
      }
      // ---- dh_{t-1} = dz A^T + sum_k dPre_k Dh_k ; dx_t = dzx Ux^T + sum_k dPre_k Dx_k ----
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
      if (xwarp && a.dx) {
        float ax[2][4];
#pragma unroll
        for (int P = 0; P < 2; ++P) {
          ax[P][0] = sx[P][0][0]; ax[P][1] = sx[P][1][0]; ax[P][2] = sx[P][0][1]; ax[P][3] = sx[P][1][1];
        }
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const int s0 = 8 * s + q, s1 = s0 + 4;
          const float av[4] = {Dz[g * PP + s0], Dz[(g + 8) * PP + s0], Dz[g * PP + s1], Dz[(g + 8) * PP + s1]};
          float ah[4], al[4];
          split4(av, ah, al);
#pragma unroll
          for (int P = 0; P < 2; ++P) {                  // B = Ux^T restricted to the zx slots: (k = slot, n = g <-> unit 8P+g)
            const int j = ubase + PS * P + g;
            const float b0 = (j < I && s0 >= RH && s0 < RH + RX) ? UxS[j * RX + (s0 - RH)] : 0.f;
            const float b1 = (j < I && s1 >= RH && s1 < RH + RX) ? UxS[j * RX + (s1 - RH)] : 0.f;
            const float b0h = tf32_rna(b0), b1h = tf32_rna(b1);
            mma_3x(ax[P], ah, al, b0h, b1h, b0 - b0h, b1 - b1h);
          }
        }
#pragma unroll
        for (int P = 0; P < 2; ++P)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int j = j0 + PS * P;
            if (sq[hf] < B) {
              float* o = dxrow[hf] + PS * P;
              if (j < I) o[0] = ax[P][2 * hf];
              if (j + 1 < I) o[1] = ax[P][2 * hf + 1];
            }
          }
      }
      if (a.dx) { dxrow[0] -= a.dxs_t; dxrow[1] -= a.dxs_t; }
      // ---- dA += Hprev^T dz: A = (m = g: unit ju, m = g+8: unit ju+1; k = sequence q / q+4) read transposed from the rows
      //      this warp touched in phase 1 (L1/L2 hits), B = the reduced dz rows (k = sequence, n = slot).  dUx = X^T dzx
      //      is NOT formed here: only the warps that own input-side units could do it, and they are the critical path of
      //      every barrier; the dzx rows go to HBM (32 B per sequence-step) for dux_rows_kernel ----
      {
        const int ju = ubase + PS * (g >> 2) + 2 * (g & 3);
        const bool uin = ju < H;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const int rr0 = 8 * ks + q, rr1 = rr0 + 4;
          const bool v0 = (b0 + rr0) < B, v1 = (b0 + rr1) < B;
          float2 h0v = make_float2(0.f, 0.f), h1v = h0v;
          if (uin) {
            if (t > 0) {
              const float* yb = a.y + (size_t)(t - 1) * a.ys_t + (size_t)b0 * a.ys_b + ju;
              if (v0) h0v = __ldg(reinterpret_cast<const float2*>(yb + (size_t)rr0 * a.ys_b));
              if (v1) h1v = __ldg(reinterpret_cast<const float2*>(yb + (size_t)rr1 * a.ys_b));
            } else if (a.h0) {
              if (v0) h0v = __ldg(reinterpret_cast<const float2*>(a.h0 + (size_t)(b0 + rr0) * H + ju));
              if (v1) h1v = __ldg(reinterpret_cast<const float2*>(a.h0 + (size_t)(b0 + rr1) * H + ju));
            }
          }
          const float hv[4] = {h0v.x, h0v.y, h1v.x, h1v.y};
          float hh[4], hl[4];
          split4(hv, hh, hl);
#pragma unroll
          for (int s = 0; s < NZ; ++s) {                 // only the dz slots (n-tiles below NZ) feed dA
            float ag[4], tg[4] = {0.f, 0.f, 0.f, 0.f};   // fresh MMA accumulator per step (see the dW block above)
            tmem_ld4(tbase + 64 + 4 * s, ag);
            const float d0 = Dz[rr0 * PP + 8 * s + g], d1 = Dz[rr1 * PP + 8 * s + g];
            const float b0h = tf32_rna(d0), b1h = tf32_rna(d1);
            mma_3x(tg, hh, hl, b0h, b1h, d0 - b0h, d1 - b1h);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 4; ++i) ag[i] += tg[i];
            tmem_st4(tbase + 64 + 4 * s, ag);
          }
        }
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

  // ---- this CTA's partial: accumulators out of tensor memory ----
  tmem_wait_st();
  const GradLayout L(I, H, RX, RH);
  float* Pout = a.partial + (size_t)blockIdx.x * L.total;
  // [z|zx|1]^T dPre, n-tile (k, P): C fragment (slot g, unit 8P+2q / +1), (slot g+8, same)
#pragma unroll
  for (int P = 0; P < 2; ++P)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float acc[4];
      tmem_ld4(tbase + 4 * (P * 4 + k), acc);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int slot = g + 8 * (i >> 1), j = j0 + PS * P + (i & 1);
        if (j >= H) continue;
        const size_t row = (size_t)k * H + j;
        if (slot < RH) Pout[L.oBm + row * RH + slot] = acc[i];
        else if (slot < RH + RX) Pout[L.oVx + row * RX + (slot - RH)] = acc[i];
        else if (slot == RH + RX) Pout[L.oBias + row] = acc[i];
      }
    }
  // dDh / dDx: sum over the eight g lanes (sequences), lanes g == 0 store units j0 + PS*P + e
#pragma unroll
  for (int P = 0; P < 2; ++P)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float gd[4], gx[4];
      tmem_ld4(tbase + 32 + 8 * P + 4 * e, gd);
      tmem_ld4(tbase + 48 + 8 * P + 4 * e, gx);
      tmem_wait_ld();
      const int j = j0 + PS * P + e;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v = gd[k], w = gx[k];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          v += __shfl_xor_sync(0xffffffffu, v, o);
          w += __shfl_xor_sync(0xffffffffu, w, o);
        }
        if (g == 0 && j < H) Pout[L.oDh + k * H + j] = v;
        if (g == 0 && j < I) Pout[L.oDx + k * I + j] = w;
      }
    }
  // dA: C fragment (m = g [+8] <-> unit ju [+1], n = slot 8s + 2q [+1]); the dUx slice of the partial stays unwritten
  // (launch_bwd_fused overwrites dUx with dux_rows_kernel's result after the reduce)
  {
    const int ju = ubase + PS * (g >> 2) + 2 * (g & 3);
#pragma unroll
    for (int s = 0; s < NZ; ++s) {
      float ag[4];
      tmem_ld4(tbase + 64 + 4 * s, ag);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = ju + (i >> 1), slot = 8 * s + 2 * q + (i & 1);
        if (j < H && slot < RH) Pout[L.oA + (size_t)j * RH + slot] = ag[i];
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tslot), "r"(tmem_cols));
  }
}

// ------------------------------------------------------------------------------------------------- //
// dUx = X^T dZX over all T*B rows: a streaming pass over x (read once, I floats per row) and the dzx rows the fused
// kernel wrote (zxp floats per row), both in x's row order.  Same staging as xproj_small_kernel (64-row tiles,
// contiguous spans copied with 16-byte loads); thread = (input unit j, 4 slots), accumulators in registers, one
// partial per block, summed by dux_reduce_kernel in block order.
// ------------------------------------------------------------------------------------------------- //
constexpr int kDuxRows = 64, kDuxThreads = 256, kDuxItems = 8;      // I * zxp / 4 <= 1024 work items
constexpr int kDuxStages = 3;
struct DuxArgs {
  const float* x; long long xs_t, xs_b;
  const float* dzx;                        // [T*B, zxp]
  float* pbuf;                             // [2 * gridDim.x, I * zxp]: two partials per block (row halves of its tiles)
  int T, B, I, zxp, contiguous;            // contiguous: rows are one span in memory (either order); else (t, b) strides
  int stages;                              // 3: cp.async pipeline (contiguous, 16-byte aligned x), 1: plain staging
};

// block = 256 threads: thread = (work item (input unit j, 4 slots), half of the tile's rows).  smem per stage: xs[64][I] | ds[64][zxp]
static __global__ void __launch_bounds__(kDuxThreads) dux_rows_kernel(const DuxArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int I = a.I, zxp = a.zxp, NQ = zxp >> 2, nitems = I * NQ;
  const size_t stage_floats = (size_t)kDuxRows * (I + zxp);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = tid >> 7, t7 = tid & 127;           // rows [32 half, 32 half + 32) of every tile
  float4 acc[kDuxItems];
#pragma unroll
  for (int it = 0; it < kDuxItems; ++it) acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long nrows = (long long)a.T * a.B;
  const long long step = (long long)gridDim.x * kDuxRows;

  auto compute = [&](const float* xs, const float* ds, int nr) {
    const int r0 = 32 * half, r1 = min(nr, r0 + 32);
#pragma unroll
    for (int it = 0; it < kDuxItems; ++it) {
      const int item = t7 + it * 128;
      if (item >= nitems) break;
      const int j = item / NQ, nq = item - j * NQ;
      float4 s = acc[it];
#pragma unroll 4
      for (int r = r0; r < r1; ++r) {
        const float xv = xs[r * I + j];
        const float4 d = *reinterpret_cast<const float4*>(ds + r * zxp + 4 * nq);
        s.x = fmaf(xv, d.x, s.x); s.y = fmaf(xv, d.y, s.y); s.z = fmaf(xv, d.z, s.z); s.w = fmaf(xv, d.w, s.w);
      }
      acc[it] = s;
    }
  };

  if (a.stages >= 3) {
    auto issue = [&](long long row0, int stage) {
      if (row0 < nrows) {
        const long long left = nrows - row0;
        const int nr = (int)(left < kDuxRows ? left : kDuxRows);
        float* xs = smem + (size_t)stage * stage_floats;
        float* ds = xs + (size_t)kDuxRows * I;
        const float* src = a.x + row0 * I;
        const int n = nr * I, n4 = n & ~3;
        for (int e = tid * 4; e < n4; e += 4 * kDuxThreads)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(xs + e)), "l"(src + e) : "memory");
        for (int t1 = n4 + tid; t1 < n; t1 += kDuxThreads) xs[t1] = __ldg(src + t1);
        const float* dsrc = a.dzx + row0 * zxp;
        for (int e = tid * 4; e < nr * zxp; e += 4 * kDuxThreads)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(ds + e)), "l"(dsrc + e) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    long long row0 = (long long)blockIdx.x * kDuxRows;
    issue(row0, 0);
    issue(row0 + step, 1);
    for (int i = 0; row0 < nrows; ++i, row0 += step) {
      issue(row0 + 2 * step, (i + 2) % 3);              // that stage was consumed in iteration i-1 (barrier below)
      asm volatile("cp.async.wait_group 2;" ::: "memory");
      __syncthreads();
      const long long left = nrows - row0;
      const float* xs = smem + (size_t)(i % 3) * stage_floats;
      compute(xs, xs + (size_t)kDuxRows * I, (int)(left < kDuxRows ? left : kDuxRows));
      __syncthreads();
    }
  } else {
    float* xs = smem;
    float* ds = smem + (size_t)kDuxRows * I;
    for (long long row0 = (long long)blockIdx.x * kDuxRows; row0 < nrows; row0 += step) {
      __syncthreads();
      const long long left = nrows - row0;
      const int nr = (int)(left < kDuxRows ? left : kDuxRows);
      if (a.contiguous) {
        const float* src = a.x + row0 * I;
        for (int t1 = tid; t1 < nr * I; t1 += kDuxThreads) xs[t1] = __ldg(src + t1);
      } else {
        for (int rr = warp; rr < nr; rr += kDuxThreads / 32) {
          const long long row = row0 + rr, t = row / a.B, b = row % a.B;
          const float* src = a.x + t * a.xs_t + b * a.xs_b;
          for (int jj = lane; jj < I; jj += 32) xs[rr * I + jj] = __ldg(src + jj);
        }
      }
      {                                                  // dzx rows: contiguous, zxp % 4 == 0
        const float4* src = reinterpret_cast<const float4*>(a.dzx + row0 * zxp);
        for (int e = tid; e < nr * NQ; e += kDuxThreads) reinterpret_cast<float4*>(ds)[e] = __ldg(src + e);
      }
      __syncthreads();
      compute(xs, ds, nr);
    }
  }
  float* out = a.pbuf + ((size_t)blockIdx.x * 2 + half) * I * zxp;
#pragma unroll
  for (int it = 0; it < kDuxItems; ++it) {
    const int item = t7 + it * 128;
    if (item >= nitems) break;
    *reinterpret_cast<float4*>(out + (size_t)item * 4) = acc[it];            // item = j * NQ + nq  ->  offset j * zxp + 4 nq
  }
}

// dUx[j, r] = sum over blocks of pbuf[block][j * zxp + r]: 4 outputs per 256-thread block, each summed by 64 threads over
// interleaved slices of the block list (independent loads in flight), then across the slices in slice order (fixed order)
static __global__ void __launch_bounds__(256) dux_reduce_kernel(const float* __restrict__ pbuf, int nblocks, int I, int RX,
                                                                int zxp, float* __restrict__ dUx) {
  __shared__ float sl[64][5];
  const int e = threadIdx.x & 3, slice = threadIdx.x >> 2;
  const int i = blockIdx.x * 4 + e;                    // output index j * RX + r
  const bool live = i < I * RX;
  const int j = live ? i / RX : 0, r = live ? i - j * RX : 0;
  const size_t stride = (size_t)I * zxp, off = (size_t)j * zxp + r;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (live) {
    int b = slice;
    for (; b + 192 < nblocks; b += 256) {
      s0 += pbuf[(size_t)b * stride + off];
      s1 += pbuf[(size_t)(b + 64) * stride + off];
      s2 += pbuf[(size_t)(b + 128) * stride + off];
      s3 += pbuf[(size_t)(b + 192) * stride + off];
    }
    for (; b < nblocks; b += 64) s0 += pbuf[(size_t)b * stride + off];
  }
  sl[slice][e] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (slice == 0 && live) {
    float t = 0.f;
#pragma unroll 8
    for (int q = 0; q < 64; ++q) t += sl[q][e];
    dUx[i] = t;
  }
}

int launch_bwd_fused(const SeqBwdFusedArgs& a, const GradOut& out, void* workspace, cudaStream_t st);
long long bwd_fused_workspace_floats(int T, int B, int I, int H, int RX, int RH);
bool bwd_fused_fits(int I, int H, int RX, int RH);

}  // namespace vmlmf
