/* vmlmf_b200.h -- C ABI of libvmlmf_b200.so (sm_100a).
 *
 * Drop-in boundary for the VMLMF compressed-LSTM hot path.  The reference
 * (snudm-starlab/VMLMF) has no FFI / plugin interface: its boundary is the Python
 * nn.Module API.  Each entry point below replaces a piece of reference Python that runs
 * once per timestep (or its autograd replay); the file:line being replaced is cited.
 * "V/" = rnn_compression_factorization_vmlmf/src/ in the reference tree.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *    (PyTorch); the library never allocates, frees, retains pointers or synchronises.
 *  - all tensors fp32, innermost dimension contiguous; x / y / dy / dx carry explicit
 *    element strides for their (time, batch) axes so batch-first (MyLSTM) and time-major
 *    (MyVMLSTM) callers share one kernel.
 *  - `stream` is a cudaStream_t passed as void*.  Entry points are re-entrant (autograd
 *    calls backward from its own thread); there is no global mutable state.
 *  - return 0 on success, a negative VMLMF_E* code for bad arguments / unsupported shapes,
 *    a positive cudaError_t for a launch failure.  vmlmf_strerror() maps either to text.
 *
 * Canonical parameters (see DESIGN.md "canonical recurrence"; gate order k = i,f,o,n):
 *   Ux[I,RX]  Vx[4H,RX]  Dx[4,I]   input side  (Dx = dia_x - diag(Ux Vx_k^T))
 *   A [H,RH]  Bm[4H,RH]  Dh[4,H]   hidden side (Dh = dia_h - diag(A  Bm_k^T))
 *   bias[4H] = b_x + b_h
 *   pre_t[b,k,j] = (x_t Ux) Vx[kH+j,:] + [j<I] x_t[b,j] Dx[k,j]
 *                + (h_{t-1} A) Bm[kH+j,:] + h_{t-1}[b,j] Dh[k,j] + bias[kH+j]
 */
#ifndef VMLMF_B200_H_
#define VMLMF_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMLMF_ABI_VERSION 5

enum {
  VMLMF_OK = 0,
  VMLMF_EINVAL = -1,       /* null pointer / non-positive size / bad stride            */
  VMLMF_ESHAPE = -2,       /* hidden_size < input_size: the reference raises TypeError
                              here too (V/models/vmlmf.py:92-94,117)                    */
  VMLMF_EUNSUPPORTED = -3, /* shape outside every compiled regime                       */
  VMLMF_EWORKSPACE = -4,   /* workspace too small (see vmlmf_seq_plan)                  */
  VMLMF_EPLAN = -5         /* plan does not match the arguments                         */
};

/* regimes (vmlmf_plan.path) */
enum {
  VMLMF_PATH_R1 = 1,  /* persistent SIMT, factors register-resident, thread = hidden unit      */
  VMLMF_PATH_G = 2,   /* generic: time-parallel XP GEMM + one fused launch per timestep        */
  VMLMF_PATH_R1M = 3, /* persistent warp-MMA (mma.sync 3xTF32) recurrence, CTA = 16 sequences;
                         needs H % 4 == 0, H <= 256, RH <= 16, RH + RX + 1 <= 32               */
  VMLMF_PATH_R2 = 4,  /* persistent tcgen05 recurrence for every other shape (large H, high ranks):
                         a thread-block cluster owns a 128-sequence tile for all T steps, the hidden
                         units are split over the cluster's CTAs, operands arrive as TMA tiles     */
  VMLMF_PATH_R3 = 5   /* the same recurrence for small batches (B <= 32, e.g. the LM at the reference's
                         20 streams): one group of ceil(H/8) CTAs, each keeps its rows of the factors
                         resident in shared memory for all T steps; only activations move per step  */
};

typedef struct vmlmf_plan {
  int path;                  /* VMLMF_PATH_*                                            */
  int zx_pitch;              /* row pitch (floats) of zx[T*B, zx_pitch]  (x_t Ux)       */
  int z_pitch;               /* row pitch (floats) of z [T*B, z_pitch]   (h_{t-1} A)    */
  int xp_cols;               /* PATH_G: columns of xp[T*B, xp_cols] (=4H), else 0       */
  long long fwd_workspace_bytes;
  long long bwd_workspace_bytes;
  long long gates_bytes;     /* size of the saved `gates` buffer (PATH_R1M pads it to whole
                                16-sequence x 16-unit fragments and stores it fragment-major;
                                the other paths use [T,B,4,H]); opaque to the caller          */
  long long cs_bytes;        /* size of the saved `cs` buffer, same remark                     */
  int reserved[8];
} vmlmf_plan;

int vmlmf_abi_version(void);
const char* vmlmf_strerror(int code);

/* Which regime runs these sizes, and how big the caller-owned scratch buffers must be. */
int vmlmf_seq_plan(int T, int B, int I, int H, int RX, int RH, vmlmf_plan* plan);

/* K0 / K5: the loop-invariant diagonal corrections, hoisted out of the time loop (the reference recomputes them
 * in every step: V/models/vmlmf.py:102-106, vmlmf_lm.py:250-255) and their chain rule.
 *   D[k,j] = dia[j] - sum_r u[j,r] v[kH+j,r]          u[n,R]  v[4H,R]  dia[n]  D[4,n]   (n = I or H, n <= H)
 *   du[j,r] = -sum_k dD[k,j] v[kH+j,r]    dv[kH+j,r] = -dD[k,j] u[j,r] (0 for j >= n)    ddia[j] = sum_k dD[k,j]   */
int vmlmf_diag_fwd(const float* u, const float* v, const float* dia, float* D, int n, int H, int R,
                   void* stream);
int vmlmf_diag_bwd(const float* u, const float* v, const float* dD, float* du, float* dv, float* ddia,
                   int n, int H, int R, void* stream);

/* K0 / K5 for a whole plain cell (MyVMLMFCell, MyVMLSTM) in ONE launch each way.
 * forward : Dx[4,I], Dh[4,H] as vmlmf_diag_fwd, and bias[4H] = b_x + b_h   (V/models/vmlmf.py:102-110).
 * backward: turns the canonical gradients that vmlmf_seq_bwd wrote into the reference parameters' gradients IN PLACE
 *           (dUx -> du_x, dVx -> dv_x, dA -> du_h, dBm -> dv_h get the diagonal-correction chain rule subtracted),
 *           writes ddia_x[I], ddia_h[H] and db_h[4H] (a copy of dbias: b_x and b_h must not share a gradient buffer). */
int vmlmf_pack_plain_fwd(const float* u_x, const float* v_x, const float* dia_x, const float* u_h, const float* v_h,
                         const float* dia_h, const float* b_x, const float* b_h, float* Dx, float* Dh, float* bias,
                         int I, int H, int RX, int RH, void* stream);
int vmlmf_pack_plain_bwd(const float* u_x, const float* v_x, const float* u_h, const float* v_h, const float* dDx,
                         const float* dDh, float* dUx, float* dVx, float* dA, float* dBm, float* ddia_x,
                         float* ddia_h, const float* dbias, float* db_h, int I, int H, int RX, int RH, void* stream);

/* K1: zx[t,b,:] = x[t,b,:] Ux for every timestep at once (time-parallel half of
 * `torch.matmul(x, self.u_x)`, V/models/vmlmf.py:98, vmlmf_group.py:98, vmlmf_lm.py:246).
 * zx is [T*B, plan.zx_pitch], pad columns written as 0.                               */
int vmlmf_xproj_fwd(const float* x, long long xs_t, long long xs_b, const float* Ux,
                    float* zx, int T, int B, int I, int RX, int zx_pitch, void* stream);

/* K2: the whole time loop of one layer -- replaces `for t in range(seqlen): h,c = cell(...)`
 * (V/models/vmlmf.py:308-310 with the cell body :78-125; vmlmf_group.py:85-155;
 * vmlmf_lm.py:272-280 with lstm_step :222-269).
 *   h0,c0      [B,H] or NULL (= zeros, MyLSTM.forward :302-303)
 *   y          h_t for every t, strides (ys_t, ys_b); may be NULL when nothing is saved and the plan
 *              is PATH_R1 / PATH_R1M: a caller that consumes only the last step (Net.forward,
 *              V/models/vmlmf.py:354-355) skips the [T,B,H] write and reads hT instead
 *   hT,cT      [B,H] final state
 *   gates,cs,z saved for backward: plan.gates_bytes / plan.cs_bytes / T*B*z_pitch floats, written by
 *              this call and read back only by vmlmf_seq_bwd (layout is private to the path);
 *              all three NULL = inference (nothing saved)
 *   PATH_R1M additionally needs y rows 8-byte aligned (ys_t, ys_b even)
 *   workspace  plan.fwd_workspace_bytes (may be NULL when that is 0)                   */
int vmlmf_seq_fwd(const vmlmf_plan* plan, const float* x, long long xs_t, long long xs_b,
                  const float* zx, const float* Ux, const float* Vx, const float* Dx,
                  const float* A, const float* Bm, const float* Dh, const float* bias,
                  const float* h0, const float* c0, float* y, long long ys_t, long long ys_b,
                  float* hT, float* cT, float* gates, float* cs, float* z, void* workspace,
                  int T, int B, int I, int H, int RX, int RH, void* stream);

/* K3: fused backward-through-time of the same loop (replaces the autograd replay of the
 * per-step ATen ops).  Accumulates every factor gradient on chip; no [H,4H] matrix or its
 * gradient is formed.  dy (strides dys_t,dys_b), dhT, dcT may each be NULL (= zeros).
 * Outputs: dx (strides dxs_t,dxs_b; may be NULL), dh0,dc0 [B,H] (may be NULL) and the
 * canonical-parameter gradients dUx,dVx,dDx,dA,dBm,dDh,dbias (overwritten, not added).
 * The sum over CTAs is a fixed-order tree: results are bit-reproducible run to run.
 * workspace: plan.bwd_workspace_bytes (per-CTA gradient partials; on PATH_R1M also the
 * dzx rows and the partials of the streaming dUx = X^T dZX pass that follows the
 * recurrence kernel).  x is read twice on that path: keep it valid until the call's work
 * on `stream` has finished, like every other argument.                                   */
int vmlmf_seq_bwd(const vmlmf_plan* plan, const float* x, long long xs_t, long long xs_b,
                  const float* zx, const float* Ux, const float* Vx, const float* Dx,
                  const float* A, const float* Bm, const float* Dh,
                  const float* h0, const float* c0, const float* y, long long ys_t,
                  long long ys_b, const float* gates, const float* cs, const float* z,
                  const float* dy, long long dys_t, long long dys_b, const float* dhT,
                  const float* dcT, float* dx, long long dxs_t, long long dxs_b, float* dh0,
                  float* dc0, float* dUx, float* dVx, float* dDx, float* dA, float* dBm,
                  float* dDh, float* dbias, void* workspace, int T, int B, int I, int H,
                  int RX, int RH, void* stream);

/* fp32-accurate GEMM on the tensor cores (tcgen05, 3xTF32, TMA):  C[M,N] (+)= A[M,K] B[N,K]^T (+ bias[N]).
 * Row pitches in floats; A and B need 16-byte aligned rows (pitch % 4 == 0, base % 16 == 0) -- other operands
 * run on the SIMT fallback.  Used for the LM vocabulary projection around the path
 * (V/models/vmlmf_lm.py:341-361, `torch.addmm(self.b, x, self.w.t())`) and its backward.  K contractions longer
 * than a tile wave are split and summed in a fixed order; `workspace` (workspace_bytes >= 16*M*N*4 lets every
 * split configuration through, 0 disables splitting) holds the partials.                                      */
int vmlmf_gemm_nt(const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc,
                  const float* bias, int M, int N, int K, int accumulate, void* workspace,
                  long long workspace_bytes, void* stream);

/* C[M,N] (+)= At[K,M]^T Bt[K,N]: both operands stored with the contraction index as the ROW (activations [rows, features]);
 * the tensor cores read them as MN-major operands, no transposed copy is made.  This is the weight-gradient product
 * dW = dY^T X of a linear layer (V/models/vmlmf_lm.py:357 under autograd) and of the low-rank factors.  Row pitches in
 * floats, rows 16-byte aligned; returns VMLMF_EUNSUPPORTED when an operand misses that.  Long contractions are split and
 * summed in a fixed order; `workspace` as for vmlmf_gemm_nt.                                                          */
int vmlmf_gemm_tn(const float* At, long long lda, const float* Bt, long long ldb, float* C, long long ldc,
                  int M, int N, long long K, int accumulate, void* workspace, long long workspace_bytes, void* stream);

/* ---- callers either side of the recurrence (SURVEY.md 8 rows f1 / f2 and the Net head, a4) ---------------- */

/* Softmax-NLL over rows of scores[rows, C] (row pitch ld floats), labels int64 in [0, C):
 *   lse[r]  = log sum_c exp(scores[r,c])            (kept for backward)
 *   *loss   = scale * sum_r (lse[r] - scores[r, labels[r]])      summed in a fixed order
 * scale = 1/rows is F.cross_entropy's mean (V/train_test/train.py:63); scale = batch/rows is the LM loss
 * `torch.mean(-log(p[y]) * batch_size)` (V/train_test/lm_test.py:140-153) without its exp() overflow.
 * workspace: vmlmf_softmax_nll_workspace_bytes(rows, C).                                                       */
long long vmlmf_softmax_nll_workspace_bytes(long long rows, int C);
int vmlmf_softmax_nll_fwd(const float* scores, long long ld, const long long* labels, float* lse, float* loss,
                          float scale, void* workspace, long long rows, int C, void* stream);
/* dscores[r,c] = (softmax(scores[r,:])[c] - [c == labels[r]]) * scale * (dloss ? *dloss : 1).  dloss is a device
 * scalar (the upstream gradient) or NULL; dscores (row pitch ldd) may alias scores.                           */
int vmlmf_softmax_nll_bwd(const float* scores, long long ld, const long long* labels, const float* lse,
                          const float* dloss, float scale, float* dscores, long long ldd, long long rows, int C,
                          void* stream);

/* The classifier head of Net, `self.lin(y[:, -1])` with lin = nn.Linear(H_last, 18) (V/models/vmlmf.py:345-347,
 * :354-355): out[B,N] = h[B,K] W[N,K]^T + bias[N] for N <= 32, K <= 1024 (h row pitch ldh, out dense).
 * Backward: dh[B,K] (pitch lddh, may be NULL), dW[N,K], db[N] (may be NULL) from dout[B,N]; per-block partial sums
 * are reduced in block order.  workspace: vmlmf_head_bwd_workspace_bytes(B, K, N).                             */
int vmlmf_head_fwd(const float* h, long long ldh, const float* W, const float* bias, float* out, int B, int K, int N,
                   void* stream);
long long vmlmf_head_bwd_workspace_bytes(int B, int K, int N);
int vmlmf_head_bwd(const float* h, long long ldh, const float* W, const float* dout, float* dh, long long lddh,
                   float* dW, float* db, void* workspace, int B, int K, int N, void* stream);

/* Optimizer steps on flat buffers of n floats (parameters, gradients and moments each one contiguous bucket --
 * the gradient bucket is the one the data-parallel all-reduce already uses).
 * Adam exactly as torch.optim.Adam(lr) (V/train_test/train.py:47,65): the step count t (>= 1, already incremented)
 * is read from the device scalar step_dev (float; lets a CUDA graph replay advance it) or, if NULL, from `step`. */
int vmlmf_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                    float eps, const float* step_dev, int step, void* stream);
/* LM input side: Embed (index -> row of W[V,E], V/models/vmlmf_lm.py:33-51) and the dropout applied to it (:436) in one pass:
 *   out[r, c] = W[tok[r], c] * (mask ? mask[r, c] * scale : 1)   c < E;   0 for E <= c < ldo
 * out has row pitch ldo (>= E): with ldo % 4 == 0 the result is directly the TMA A operand of the x-projection GEMM (rows of
 * E = 650 floats are not 16-byte multiples).  mask: uint8 [rows, E], 1 = keep, or NULL (eval / p = 0); scale = 1/(1-p).
 * An out-of-range token yields NaN rows (the eager indexing this replaces device-asserts).                              */
int vmlmf_embed_dropout_fwd(const long long* tok, const float* W, const unsigned char* mask, float scale, float* out,
                            long long ldo, long long rows, int E, int V, void* stream);

/* Data-parallel Adam step with the gradient all-reduce FUSED in, over NVLink peer memory (one-shot all-reduce): peer_grads is
 * a device array of `world` pointers, peer_grads[r] = rank r's flat gradient bucket of n floats mapped into this process
 * (torch symmetric memory / cudaIpc).  Every rank adds the buckets in rank order (bit-identical replicas), multiplies by
 * `scale` (1/world for the batch-mean HAR loss, V/train_test/train.py:63) and applies Adam as vmlmf_adam_step does.  The
 * caller brackets the call with cross-rank barriers on the stream (all buckets written before / all read after).  Meant for
 * the small factor-gradient buckets (latency-bound); large buckets (the LM's 67.7 MB) belong on ncclAllReduce.            */
int vmlmf_p2p_adam_step(float* p, float* m, float* v, const float* const* peer_grads, int world, long long n, float scale,
                        float lr, float beta1, float beta2, float eps, const float* step_dev, int step, void* stream);
/* clip_grad_norm_(max_norm) then `param -= lr * param.grad` (V/train_test/lm_test.py:203-209) in two launches:
 * coef = min(1, max_norm / (||g||_2 + 1e-6)) (max_norm <= 0: no clipping); scale_grads != 0 also writes g *= coef
 * back like clip_grad_norm_ does.  norm_out (device scalar, may be NULL) receives ||g||_2.
 * workspace: vmlmf_sgd_clip_workspace_bytes(n).                                                                */
long long vmlmf_sgd_clip_workspace_bytes(long long n);
int vmlmf_sgd_clip_step(float* p, float* g, long long n, float lr, float max_norm, int scale_grads,
                        float* norm_out, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VMLMF_B200_H_ */
