import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vmlmf_b200 import _lib
lib = _lib.lib()
dev = "cuda:0"
def run(K, M, N, lda=None, ldb=None):
    lda = lda or M; ldb = ldb or N
    g = torch.Generator(device=dev).manual_seed(K + M + N)
    At = torch.randn(K, lda, device=dev, generator=g)
    Bt = torch.randn(K, ldb, device=dev, generator=g)
    Cc = torch.zeros(M, N, device=dev)
    ws = torch.empty(64 * M * N, device=dev)
    rc = lib.vmlmf_gemm_tn(C.c_void_p(At.data_ptr()), lda, C.c_void_p(Bt.data_ptr()), ldb, C.c_void_p(Cc.data_ptr()), N, M, N, K, 0,
                           C.c_void_p(ws.data_ptr()), ws.numel() * 4, None)
    torch.cuda.synchronize()
    ref = (At[:, :M].double().t() @ Bt[:, :N].double())
    err = ((Cc.double() - ref).norm() / ref.norm()).item()
    print(f"K={K} M={M} N={N} lda={lda} ldb={ldb} rc={rc} rel err {err:.3e}")
    return Cc, ref
run(32, 128, 128)
run(8, 128, 128)
run(64, 128, 128)
run(700, 2688, 300, 2688, 300)
run(4096, 256, 64)
c, r = run(32, 32, 32)
print(c[:4, :4]); print(r[:4, :4].float())
