#!/usr/bin/env python
"""Debug aid: builds a -DVMLMF_R2_TRACE copy of the library, runs one R2 forward and prints the per-phase cycle trace of
CTA 0 (events: 1/2 producer Z begin/end, 3/4 producer G, 10/11 MMA Z first issue / commit, 12/13 MMA G, 20/21 epilogue Z,
30/31/32 exchange, 22/23/24 epilogue G wait-begin / wait-end / chunk done).  usage: trace_r2.py B T I H RX RH"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
csrc = os.path.join(ROOT, "vmlmf_b200", "csrc")
out = "/tmp/libvmlmf_trace.so"
srcs = ["vmlmf_api.cu", "seq_r1_rx4.cu", "seq_r1_rx8.cu", "seq_r1_rx16.cu", "seq_mma.cu", "seq_bwd_mma.cu", "seq_bwd_fused.cu", "seq_r2.cu", "seq_r3.cu"]
subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                       "--expt-relaxed-constexpr", "-DVMLMF_R2_TRACE", "-shared", "-o", out] + srcs, cwd=csrc)
import torch
from vmlmf_b200 import _lib
_lib.LIB_PATH = out
from vmlmf_b200.functional import vmlmf_sequence
B, T, I, H, RX, RH = [int(v) for v in sys.argv[1:7]]
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
sc = 0.05 if H >= 300 else 0.1
r = lambda *s: torch.randn(*s, device=dev, generator=g) * sc
canon = [r(I, RX), r(4 * H, RX), r(4, I), r(H, RH), r(4 * H, RH), r(4, H), r(4 * H)]
x = torch.randn(T, B, I, device=dev, generator=g)
bwd = len(sys.argv) > 7 and sys.argv[7] == "bwd"
tr_buf0 = (ctypes.c_longlong * 8192)()
if bwd:
    for p in canon:
        p.requires_grad_(True)
    y, hT, cT = vmlmf_sequence(x, None, None, canon, False)
    ctypes.CDLL(out).vmlmf_r2_trace_read(tr_buf0, 4096)          # drop the forward's events
    ctypes.CDLL(out).vmlmf_r3_trace_read(tr_buf0, 4096)
    (y.sum() + hT.sum()).backward()
else:
    with torch.no_grad():
        vmlmf_sequence(x, None, None, canon, False)
torch.cuda.synchronize()
tr_buf = (ctypes.c_longlong * 8192)()
h = ctypes.CDLL(out)
getn = h.vmlmf_r3_trace_read if _lib.plan(T, B, I, H, RX, RH).path == _lib.PATH_R3 else h.vmlmf_r2_trace_read
getn.argtypes = [ctypes.c_void_p, ctypes.c_int]
getn.restype = ctypes.c_int
n = getn(tr_buf, 4096)
ev = sorted(((tr_buf[2 * i + 1], tr_buf[2 * i]) for i in range(n)))
t0 = ev[0][0]
prev = t0
for c, e in ev:
    print(f"t={e // 1000:3d} ev={e % 1000:3d}  +{(c - t0) / 1.965e3:9.2f} us  (d {(c - prev) / 1.965e3:7.2f})")
    prev = c
