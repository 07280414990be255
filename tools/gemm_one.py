"""Single-shape GEMM runner for ncu captures: python tools/gemm_one.py M N K [planes|f32|both|geglu] [nmma] [bn]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from odise_b200 import lib  # noqa: E402

M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else "f32"
nmma = int(sys.argv[5]) if len(sys.argv) > 5 else 3
bn = int(sys.argv[6]) if len(sys.argv) > 6 else 0
dev = torch.device("cuda")
lo = lib.Q8 if nmma == 2 else True          # nmma 2: F16Q8 operands (fp16 hi*hi + e5m2 cross terms)
a = lib.split(torch.randn(M, K, device=dev), lo=lo)
b = lib.split(torch.randn(N, K, device=dev) * 0.05, lo=lo)
bias = torch.randn(N, device=dev)
out = torch.empty(M, N, device=dev) if mode in ("f32", "both") else None
outp = lib.Planes.empty(M, N, dev, lo=lo) if mode in ("planes", "both") else None
kw = {}
if mode == "geglu":
    outp, kw = lib.Planes.empty(M, N // 2, dev, lo=lo), dict(geglu=True)
for i in range(5):
    lib.gemm(a, b, nmma=nmma, bias=bias, out=out, out_planes=outp, force_bn=bn, **kw)
torch.cuda.synchronize()
s, e = torch.cuda.Event(True), torch.cuda.Event(True)
s.record()
for i in range(10):
    lib.gemm(a, b, nmma=nmma, bias=bias, out=out, out_planes=outp, force_bn=bn, **kw)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
print(f"M={M} N={N} K={K} {mode} nmma={nmma} bn={bn}: {ms*1000:.1f} us, {2*M*N*K/ms/1e9:.1f} TFLOP/s")
