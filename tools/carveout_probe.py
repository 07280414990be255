"""Does alternating small-smem elementwise kernels with max-smem GEMMs cost an SM re-partition per switch?
Times a dependent chain [layer_norm -> tiny GEMM] x 200 with and without the prefer-shared device policy."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from odise_b200 import lib, ops  # noqa: E402

dev = torch.device("cuda")
x = torch.randn(400, 256, device=dev)
g_, b_ = torch.ones(256, device=dev), torch.zeros(256, device=dev)
w = lib.split(torch.randn(256, 256, device=dev) * 0.05)
out = torch.empty(400, 256, device=dev)


def chain(n=200):
    h = x
    for _ in range(n):
        _, p = ops.layer_norm(h, g_, b_)
        lib.gemm(p, w, out=out)
        h = out


def gemm_only(n=200):
    p = lib.split(x)
    for _ in range(n):
        lib.gemm(p, w, out=out)


def timed(fn):
    gr = torch.cuda.CUDAGraph()
    fn(3)
    torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        fn()
    gr.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(5):
        gr.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / 5 / 200 * 1000


for pol in (0, 1, 0, 1):
    lib._check(lib.load().odise_set_carveout_policy(pol), "policy")
    print(f"policy prefer_shared={pol}: LN+GEMM pair {timed(chain):.2f} us, GEMM alone {timed(gemm_only):.2f} us", flush=True)
