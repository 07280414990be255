"""Runs ONE named hot kernel at its release shape (for `ncu --set full -k regex:... -c 1`), prints a CUDA-event timing.

    python tools/ncu_targets.py conv | attn | msda | mha | clip | gnapply | vaeconv
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from odise_b200 import lib, ops  # noqa: E402

which = sys.argv[1]
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


if which in ("conv", "vaeconv"):
    # ResBlock conv3x3: 16 crops x 64x64 x 320 -> 320 (UNet level 1), or VAE 128x128x512
    B, H, C, Co = (16, 64, 320, 320) if which == "conv" else (16, 128, 512, 512)
    x = lib.split(torch.randn(B * H * H, C, device=dev))
    w = lib.split(torch.randn(Co, 9 * C, device=dev) * 0.02)
    bias = torch.randn(Co, device=dev)
    emb = torch.randn(B, Co, device=dev)
    res = torch.randn(B * H * H, Co, device=dev)
    out = torch.empty(B * H * H, Co, device=dev)
    fn = lambda: lib.gemm(x, w, M=B * H * H, N=Co, conv=(C, H, H), bias=bias, rowbias=emb, rows_per_group=H * H,
                          residual=res, out=out)
    ms = timeit(fn)
    fl = 2.0 * B * H * H * Co * 9 * C
    print(f"{which}: {ms*1000:.1f} us, {fl/ms/1e9:.1f} TFLOP/s algorithmic (bf16x3), {3*fl/ms/1e9:.1f} MMA-TFLOP/s")
elif which == "convgn":
    # the same ResBlock conv with the GroupNorm records of its output written by the epilogue (desc.gn_partial),
    # followed by the record merge + apply + SiLU + split of the consumer (no statistics pass over the activation)
    B, H, C, Co = 16, 64, 320, 320
    x = lib.split(torch.randn(B * H * H, C, device=dev))
    w = lib.split(torch.randn(Co, 9 * C, device=dev) * 0.02)
    bias, emb = torch.randn(Co, device=dev), torch.randn(B, Co, device=dev)
    res, out = torch.randn(B * H * H, Co, device=dev), torch.empty(B * H * H, Co, device=dev)
    st = lib.GnStats(B * H * H, Co, dev)
    ga, be = torch.randn(Co, device=dev), torch.randn(Co, device=dev)
    f1 = lambda: lib.gemm(x, w, M=B * H * H, N=Co, conv=(C, H, H), bias=bias, rowbias=emb, rows_per_group=H * H,
                          residual=res, out=out, gn=st)
    f0 = lambda: lib.gemm(x, w, M=B * H * H, N=Co, conv=(C, H, H), bias=bias, rowbias=emb, rows_per_group=H * H,
                          residual=res, out=out)
    g1 = lambda: ops.group_norm(out, B, H * H, ga, be, 1e-5, act=2, stats=st)
    g0 = lambda: ops.group_norm(out, B, H * H, ga, be, 1e-5, act=2)
    print(f"convgn: conv {timeit(f0)*1000:.1f} us -> with records {timeit(f1)*1000:.1f} us; gn(stats pass + apply) "
          f"{timeit(g0)*1000:.1f} us -> gn(record merge + apply) {timeit(g1)*1000:.1f} us")
elif which == "gnapply":
    B, HW, C = 16, 4096, 320
    x = torch.randn(B * HW, C, device=dev)
    ga, be = torch.randn(C, device=dev), torch.randn(C, device=dev)
    fn = lambda: ops.group_norm(x, B, HW, ga, be, 1e-5, act=2)
    ms = timeit(fn)
    by = B * HW * C * (4 + 4 + 4)      # stats read + apply read + planes write
    print(f"gn(stats+apply+silu+split): {ms*1000:.1f} us, {by/ms/1e6:.1f} GB/s of algorithmic bytes")
elif which == "attn":
    B, heads, d, T = 16, 8, 40, 4096
    HS = 64
    q = lib.split(torch.randn(B * T, heads * HS, device=dev))
    k = lib.split(torch.randn(B * T, heads * HS, device=dev))
    vt = lib.split(torch.randn(heads * HS, B * T, device=dev), f16=True)
    fn = lambda: ops.attention_tc(q, k, vt, B, heads, d, T, T, d ** -0.5, 3)
    ms = timeit(fn, 5)
    fl = 4.0 * B * heads * T * T * d
    print(f"attn_tc d=40 T=4096 B=16: {ms*1000:.1f} us, {fl/ms/1e9:.1f} TFLOP/s algorithmic, {3*fl*48/40/ms/1e9:.1f} MMA-TFLOP/s issued")
elif which == "msda":
    N, M, D, P = 4, 8, 32, 4
    shapes = [(32, 32), (64, 64), (128, 128)]
    S = sum(h * w for h, w in shapes)
    ss = torch.tensor(shapes, dtype=torch.int64, device=dev)
    lsi = torch.tensor([0, 1024, 1024 + 4096], dtype=torch.int64, device=dev)
    value = torch.randn(N * S, M * D, device=dev)
    ref = torch.rand(N, S, 3, 2, device=dev)
    offs = torch.randn(N * S, M * 3 * P * 2, device=dev) * 2
    logits = torch.randn(N * S, M * 3 * P, device=dev)
    fn = lambda: ops.msda_fused(value, ss, lsi, ref, offs, logits, N, S, M, D, 3, S, P)
    ms = timeit(fn)
    by = 4 * N * (S * 256 + S * M * 3 * P * 3 + S * 256)
    print(f"msda_fused N=4 S=21504: {ms*1000:.1f} us, {by/ms/1e6:.1f} GB/s of compulsory bytes ({by/1e6:.1f} MB)")
elif which == "mha":
    B, Q, heads, hw = 4, 100, 8, 128 * 128
    q = torch.randn(B * Q, 256, device=dev)
    kk = torch.randn(B * hw, 768, device=dev)
    vv = torch.randn(B * hw, 768, device=dev)
    ml = torch.randn(B, Q, 256, 256, device=dev) * 3 - 1
    bits, ra = ops.attn_mask_bits(ml, B, Q, 256, 256, 128, 128)
    fn = lambda: ops.mha_d32(q, 256, kk, vv, 768, B, Q, hw, heads, 32 ** -0.5, bits, ra)
    ms = timeit(fn)
    by = 4 * B * (2 * hw * 256 + 2 * Q * 256) + B * Q * hw // 8
    print(f"mha_d32 masked Q=100 keys=16384 B=4: {ms*1000:.1f} us, {by/ms/1e6:.1f} GB/s of compulsory bytes")
    fn2 = lambda: ops.attn_mask_bits(ml, B, Q, 256, 256, 128, 128)
    ms2 = timeit(fn2)
    print(f"attn_mask_bits: {ms2*1000:.1f} us, {B*Q*256*256*4/ms2/1e6:.1f} GB/s")
elif which == "clip":
    for rows, kp in ((400, 403), (800, 1342), (6400, 1342)):
        me = lib.split(torch.nn.functional.normalize(torch.randn(rows, 256, device=dev), dim=-1))
        te = lib.split(torch.nn.functional.normalize(torch.randn(kp, 256, device=dev), dim=-1))
        out = torch.empty(rows, kp, device=dev)
        fn = lambda: lib.gemm(me, te, alpha=14.2857, out=out)
        ms = timeit(fn)
        print(f"clip match [{rows},256]x[{kp},256]^T: {ms*1000:.1f} us, {2.0*rows*kp*256/ms/1e9:.2f} TFLOP/s")
elif which == "clipattn":
    # CLIP ViT-L/14-336 self-attention of 16 crops: 16 heads x d=64, 577 tokens (584 rows per image)
    B, heads, d, T, TS = 16, 16, 64, 577, 584
    qk = lib.split(torch.randn(B * TS, 2 * heads * d, device=dev))
    vt = lib.split(torch.randn(heads * d, B * TS, device=dev), f16=True)
    fn = lambda: ops.attention_tc(qk.col_slice(0, heads * d), qk.col_slice(heads * d, heads * d), vt, B, heads, d, TS, T,
                                  d ** -0.5, 3, tk_stride=TS)
    ms = timeit(fn, 5)
    fl = 4.0 * B * heads * TS * T * d
    print(f"attn_tc d=64 T=577 B=16: {ms*1000:.1f} us, {fl/ms/1e9:.1f} TFLOP/s algorithmic, {3*fl/ms/1e9:.1f} MMA-TFLOP/s issued")
elif which == "clipmlp":
    # CLIP MLP c_fc: [16*584, 1024] x [4096, 1024]^T with bias + QuickGELU + (hi, lo) plane epilogue
    M, K, N = 16 * 584, 1024, 4096
    a = lib.split(torch.randn(M, K, device=dev))
    w = lib.split(torch.randn(N, K, device=dev) * 0.03)
    bias = torch.randn(N, device=dev)
    u = lib.Planes.empty(M, N, dev)
    fn = lambda: lib.gemm(a, w, bias=bias, act=4, out_planes=u)
    ms = timeit(fn)
    fl = 2.0 * M * N * K
    print(f"clip c_fc {M}x{N}x{K}: {ms*1000:.1f} us, {fl/ms/1e9:.1f} TFLOP/s algorithmic, {3*fl/ms/1e9:.1f} MMA-TFLOP/s")
