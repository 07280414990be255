"""Times odise_gemm_bf16 on UNet-like shapes (CUDA events, L2-flushing rotation) -> gpurun_out/gemm_probe.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from odise_b200 import lib  # noqa: E402


def time_fn(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = torch.device("cuda")
    res = []
    shapes = [  # (name, M, N, K, conv)
        ("lin_16384x320x320", 16384, 320, 320, None),
        ("ff1_16384x2560x320", 16384, 2560, 320, None),
        ("ff2_16384x320x1280", 16384, 320, 1280, None),
        ("conv64_320", 4 * 64 * 64, 320, 9 * 320, (320, 64, 64)),
        ("conv64_640to320", 4 * 64 * 64, 320, 9 * 640, (640, 64, 64)),
        ("conv32_640", 4 * 32 * 32, 640, 9 * 640, (640, 32, 32)),
        ("conv16_1280", 4 * 16 * 16, 1280, 9 * 1280, (1280, 16, 16)),
        ("conv8_2560to1280", 4 * 8 * 8, 1280, 9 * 2560, (2560, 8, 8)),
        ("big_8192^3", 8192, 8192, 8192, None),
    ]
    for name, M, N, K, conv in shapes:
        a = lib.split(torch.randn(M, K if not conv else conv[0], device=dev))
        b = lib.split(torch.randn(N, K, device=dev) * 0.05)
        out = torch.empty(M, N, device=dev)
        for nmma in (1, 3):
            for bn in (0, 128, 256):
                try:
                    ms = time_fn(lambda: lib.gemm(a, b, M=M, N=N, K=K, nmma=nmma, conv=conv, out=out, force_bn=bn))
                except Exception as ex:  # noqa
                    res.append(dict(name=name, nmma=nmma, bn=bn, error=str(ex)))
                    continue
                tf = 2.0 * M * N * K / ms / 1e9
                res.append(dict(name=name, nmma=nmma, bn=bn, ms=ms, tflops=tf, mma_tflops=tf * nmma))
                print(res[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/gemm_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
