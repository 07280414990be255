"""Runs the eager (non-graph) hot-path step a few times — the command ncu wraps for the launch list / captures."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from odise_b200 import lib, spec  # noqa: E402
from odise_b200.pipeline import ODISEEngine, full_param_list, synthetic_vocabulary  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--precision", default="f16q8", choices=["f16q8", "bf16x3", "bf16"])
ap.add_argument("--full", action="store_true", help="include the VAE (f-1) and CLIP image tower (f-2) stages")
a = ap.parse_args()
dev = torch.device("cuda:0")
sd = spec.synth_state_dict(full_param_list(with_vae=a.full, with_clip=a.full), 0)
eng = ODISEEngine(sd, dev, nmma={"f16q8": 2, "bf16x3": 3, "bf16": 1}[a.precision], synthetic_uncond=True, with_vae=a.full, with_clip=a.full)
eng.set_synthetic_vocabulary("ade150", 150, 403)
n0 = lib.launch_count()
for i in range(a.iters):
    last = i == a.iters - 1
    if last:                      # ncu --profile-from-start off: only the LAST eager step is profiled (weights are prepared)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    with lib.nvtx("odise_step"):
        eng.step_full(a.batch, a.size, a.size)
    torch.cuda.synchronize()
    if last:
        torch.cuda.profiler.stop()
    print("iter", i, "launches so far", lib.launch_count() - n0, flush=True)
