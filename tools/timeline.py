"""Kernel timeline of CUDA-graph replays of the full step (torch.profiler / CUPTI): busy time, idle gaps between kernels
and warm per-kernel totals (ncu launch lists are cold-cache and serialised) -> gpurun_out/timeline_summary.txt"""
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from odise_b200 import spec  # noqa: E402
from odise_b200.pipeline import ODISEEngine, full_param_list, synthetic_vocabulary  # noqa: E402

full = "--hot-path-only" not in sys.argv
dev = torch.device("cuda:0")
eng = ODISEEngine(spec.synth_state_dict(full_param_list(with_vae=full, with_clip=full), 0), dev, synthetic_uncond=True, nmma=3, with_vae=full,
                  with_clip=full)
eng.set_synthetic_vocabulary("ade150", 150, 403)
g, out = eng.capture(4, 1024, 1024)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
REPS = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(REPS):
        g.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy, gaps, last_end = 0.0, [], None
tot = defaultdict(lambda: [0, 0.0])
for e in ev:
    s, en = e.time_range.start, e.time_range.end
    busy += en - s
    if last_end is not None and s > last_end:
        gaps.append(s - last_end)
    last_end = en if last_end is None else max(last_end, en)
    k = e.name.split("(")[0]
    tot[k][0] += 1
    tot[k][1] += en - s
span = t1 - t0
lines = [f"{REPS} graph replays: span {span/1000/REPS:.2f} ms/step, kernel busy {busy/1000/REPS:.2f} ms/step, "
         f"idle between kernels {sum(gaps)/1000/REPS:.2f} ms/step over {len(gaps)//REPS} gaps "
         f"(median gap {sorted(gaps)[len(gaps)//2] if gaps else 0:.2f} us), {len(ev)//REPS} kernels/step"]
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    lines.append(f"{us/1000/REPS:9.3f} ms {100*us/busy:5.1f}%  n={n//REPS:5d}  avg={us/n:9.1f} us  {k[:100]}")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/timeline_summary.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
