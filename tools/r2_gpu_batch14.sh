#!/usr/bin/env bash
T=${1:-r2s}
O=gpurun_out
mkdir -p $O
(timeout 600 python -m pytest tests/test_gpu_q8.py tests/test_gpu_vae.py -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -20) > $O/${T}_pytest_sel.log
tail -3 $O/${T}_pytest_sel.log
timeout 120 python - > $O/${T}_softmax_timing.txt 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, ".")
from odise_b200 import lib, ops
x = torch.randn(65536, 4096, device="cuda")
for lo in (True, lib.Q8):
    for _ in range(3): p = ops.softmax_split(x, 65536, 4096, 4096, 0.044, lo=lo)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(5): p = ops.softmax_split(x, 65536, 4096, 4096, 0.044, lo=lo)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(f"softmax_split 65536 x 4096 lo={lo}: {ms*1000:.1f} us, {2*65536*4096*4/ms/1e6:.0f} GB/s of read + write")
PY
cat $O/${T}_softmax_timing.txt
timeout 420 python bench.py --steps 10 --warmup 3 > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err
tail -c 200 $O/${T}_bench_c2.err
