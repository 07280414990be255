#!/usr/bin/env bash
# round-2 batch: measured tile / pair policy: full suite, smoke, bench (default f16q8, bf16x3), per-shape GEMM list, launch list
T=${1:-r2k}
O=gpurun_out
mkdir -p $O
(timeout 1000 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider 2>&1 | tail -60) > $O/${T}_pytest.log
tail -3 $O/${T}_pytest.log
timeout 300 python __graft_entry__.py --smoke > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
timeout 420 python bench.py --steps 5 --warmup 3 > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err
timeout 420 python bench.py --steps 5 --warmup 3 --precision bf16x3 --no-cpu-baseline > $O/${T}_bench_c2_bf16x3.json 2> $O/${T}_bench_c2_bf16x3.err
timeout 300 python tools/gemm_shapes.py --full > $O/${T}_gemm_shapes_full.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/${T}_launches.csv python tools/profile_step.py --full --iters 2 > $O/${T}_profile_step.log 2>&1
python tools/summarize_launches.py $O/${T}_launches.csv > $O/${T}_launch_summary.txt 2>&1
tail -c 300 $O/${T}_bench_c2.err
