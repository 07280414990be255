#!/usr/bin/env bash
# round-2 final evidence: full GPU suite, smoke, bench lines c2 (default + bf16x3 on the same box) / c3 / c4 / c5, per-shape GEMM
# list, ncu launch list of one eager step
T=${1:-r2z}
O=gpurun_out
mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider 2>&1 | tail -60) > $O/${T}_pytest.log
tail -3 $O/${T}_pytest.log
timeout 300 python __graft_entry__.py --smoke > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
timeout 420 python bench.py --steps 10 --warmup 3 > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err
timeout 420 python bench.py --steps 5 --warmup 3 --precision bf16x3 --no-cpu-baseline > $O/${T}_bench_c2_bf16x3.json 2> $O/${T}_bench_c2_bf16x3.err
timeout 240 python bench.py --config c4 > $O/${T}_bench_c4.json 2> $O/${T}_bench_c4.err
timeout 420 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_bench_c3_n1.json 2> $O/${T}_bench_c3_n1.err
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_bench_c5.json 2> $O/${T}_bench_c5.err
timeout 300 python tools/gemm_shapes.py --full > $O/${T}_gemm_shapes_full.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/${T}_launches.csv python tools/profile_step.py --full --iters 2 > $O/${T}_profile_step.log 2>&1
python tools/summarize_launches.py $O/${T}_launches.csv > $O/${T}_launch_summary.txt 2>&1
tail -c 300 $O/${T}_bench_c2.err
