#!/usr/bin/env bash
# round-2 batch: long-row softmax, one-pass gn_finalize, VAE conv_in: full suite + bench + launch summary
T=${1:-r2r}
O=gpurun_out
mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider 2>&1 | tail -60) > $O/${T}_pytest.log
tail -3 $O/${T}_pytest.log
timeout 420 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/${T}_launches.csv python tools/profile_step.py --full --iters 2 > $O/${T}_profile_step.log 2>&1
python tools/summarize_launches.py $O/${T}_launches.csv > $O/${T}_launch_summary.txt 2>&1
tail -c 300 $O/${T}_bench_c2.err
