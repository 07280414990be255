#!/usr/bin/env bash
# One gpurun call of round 2: tests, bench lines, CUDA-event kernel timings, ncu launch list + full captures.
# usage (on the box): bash tools/r2_gpu_batch.sh <tag>
T=${1:-r2x}
O=gpurun_out
mkdir -p $O
(timeout 1000 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider 2>&1 | tail -120) > $O/${T}_pytest.log
timeout 240 python bench.py --config c4 > $O/${T}_c4.json 2> $O/${T}_c4.err
timeout 420 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err
for t in conv convgn attn clipattn msda gnapply; do timeout 120 python tools/ncu_targets.py $t; done > $O/${T}_targets_timing.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/${T}_launches.csv python tools/profile_step.py --full --iters 2 > $O/${T}_profile_step.log 2>&1
python tools/summarize_launches.py $O/${T}_launches.csv > $O/${T}_launch_summary.txt 2>&1
for t in msda attn; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"msda_d32|attn_tc" -c 1 -f -o $O/${T}_ncu_$t \
      python tools/ncu_targets.py $t > $O/${T}_ncu_$t.log 2>&1
done
tail -3 $O/${T}_pytest.log
