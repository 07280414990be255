#!/usr/bin/env bash
# round-2 batch: attention with O accumulated in TMEM (lazy rescale): tests, timings, ncu; ncu of post_fused_kernel
T=${1:-r2m}
O=gpurun_out
mkdir -p $O
(timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_unet.py tests/test_gpu_clip.py tests/test_gpu_head.py -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -40) > $O/${T}_pytest_attn.log
tail -3 $O/${T}_pytest_attn.log
for t in attn clipattn; do timeout 120 python tools/ncu_targets.py $t; done > $O/${T}_targets_timing.txt 2>&1
cat $O/${T}_targets_timing.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc -c 1 -f -o $O/${T}_ncu_attn \
    python tools/ncu_targets.py attn > $O/${T}_ncu_attn.log 2>&1
timeout 420 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:post_fused --profile-from-start off -c 1 -f -o $O/${T}_ncu_post \
    python tools/profile_step.py --full --iters 1 > $O/${T}_ncu_post.log 2>&1
tail -c 300 $O/${T}_bench_c2.err
