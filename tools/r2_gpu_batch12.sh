#!/usr/bin/env bash
# round-2 batch: split-K late MaskCLIP pass + fast sigmoid in post_fused: tests, then A/B of the joint CLIP pass on one box
T=${1:-r2p}
O=gpurun_out
mkdir -p $O
(timeout 900 python -m pytest tests/test_gpu_clip.py tests/test_gpu_postprocess.py tests/test_gpu_pipeline.py -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -40) > $O/${T}_pytest_sel.log
tail -3 $O/${T}_pytest_sel.log
timeout 420 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_bench_c2_joint.json 2> $O/${T}_bench_c2_joint.err
ODISE_NO_CLIP_JOINT=1 timeout 420 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_bench_c2_standalone.json 2> $O/${T}_bench_c2_standalone.err
timeout 420 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_bench_c2_joint2.json 2> $O/${T}_bench_c2_joint2.err
tail -c 300 $O/${T}_bench_c2_joint.err
