#!/usr/bin/env bash
# round-2 batch: MaskCLIP image tokens share the crop pass of the CLIP tower: tests + bench
T=${1:-r2o}
O=gpurun_out
mkdir -p $O
(timeout 1200 python -m pytest tests/test_gpu_clip.py tests/test_gpu_pipeline.py tests/test_gpu_pipeline_q8.py tests/test_gpu_plugin.py tests/test_gpu_attention.py -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -60) > $O/${T}_pytest_clip.log
tail -3 $O/${T}_pytest_clip.log
timeout 420 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err
timeout 300 python tools/gemm_shapes.py --full > $O/${T}_gemm_shapes_full.txt 2>&1
tail -c 300 $O/${T}_bench_c2.err
