#!/usr/bin/env bash
# round-2 batch: per-shape GEMM autotune: GEMM / pipeline tests, bench c2 (+ per-shape list), c3 / c4 / c5 lines
T=${1:-r2l}
O=gpurun_out
mkdir -p $O
(timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_q8.py tests/test_gpu_pipeline_q8.py tests/test_gpu_unet.py -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -40) > $O/${T}_pytest_sel.log
tail -3 $O/${T}_pytest_sel.log
timeout 420 python bench.py --steps 5 --warmup 3 > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err
ODISE_VERBOSE=1 timeout 300 python tools/gemm_shapes.py --full > $O/${T}_gemm_shapes_full.txt 2> $O/${T}_gemm_autotune_choices.txt
timeout 240 python bench.py --config c4 > $O/${T}_bench_c4.json 2> $O/${T}_bench_c4.err
timeout 420 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_bench_c3_n1.json 2> $O/${T}_bench_c3_n1.err
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 > $O/${T}_bench_c5.json 2> $O/${T}_bench_c5.err
tail -c 300 $O/${T}_bench_c2.err
