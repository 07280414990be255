#!/usr/bin/env bash
# round-2 batch for the TMA-store epilogue: GEMM tests first (fail fast), A/B single-shape timings, full suite, bench A/B
T=${1:-r2e}
O=gpurun_out
mkdir -p $O
(timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -40) > $O/${T}_pytest_gemm.log
tail -2 $O/${T}_pytest_gemm.log
{
for shape in "65536 320 320 f32" "65536 640 320 planes" "65536 320 1280 f32" "16384 1280 640 planes" "9344 4096 1024 planes" "86016 256 1024 f32"; do
  echo "== $shape  TMA store"; timeout 60 python tools/gemm_one.py $shape
  echo "== $shape  plain epilogue"; ODISE_NO_TMA_STORE=1 timeout 60 python tools/gemm_one.py $shape
done
} > $O/${T}_gemm_ab.txt 2>&1
(timeout 1000 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider 2>&1 | tail -60) > $O/${T}_pytest.log
timeout 420 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err
ODISE_NO_TMA_STORE=1 timeout 420 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_bench_no_tma.json 2> $O/${T}_bench_no_tma.err
tail -3 $O/${T}_pytest.log
