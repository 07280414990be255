#!/usr/bin/env bash
# round-2 batch: F16Q8 unit + pipeline tests, ncu full captures of the conv-shaped GEMM in both modes, bench A/B
T=${1:-r2g}
O=gpurun_out
mkdir -p $O
(timeout 900 python -m pytest tests/test_gpu_q8.py -m gpu -q --maxfail=30 -p no:cacheprovider 2>&1 | tail -80) > $O/${T}_pytest_q8.log
tail -3 $O/${T}_pytest_q8.log
for m in 3 2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 1 -f -o $O/${T}_ncu_gemm_nmma$m \
      python tools/gemm_one.py 65536 512 4608 f32 $m > $O/${T}_ncu_gemm_nmma$m.log 2>&1
done
{
for shape in "65536 512 4608 f32" "65536 640 320 planes" "65536 2560 320 geglu"; do
  for bn in 128 256; do
    echo "== $shape bn=$bn bf16x3"; timeout 60 python tools/gemm_one.py $shape 3 $bn
    echo "== $shape bn=$bn f16q8"; timeout 60 python tools/gemm_one.py $shape 2 $bn
  done
done
} > $O/${T}_gemm_q8_ab.txt 2>&1
grep -v "^==" $O/${T}_gemm_q8_ab.txt | tail -14
(timeout 1200 python -m pytest tests/test_gpu_pipeline_q8.py -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -60) > $O/${T}_pytest_pipeline_q8.log
tail -3 $O/${T}_pytest_pipeline_q8.log
timeout 420 python bench.py --steps 5 --warmup 3 --precision f16q8 --no-cpu-baseline > $O/${T}_bench_f16q8.json 2> $O/${T}_bench_f16q8.err
timeout 420 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_bench_bf16x3.json 2> $O/${T}_bench_bf16x3.err
tail -c 600 $O/${T}_bench_f16q8.err
