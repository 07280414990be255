#!/usr/bin/env bash
# round-2 batch: 2-SM MMAs (cta_group::2): GEMM tests forced on every shape, A/B timings vs single CTAs / multicast pairs, ncu, suite, bench
T=${1:-r2i}
O=gpurun_out
mkdir -p $O
(ODISE_GEMM_CLUSTER=2 timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_q8.py -m gpu -q --maxfail=8 -p no:cacheprovider 2>&1 | tail -60) > $O/${T}_pytest_gemm_sm2.log
tail -3 $O/${T}_pytest_gemm_sm2.log
{
for shape in "65536 512 4608 f32" "65536 320 2880 f32" "65536 640 320 planes" "9344 4096 1024 planes" "65536 2560 320 geglu" "86016 256 1024 f32"; do
  for m in 3 2; do
    echo "== $shape nmma=$m single"; ODISE_GEMM_CLUSTER=0 timeout 60 python tools/gemm_one.py $shape $m
    echo "== $shape nmma=$m multicast pairs"; ODISE_GEMM_CLUSTER=3 timeout 60 python tools/gemm_one.py $shape $m
    echo "== $shape nmma=$m 2-SM"; ODISE_VERBOSE=1 timeout 60 python tools/gemm_one.py $shape $m
  done
done
} > $O/${T}_gemm_sm2_ab.txt 2>&1
grep -v "^==" $O/${T}_gemm_sm2_ab.txt | grep -v co-resident | tail -40
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 1 -f -o $O/${T}_ncu_gemm_nmma2_sm2 \
    python tools/gemm_one.py 65536 512 4608 f32 2 > $O/${T}_ncu_gemm_nmma2_sm2.log 2>&1
(timeout 1000 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider 2>&1 | tail -60) > $O/${T}_pytest.log
tail -3 $O/${T}_pytest.log
timeout 420 python bench.py --steps 5 --warmup 3 --precision f16q8 --no-cpu-baseline > $O/${T}_bench_f16q8_sm2.json 2> $O/${T}_bench_f16q8_sm2.err
ODISE_GEMM_CLUSTER=0 timeout 420 python bench.py --steps 5 --warmup 3 --precision f16q8 --no-cpu-baseline > $O/${T}_bench_f16q8_single.json 2> $O/${T}_bench_f16q8_single.err
timeout 420 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_bench_bf16x3_sm2.json 2> $O/${T}_bench_bf16x3_sm2.err
tail -c 400 $O/${T}_bench_f16q8_sm2.err
