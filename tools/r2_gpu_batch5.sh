#!/usr/bin/env bash
# round-2 batch: CTA-pair multicast of the B tile (GemmParams::cl): GEMM tests in both settings, A/B timings, ncu, bench A/B
T=${1:-r2h}
O=gpurun_out
mkdir -p $O
(ODISE_GEMM_CLUSTER=2 timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_q8.py -m gpu -q --maxfail=20 -p no:cacheprovider 2>&1 | tail -60) > $O/${T}_pytest_gemm_cl2.log
tail -3 $O/${T}_pytest_gemm_cl2.log
{
for shape in "65536 512 4608 f32" "65536 320 2880 f32" "65536 640 320 planes" "9344 4096 1024 planes" "65536 2560 320 geglu" "86016 256 1024 f32"; do
  for m in 3 2; do
    echo "== $shape nmma=$m no cluster"; ODISE_GEMM_CLUSTER=0 timeout 60 python tools/gemm_one.py $shape $m
    echo "== $shape nmma=$m pairs"; ODISE_VERBOSE=1 timeout 60 python tools/gemm_one.py $shape $m
  done
done
} > $O/${T}_gemm_cluster_ab.txt 2>&1
grep -v "^==" $O/${T}_gemm_cluster_ab.txt | tail -30
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 1 -f -o $O/${T}_ncu_gemm_nmma2_pairs \
    python tools/gemm_one.py 65536 512 4608 f32 2 > $O/${T}_ncu_gemm_nmma2_pairs.log 2>&1
(timeout 1000 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider 2>&1 | tail -60) > $O/${T}_pytest.log
tail -3 $O/${T}_pytest.log
timeout 420 python bench.py --steps 5 --warmup 3 --precision f16q8 --no-cpu-baseline > $O/${T}_bench_f16q8_pairs.json 2> $O/${T}_bench_f16q8_pairs.err
ODISE_GEMM_CLUSTER=0 timeout 420 python bench.py --steps 5 --warmup 3 --precision f16q8 --no-cpu-baseline > $O/${T}_bench_f16q8_nocl.json 2> $O/${T}_bench_f16q8_nocl.err
timeout 420 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_bench_bf16x3_pairs.json 2> $O/${T}_bench_bf16x3_pairs.err
tail -c 400 $O/${T}_bench_f16q8_pairs.err
