#!/usr/bin/env bash
# round-2 batch for the F16Q8 operand mode: unit tests first (fail fast), single-shape A/B timings bf16x3 vs F16Q8, UNet taps
T=${1:-r2f}
O=gpurun_out
mkdir -p $O
(timeout 900 python -m pytest tests/test_gpu_q8.py -m gpu -q --maxfail=30 -p no:cacheprovider 2>&1 | tail -80) > $O/${T}_pytest_q8.log
tail -3 $O/${T}_pytest_q8.log
{
for shape in "65536 512 4608 f32" "65536 320 2880 f32" "65536 320 320 f32" "65536 640 320 planes" "16384 1280 640 planes" "9344 4096 1024 planes" "65536 2560 320 geglu"; do
  echo "== $shape  bf16x3"; timeout 60 python tools/gemm_one.py $shape 3
  echo "== $shape  f16q8"; timeout 60 python tools/gemm_one.py $shape 2
done
} > $O/${T}_gemm_q8_ab.txt 2>&1
cat $O/${T}_gemm_q8_ab.txt | grep -v "^==" | tail -20
(timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_gemm.py tests/test_gpu_attention.py -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -30) > $O/${T}_pytest_unet.log
tail -3 $O/${T}_pytest_unet.log
