#!/usr/bin/env bash
# 2-SM policy data: single CTAs vs 2-SM MMAs per (shape, BN), F16Q8 and bf16x3
T=${1:-r2j}
O=gpurun_out
mkdir -p $O
{
for shape in "1048576 128 1152 f32 128" "262144 256 2304 f32 256" "262144 256 2304 f32 128" "65536 512 4608 f32 128" "16384 640 5760 f32 128" "16384 640 5760 f32 256" "4096 1280 11520 f32 128" "4096 1280 11520 f32 256" "4096 1280 11520 f32 160" "9344 1024 4096 f32 256" "9344 1024 4096 f32 128" "9344 4096 1024 planes 128" "65536 320 2880 f32 128" "65536 320 1280 f32 160" "65536 320 1280 f32 128"; do
  set -- $shape
  for m in 2 3; do
    echo "== $1 $2 $3 $4 bn=$5 nmma=$m single"; ODISE_GEMM_CLUSTER=0 timeout 60 python tools/gemm_one.py $1 $2 $3 $4 $m $5
    echo "== $1 $2 $3 $4 bn=$5 nmma=$m 2-SM"; ODISE_GEMM_CLUSTER=2 timeout 60 python tools/gemm_one.py $1 $2 $3 $4 $m $5
  done
done
} > $O/${T}_gemm_sm2_policy.txt 2>&1
grep -v "^==" $O/${T}_gemm_sm2_policy.txt | tail -70
