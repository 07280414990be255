for bn in 0 128 256; do python tools/gemm_one.py 65536 2560 320 geglu 3 $bn; done
python tools/gemm_one.py 16384 5120 640 geglu 3 0
python tools/gemm_one.py 65536 320 320 f32 3 0
python tools/gemm_one.py 86016 1024 256 f32 3 0
python tools/gemm_one.py 9344 4096 1024 planes 3 0
python tools/gemm_one.py 9344 1024 1024 f32 3 0
python tools/gemm_one.py 4096 1280 1280 f32 3 0
