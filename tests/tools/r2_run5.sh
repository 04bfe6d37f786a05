#!/bin/bash
mkdir -p gpurun_out
python tests/gpu_microbench.py gram 10 65536 256 2>&1 | tee gpurun_out/r2_micro_gram.log
python tests/gpu_microbench.py gram 10 262144 128 2>&1 | tee -a gpurun_out/r2_micro_gram.log
python tests/gpu_microbench.py gram 10 1048576 64 2>&1 | tee -a gpurun_out/r2_micro_gram.log
timeout 900 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider -x 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --detail > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
head -c 400 gpurun_out/r2_bench5.json
