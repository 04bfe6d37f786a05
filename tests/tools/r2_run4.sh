#!/bin/bash
mkdir -p gpurun_out
python tests/gpu_microbench.py sample 2>&1 | tee gpurun_out/r2_micro_misc.log
python tests/gpu_microbench.py mcreduce 30 1048576 7 5 2>&1 | tee -a gpurun_out/r2_micro_misc.log
python tests/gpu_microbench.py gram 10 65536 256 2>&1 | tee -a gpurun_out/r2_micro_misc.log
python tests/gpu_microbench.py gram 10 262144 128 2>&1 | tee -a gpurun_out/r2_micro_misc.log
timeout 900 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider -x 2>&1 | tail -5
