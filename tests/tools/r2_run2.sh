#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider -s 2>&1 > gpurun_out/r2_pytest_full.log
grep -E "max abs err|argmax|decided|passed|failed|FAIL|Error" gpurun_out/r2_pytest_full.log | tail -40
for shape in "10 1048576 256 64" "10 262144 512 128" "10 65536 1024 256"; do timeout 120 python tests/gpu_microbench.py gemm_bn $shape 5; done 2>&1 | tee gpurun_out/r2_micro_gemm_bn.log
timeout 120 python tests/gpu_microbench.py gemm 10 1048576 256 64 5 2>&1 | tee -a gpurun_out/r2_micro_gemm_bn.log
timeout 120 python tests/gpu_microbench.py gemm 10 1048576 64 256 5 2>&1 | tee -a gpurun_out/r2_micro_gemm_bn.log
timeout 900 python bench.py --steps 5 --warmup 3 --detail --no-cpu-baseline --no-train-leg > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
head -c 900 gpurun_out/r2_bench2.json; head -12 gpurun_out/r2_bench2.err
