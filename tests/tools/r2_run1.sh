#!/bin/bash
# round-2 GPU validation pass: full GPU test suite, HBM write probes, bench with per-shape detail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r2_pytest.log
timeout 300 python tests/gpu_microbench.py hbmwrite > gpurun_out/r2_hbmwrite.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --detail > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -5 gpurun_out/r2_pytest.log
cat gpurun_out/r2_hbmwrite.log
head -c 1500 gpurun_out/r2_bench.json
