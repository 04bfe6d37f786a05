#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider -x 2>&1 | tail -40 > gpurun_out/r2_pytest3.log
tail -30 gpurun_out/r2_pytest3.log
timeout 900 python bench.py --steps 5 --warmup 3 --detail --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
head -c 600 gpurun_out/r2_bench3.json; echo; head -30 gpurun_out/r2_bench3.err
