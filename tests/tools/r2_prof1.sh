#!/bin/bash
# ncu --set full of three kernels under investigation (each command first run plain)
mkdir -p gpurun_out
P() { name=$1; shift
  python tests/gpu_microbench.py "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s ${SKIP:-3} -c 1 -f -o gpurun_out/r2_$name python tests/gpu_microbench.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name exit $?"; cat gpurun_out/plain_$name.log; }
KREGEX=gemm_f16_tc P gram_K64 gram 10 1048576 64
KREGEX=gemm_f16_tc SKIP=8 P fused_N512_K128 gemm_bn 10 262144 512 128 3
KREGEX=gemm_f16_tc P gemm_N256_K64 gemm 10 1048576 256 64 3
