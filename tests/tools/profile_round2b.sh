#!/bin/bash
mkdir -p gpurun_out
P() { name=$1; shift
  python tests/gpu_microbench.py "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s ${SKIP:-3} -c 1 -f -o gpurun_out/r2_$name python tests/gpu_microbench.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name exit $?"; cat gpurun_out/plain_$name.log; }
KREGEX=gemm_f16_tc SKIP=8 P gemm_bn_fused_N512_K128 gemm_bn 10 262144 512 128 3
KREGEX=gemm_f16_tc SKIP=8 P gemm_bn_fused_N1024_K256 gemm_bn 10 65536 1024 256 3
KREGEX=conv3x3_c64_stream SKIP=2 P conv3x3_c64_stream conv 10 256 64 64 64 64 3 1 1 3
KREGEX=mc_reduce P mc_reduce mcreduce 30 1048576 7 3
KREGEX=gram_quadform SKIP=5 P gram_quadform_K256 gram 10 65536 256
KREGEX=kl_kernel P kl kl
