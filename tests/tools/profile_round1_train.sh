set -x
P() { # name, args...
  name=$1; shift
  python tests/gpu_microbench.py "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s ${SKIP:-3} -c 1 -o gpurun_out/r1_$name python tests/gpu_microbench.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name exit $?"; cat gpurun_out/plain_$name.log
}
KREGEX=gemm_f16_tc P wgrad_3x3_256 wgrad 30 8 16 16 256 256 3 1 1 1 3
KREGEX=gemm_f16_tc P wgrad_1x1_64_256 wgrad 30 8 64 64 64 256 1 1 0 4 3
KREGEX=bn_bwd_apply SKIP=2 P bn_bwd_apply bnbwd 30 32768 256 3
KREGEX=bn_bwd_reduce SKIP=2 P bn_bwd_reduce bnbwd 30 32768 256 3
ls -la gpurun_out | tail -12
KREGEX=conv3x3_c64_stream SKIP=2 P conv3x3_c64_stream conv 10 256 64 64 64 64 3 1 1 3
