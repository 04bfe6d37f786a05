set -x
P() { # name, args...
  name=$1; shift
  python tests/gpu_microbench.py "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s ${SKIP:-3} -c 1 -o gpurun_out/r1_$name python tests/gpu_microbench.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name exit $?"; cat gpurun_out/plain_$name.log
}
KREGEX=gemm_f16_tc P conv_3x3_N256_K2304 conv 2 256 16 16 256 256 3 1 1 3
KREGEX=gemm_f16_tc P gemm_N256_K64 gemm 2 1048576 256 64 3
KREGEX=gemm_f16_tc SKIP=6 P gemm_bn_fused_N256_K64 gemm_bn 2 1048576 256 64 3
KREGEX=bn_act_kernel P bn_act bnact 2 1048576 256 3
KREGEX=mc_reduce P mc_reduce mcreduce 30 1048576 7 3
KREGEX=kl_kernel P kl kl
KREGEX=sample_weights P sample sample
python bench.py --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 22000 -c 2500 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "launchlist exit $?"
tail -2 gpurun_out/plain_bench.log | cut -c1-400
ls -la gpurun_out | tail -25
