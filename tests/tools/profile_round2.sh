#!/bin/bash
# Round-2 ncu evidence: every command is first run plain (must exit 0), then under ncu. One gpurun call.
mkdir -p gpurun_out
P() { name=$1; shift
  python tests/gpu_microbench.py "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s ${SKIP:-3} -c 1 -f -o gpurun_out/r2_$name python tests/gpu_microbench.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name exit $?"; cat gpurun_out/plain_$name.log; }
KREGEX=gemm_f16_tc SKIP=8 P gemm_bn_fused_N256_K64 gemm_bn 10 1048576 256 64 3
KREGEX=gemm_f16_tc P gemm_N256_K64 gemm 10 1048576 256 64 3
KREGEX=gemm_f16_tc P conv_3x3_N256_K2304 conv 10 256 16 16 256 256 3 1 1 3
KREGEX=gemm_f16_tc P conv_3x3_N128_K1152 conv 10 256 32 32 128 128 3 1 1 3
KREGEX=conv3x3_c64_stream SKIP=5 P conv3x3_c64_stream conv 10 256 64 64 64 64 3 1 1 3
KREGEX=gemm_f16_tc P gram_K64 gram 10 1048576 64
KREGEX=gram_quadform SKIP=5 P gram_quadform_K256 gram 10 65536 256
KREGEX=mc_reduce P mc_reduce mcreduce 30 1048576 7 3
KREGEX=kl_kernel P kl kl
KREGEX=sample_weights P sample sample
KREGEX=adam_update P adam adam
KREGEX=bn_act_kernel P bn_act bnact 10 1048576 64 3
python bench.py --no-cpu-baseline --no-x3 --no-train-leg --steps 1 --warmup 3 > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 9000 -c 4200 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --no-cpu-baseline --no-x3 --no-train-leg --steps 1 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "launchlist exit $?"
tail -1 gpurun_out/plain_bench.log | cut -c1-300
ls -la gpurun_out | grep r2_ | tail -20
