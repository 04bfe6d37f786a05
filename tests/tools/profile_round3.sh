#!/bin/bash
# Round 3: ncu --set full of the kernels added / changed this round (each command first run plain, must exit 0; numbers printed
# under ncu are never bench values). Run through gpurun; summaries: tests/tools/ncu_summary.py -> profiles/r3_ncu_*.md
mkdir -p gpurun_out
P() { name=$1; shift
  python tests/gpu_microbench.py "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s ${SKIP:-3} -c 1 -f -o gpurun_out/r3_$name python tests/gpu_microbench.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name exit $?"; cat gpurun_out/plain_$name.log; }
KREGEX=stem_conv_pool SKIP=2 P stem_pool stem 15 256 3
KREGEX=gemm_f16_tc SKIP=3 P conv_3x3_N128_K1152_mstack conv 10 256 32 32 128 128 3 1 1
# tail microbench, iters = 3: gemm_f16_tc launches 0-5 second moments (materialised a2), 6-10 fused conv3, 11-15 second moments with
# the operand transform, 16-20 fused conv3 with the operand transform
KREGEX=gemm_f16_tc SKIP=8 P tail_fused_K64 tail 15 1048576 256 64 3
KREGEX=gemm_f16_tc SKIP=13 P tail_gram_xf_K64 tail 15 1048576 256 64 3
KREGEX=gemm_f16_tc SKIP=18 P tail_fused_xf_K64 tail 15 1048576 256 64 3
# launch list of the default bench command (one whole step = the launches between two mc_reduce_kernel launches; a step is ~860
# launches at the default group of 30 samples)
python bench.py --no-cpu-baseline --no-x3 --no-train-leg --steps 1 --warmup 3 > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 3500 -c 2000 --csv --log-file gpurun_out/r3_launches_bench.csv python bench.py --no-cpu-baseline --no-x3 --no-train-leg --steps 1 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "launchlist exit $?"
