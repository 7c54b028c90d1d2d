#!/bin/bash
# Second GPU visit of a round: the remaining `ncu --set full` captures (FFN1, attention forward) and the other model
# families' bench lines.   usage: gpu_round_more.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
N="ncu --set full --clock-control none --import-source on"
timeout 600 $N -k regex:gemm_f16out_kernel -s 2 -c 1 -o gpurun_out/${TAG}_ffn1 -f $B > gpurun_out/${TAG}_ncu_full2.log 2>&1
timeout 600 $N -k regex:attn_lists_fwd -s 1 -c 1 -o gpurun_out/${TAG}_attn_fwd -f $B >> gpurun_out/${TAG}_ncu_full2.log 2>&1
for m in bicut attncut mtchoopy mtattncut; do
  timeout 300 python bench.py --model $m --no-cpu-baseline > gpurun_out/${TAG}_bench_${m}.json 2>> gpurun_out/${TAG}_bench2.err; echo "$m rc=$?"
done
timeout 300 python bench.py --model mmoecut --groups 32 --no-cpu-baseline > gpurun_out/${TAG}_bench_mmoecut.json 2>> gpurun_out/${TAG}_bench2.err; echo "mmoecut rc=$?"
cut -c1-220 gpurun_out/${TAG}_bench_*.json
ls -la gpurun_out | tail -12
