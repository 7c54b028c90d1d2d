#!/bin/bash
# One GPU visit (1 GPU): parity tests, the default bench line (both arms), the ncu launch list of the same bench command
# and `--set full` captures of the dominant kernels.  Outputs under gpurun_out/<tag>_*.   usage: gpu_round.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -3 gpurun_out/${TAG}_gpu_tests.log
# rows N3 / N4: roofline lines of the gather and rank-metric kernels (never measured in round 1)
timeout 300 python tools/bench_data.py > gpurun_out/${TAG}_bench_data.txt 2>&1; tail -6 gpurun_out/${TAG}_bench_data.txt
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2>> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench_ref.json gpurun_out/${TAG}_bench.json | cut -c1-600
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
# launch list of the same bench command (1 warm-up + 1 timed step, full batch)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_bench.log 2>&1
# full captures (each .ncu-rep is ~20 MB with sources and gpurun copies back at most 64 MiB per call: two here, the
# others with tools/gpu_round_more.sh): fused FFN backward (2nd launch of a step), attention backward
N="ncu --set full --clock-control none --import-source on"
timeout 600 $N -k regex:ffn_bwd_kernel -s 1 -c 1 -o gpurun_out/${TAG}_ffn_bwd -f $B > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 $N -k regex:attn_lists_bwd -s 1 -c 1 -o gpurun_out/${TAG}_attn_bwd -f $B >> gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -14
