#!/bin/bash
# One GPU visit: parity tests, the default bench line (both arms), the ncu launch list of the same bench command and one
# `--set full` capture of the kernel named in $1 (regex).  Outputs under gpurun_out/<tag>_*.
TAG=${2:-r1}
KREGEX=${1:-gemm_tn_kernel}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -3 gpurun_out/${TAG}_gpu_tests.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_ref.json
# launch list of the same bench command (short: 2 steps, smaller batch to keep ncu serialisation bounded)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --groups 16 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
# full capture of the dominant kernel
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${KREGEX}" -s 6 -c 3 -o gpurun_out/${TAG}_prof -f \
  python bench.py --steps 1 --warmup 1 --groups 16 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
