B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_choopy.csv $B > gpurun_out/r02_ncu_bench.log 2>&1
N="ncu --set full --clock-control none --import-source on"
timeout 600 $N -k regex:ffn_fwd_kernel -s 4 -c 1 -o gpurun_out/r02_ffn_fwd -f $B > gpurun_out/r02_ncu_full.log 2>&1
timeout 600 $N -k regex:attn_lists_fwd_tc -s 4 -c 1 -o gpurun_out/r02_attn_tc -f $B >> gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out | tail -6
