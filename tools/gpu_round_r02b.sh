#!/bin/bash
# Round-2 evidence, second visit: ncu --set full of the forward-only fused FFN launch and of the rewritten K3 / K4 kernels.
N="ncu --set full --clock-control none --import-source on"
timeout 600 $N -k regex:ffn_fwd_kernel -s 2 -c 1 -o gpurun_out/r02_ffn_fwd_infer -f python tools/bench_ffn.py > gpurun_out/r02_ncu_b.log 2>&1
timeout 600 $N -k regex:cut_loss_pair_kernel -s 8 -c 1 -o gpurun_out/r02_k3_js -f python tools/bench_heads.py >> gpurun_out/r02_ncu_b.log 2>&1
timeout 600 $N -k regex:eval_cut_fast_kernel -s 2 -c 1 -o gpurun_out/r02_k4 -f python tools/bench_heads.py >> gpurun_out/r02_ncu_b.log 2>&1
ls -la gpurun_out | tail -5
