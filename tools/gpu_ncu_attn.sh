#!/bin/bash
# ncu --set full of one fused-attention launch of the sampler (after 3 warm-up forwards = 72 launches)
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:attn_fused --launch-skip 80 --launch-count 1 -o gpurun_out/r02_ncu_attn -f python tools/profile_sampler_step.py fp32 > gpurun_out/r02_ncu_attn.log 2>&1; tail -2 gpurun_out/r02_ncu_attn.log
ncu -i gpurun_out/r02_ncu_attn.ncu-rep --page details --csv > gpurun_out/r02_ncu_attn_details.csv 2>/dev/null; wc -l gpurun_out/r02_ncu_attn_details.csv
