#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py tests/test_gpu_attn_fused.py tests/test_gpu_train.py tests/test_gpu_conv_grad.py -x -q 2>&1 | tail -4
timeout 300 python tools/sampler_timeline.py fp32 2>&1 | tail -45 > gpurun_out/r02_sampler_timeline_v2.txt; sed -n 7,13p gpurun_out/r02_sampler_timeline_v2.txt; tail -2 gpurun_out/r02_sampler_timeline_v2.txt
timeout 300 python tools/bench_sampler.py fp32 64 2>&1 | tail -1
# ncu full capture of the fc1 GEMM (planes + GELU epilogue) of one layer: 3 warm-up forwards x 97 tap-GEMM launches, then qkv, proj, fc1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:tapgemm_kernel --launch-skip 293 --launch-count 1 -o gpurun_out/r02_ncu_sampler_fc1 -f python tools/profile_sampler_step.py fp32 > gpurun_out/r02_ncu_sampler_fc1.log 2>&1; tail -2 gpurun_out/r02_ncu_sampler_fc1.log
