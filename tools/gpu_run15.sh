#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py tests/test_gpu_indexpred.py tests/test_gpu_attn_fused.py tests/test_gpu_train.py tests/test_gpu_baseline_configs.py -x -q 2>&1 | tail -12
timeout 300 python tools/sampler_timeline.py fp32 2>&1 | tail -45 > gpurun_out/r02_sampler_timeline_v1.txt; sed -n 1,14p gpurun_out/r02_sampler_timeline_v1.txt; tail -3 gpurun_out/r02_sampler_timeline_v1.txt
timeout 300 python tools/bench_sampler.py fp32 64 2>&1 | tail -2
