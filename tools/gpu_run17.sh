#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py tests/test_gpu_indexpred.py tests/test_gpu_attn_fused.py tests/test_gpu_train.py tests/test_gpu_conv_grad.py tests/test_gpu_baseline_configs.py -x -q 2>&1 | tail -4
timeout 300 python tools/sampler_timeline.py fp32 2>&1 | tail -45 > gpurun_out/r02_sampler_timeline_v3.txt; sed -n 7,13p gpurun_out/r02_sampler_timeline_v3.txt; tail -2 gpurun_out/r02_sampler_timeline_v3.txt
timeout 300 python tools/bench_sampler.py fp32 64 2>&1 | tail -1
echo "--- T2H_PAIR=1"; T2H_PAIR=1 timeout 300 python tools/bench_sampler.py fp32 64 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 --no-train --no-extra --no-cpu-baseline > gpurun_out/r2_bench_quick.json 2> gpurun_out/r2_bench_quick.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_quick.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['clocks'])"
