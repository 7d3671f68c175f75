#!/bin/bash
# fused attention kernel: unit tests, then the config-4 sampler with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attn_fused.py -x -q -s 2>&1 | grep -v "^$" | tail -30
echo "--- fused"; timeout 300 python tools/bench_sampler.py fp32 32 2>&1 | tail -2
echo "--- three-launch"; T2H_FUSED_ATTN=0 timeout 300 python tools/bench_sampler.py fp32 32 2>&1 | tail -2
echo "--- fp16 fused"; timeout 300 python tools/bench_sampler.py fp16 32 2>&1 | tail -2
