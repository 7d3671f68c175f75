#!/bin/bash
# round-2 GPU check 1: new training kernels + full-size parity + bench with the eager baseline
mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv_grad.py -x -q -m gpu -s 2>&1 | tail -60 > gpurun_out/r2_conv_grad.log
python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_modules.py -q -m gpu -s 2>&1 | tail -80 > gpurun_out/r2_configs.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -5 gpurun_out/r2_conv_grad.log; tail -30 gpurun_out/r2_configs.log; head -c 3000 gpurun_out/r2_bench1.json
