#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_grad.py -q -m gpu -s 2>&1 | tail -120 > gpurun_out/r2_conv_grad.log
timeout 600 python -m pytest tests/test_gpu_vqgan_train.py -q -m gpu -s -x 2>&1 | tail -120 > gpurun_out/r2_vqgan_train.log
timeout 600 python -m pytest tests/test_gpu_baseline_configs.py -q -m gpu -s -k sample_fn 2>&1 | tail -60 > gpurun_out/r2_sample_fn.log
timeout 300 python -m pytest tests/test_gpu_modules.py -q -m gpu -k sampler 2>&1 | tail -30 > gpurun_out/r2_sampler_mod.log
tail -40 gpurun_out/r2_conv_grad.log; tail -60 gpurun_out/r2_vqgan_train.log; tail -30 gpurun_out/r2_sample_fn.log; tail -8 gpurun_out/r2_sampler_mod.log
