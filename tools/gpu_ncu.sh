#!/bin/bash
# r02 ncu captures: weight-gradient kernel (pair mode, 128->128 3x3 @512x256, B=8, single product) and the swapped forward / data-gradient kernel
mkdir -p gpurun_out
CASES=0 TERMS=1 ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -c 2 -f -o gpurun_out/r02_wgrad_fp16 python tools/bench_conv_grad.py > gpurun_out/r02_ncu_wgrad.log 2>&1
CASES=0 TERMS=2 ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -c 2 -f -o gpurun_out/r02_wgrad_fp32 python tools/bench_conv_grad.py >> gpurun_out/r02_ncu_wgrad.log 2>&1
CASES=0 TERMS=1 ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tapgemm_swap -c 2 -f -o gpurun_out/r02_swap_fp16 python tools/bench_conv_grad.py >> gpurun_out/r02_ncu_wgrad.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:norm_bwd -c 4 -f -o gpurun_out/r02_norm_bwd python tools/profile_train_step.py --batch 4 >> gpurun_out/r02_ncu_wgrad.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
CASES=0,1,2 timeout 300 python tools/bench_conv_grad.py 2>&1 | tail -8
