#!/bin/bash
# fused norm-backward sums: unit test, the trainer / boundary tests, then the training step with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_grad.py -x -q -s -k "norm_backward_sums" 2>&1 | grep -E "\[nb\]|passed|failed|Error|error" | tail -8
timeout 900 python -m pytest tests/test_gpu_vqgan_train.py tests/test_gpu_boundary.py tests/test_gpu_conv_grad.py -x -q 2>&1 | tail -3
for f in 1 0; do
T2H_FUSE_NB=$f timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_bench_nb$f.json 2> gpurun_out/r2_bench_nb$f.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_nb$f.json')); t=d['ddp_train']; print('FUSE_NB=$f', 'value', round(d['value'],1), 'train', round(t['ms_per_step'],1), 'ms', round(t['img_per_s'],1), 'img/s', 'launches', t['launches_per_step'], d['clocks'])"
done
