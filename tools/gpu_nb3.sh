#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_grad.py -x -q -k "norm_backward_sums" 2>&1 | tail -2
for f in 1 0; do T2H_FUSE_NB=$f timeout 600 python tools/profile_train_step.py --events --batch 8 2>/dev/null | sed -n '1,4p;/by shape/,$p' > gpurun_out/r2_train_events_nb$f.txt; done
paste -d'|' <(cut -c1-75 gpurun_out/r2_train_events_nb1.txt) <(cut -c1-75 gpurun_out/r2_train_events_nb0.txt) | head -12
for f in 1 0; do
T2H_FUSE_NB=$f timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_bench_nb$f.json 2> gpurun_out/r2_bench_nb$f.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_nb$f.json')); t=d['ddp_train']; print('FUSE_NB=$f', 'value', round(d['value'],1), 'train', round(t['ms_per_step'],1), 'ms', round(t['img_per_s'],1), 'img/s', 'launches', t['launches_per_step'])"
done
