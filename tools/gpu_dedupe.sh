#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vqgan_train.py tests/test_gpu_boundary.py tests/test_gpu_ddp.py -x -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_final2.json')); t=d['ddp_train']
print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks'], 'mixed', d['mixed_precision']['img_per_s'], 'train', t['ms_per_step'], t['img_per_s'], t['launches_per_step'], 'sampler', d['extra']['config4_sampler']['ms_per_diffusion_step'])"
