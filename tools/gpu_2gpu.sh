#!/bin/bash
# 2-GPU box: the DDP equivalence test and the bench exactly as the driver launches it for N=2
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_ddp.py -x -q -s 2>&1 | grep -E "ddp\]|passed|failed" | tail -3
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu_v2.json 2> gpurun_out/r2_bench_2gpu_v2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_2gpu_v2.json').read().strip().splitlines()[-1])
print('N=2 value', d['value'], 'e2e', d['e2e']['value']); t=d['ddp_train']; print({k: t[k] for k in ('ms_per_step','img_per_s','ms_per_step_without_allreduce','exposed_allreduce_ms','frac_of_peak')})"
