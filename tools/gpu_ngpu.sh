#!/bin/bash
# N-GPU box: the bench exactly as the driver launches it for N = number of visible GPUs
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('N=$N value', d['value'], 'e2e', d['e2e']['value']); t=d['ddp_train']; print({k: t[k] for k in ('ms_per_step','img_per_s','ms_per_step_without_allreduce','exposed_allreduce_ms','frac_of_peak','per_gpu_batch','micro_batch')})" || tail -5 gpurun_out/r2_bench_${N}gpu.err
