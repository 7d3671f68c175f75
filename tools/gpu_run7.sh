#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_vqgan_train.py tests/test_gpu_conv_grad.py -q -m gpu -s -k "not sample_step" 2>&1 | grep -E "passed|failed|FAILED|Error|\[vqgan|assert|^E " | tail -12
timeout 600 python -m pytest tests/test_gpu_ddp.py -q -m gpu -s 2>&1 | grep -E "passed|failed|skipped|FAILED|Error|\[ddp|assert|^E " | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_2gpu.json"))
    print("N=2 value", d["value"], "e2e", d["e2e"]["value"])
    print("ddp_train", json.dumps(d.get("ddp_train"))[:1400])
except Exception as e:
    print("bench parse failed", e)
PY
tail -4 gpurun_out/r2_bench_2gpu.err
