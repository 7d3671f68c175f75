#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ddp.py -q -m gpu -s 2>&1 | grep -E "passed|failed|skipped|FAILED|Error|\[ddp|assert|^E " | tail -12
timeout 400 python -m pytest tests/test_gpu_fused_gn.py -q -m gpu -s 2>&1 | grep -E "passed|failed|FAILED|Error|\[fused|assert|^E " | tail -12
