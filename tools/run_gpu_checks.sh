#!/bin/bash
# everything the round-end driver runs on the GPU box, in one gpurun call
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
