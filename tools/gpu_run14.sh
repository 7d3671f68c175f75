#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sampler_timeline.py fp32 2>&1 | tail -45 | tee gpurun_out/r02_sampler_timeline.txt
