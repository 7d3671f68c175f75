#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_sampler_launches.csv python tools/profile_sampler_step.py fp32 > /dev/null 2>&1
python tools/ncu_by_shape.py gpurun_out/r02_sampler_launches.csv > gpurun_out/r02_sampler_launches_by_shape.txt; cat gpurun_out/r02_sampler_launches_by_shape.txt
