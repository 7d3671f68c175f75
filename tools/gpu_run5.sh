#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vqgan_train.py tests/test_gpu_boundary.py -q -m gpu -s 2>&1 | grep -E "passed|failed|FAILED|Error|\[boundary|\[vqgan|assert" | tail -30
timeout 600 python tools/profile_train_step.py --events --batch 8 > gpurun_out/r2_train_events.log 2>&1; tail -34 gpurun_out/r2_train_events.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_train_launches.csv python tools/profile_train_step.py --batch 8 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r2_train_launches.csv > gpurun_out/r2_train_launches.txt; head -40 gpurun_out/r2_train_launches.txt
for st in 1 2 3; do timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-extra --no-cpu-baseline --streams $st 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('streams', $st, 'img/s', d['value'], 'e2e', d['e2e']['value'])"; done
