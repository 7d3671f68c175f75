#!/bin/bash
# Round-end evidence in one gpurun call (1 GPU): full GPU test suite, smoke, default bench, the ncu launch lists of
# the bench command / one training micro-batch / the sampler forward, the sampler's per-launch timeline and the
# launch-latency floor.  Outputs land in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_pytest_final.log; tail -3 gpurun_out/r2_pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
timeout 1200 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_final.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks'], 'mixed', (d.get('mixed_precision') or {}).get('img_per_s'), 'ddp', (d.get('ddp_train') or {}).get('img_per_s'), 'sampler', d['extra']['config4_sampler'])
print({k: v for k, v in d['gpu_eager_baseline'].items() if k.startswith('config5')})"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_step_launches.csv python bench.py --steps 2 --warmup 3 --no-train --no-extra --no-cpu-baseline > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_step_launches.csv > gpurun_out/r02_step_launches.txt; head -14 gpurun_out/r02_step_launches.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_train_launches.csv python tools/profile_train_step.py --batch 8 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_train_launches.csv > gpurun_out/r02_train_launches.txt; head -12 gpurun_out/r02_train_launches.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_sampler_launches.csv python tools/profile_sampler_step.py fp32 > /dev/null 2>&1
python tools/ncu_by_shape.py gpurun_out/r02_sampler_launches.csv > gpurun_out/r02_sampler_launches_by_shape.txt; cat gpurun_out/r02_sampler_launches_by_shape.txt
timeout 300 python tools/sampler_timeline.py fp32 2>&1 | tail -45 > gpurun_out/r02_sampler_timeline_final.txt; tail -3 gpurun_out/r02_sampler_timeline_final.txt
timeout 120 ./tools/launch_floor > gpurun_out/r02_launch_floor.txt 2>&1; head -4 gpurun_out/r02_launch_floor.txt
