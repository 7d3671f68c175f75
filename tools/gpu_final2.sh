#!/bin/bash
# round-2 final evidence: full GPU test suite, smoke, default bench, ncu launch list of the bench command, sampler launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_pytest_final.log; tail -3 gpurun_out/r2_pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench5.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks'], 'mixed', (d.get('mixed_precision') or {}).get('img_per_s'), 'ddp', (d.get('ddp_train') or {}).get('img_per_s'), 'sampler', d['extra']['config4_sampler']['ms_per_diffusion_step'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_step_launches.csv python bench.py --steps 2 --warmup 3 --no-train --no-extra --no-cpu-baseline > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_step_launches.csv > gpurun_out/r02_step_launches.txt; head -14 gpurun_out/r02_step_launches.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_train_launches.csv python tools/profile_train_step.py --batch 8 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_train_launches.csv > gpurun_out/r02_train_launches.txt; head -12 gpurun_out/r02_train_launches.txt
