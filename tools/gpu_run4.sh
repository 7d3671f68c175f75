#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s 2>&1 | tail -150 > gpurun_out/r2_pytest_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest_all.log | tail -20
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench2.json"))
print("value", d["value"], "e2e", d["e2e"]["value"])
print("ddp_train", json.dumps(d.get("ddp_train"))[:1500])
print("sampler", json.dumps(d["extra"].get("config4_sampler"))[:600])
print("hier", json.dumps(d["extra"].get("config3_hierarchy_forward_step"))[:300])
PY
tail -5 gpurun_out/r2_bench2.err
