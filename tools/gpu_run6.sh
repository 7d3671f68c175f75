#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_baseline_configs.py tests/test_gpu_dataprep.py -q -m gpu -s 2>&1 | grep -E "passed|failed|FAILED|Error|\[boundary|\[config2 mixed|assert|^E " | tail -30
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench3.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"])
print("mixed", json.dumps(d.get("mixed_precision"))[:400])
t = d.get("ddp_train") or {}
print("ddp_train", {k: t.get(k) for k in ("img_per_s", "ms_per_step", "frac_of_peak", "error")})
print("sampler", d["extra"]["config4_sampler"]["ms_per_diffusion_step"], "hier", d["extra"]["config3_hierarchy_forward_step"]["img_per_s"])
print("eager", json.dumps(d.get("gpu_eager_baseline"))[:700])
PY
timeout 600 python tools/bench_conv_grad.py 2>&1 | tail -14 | tee gpurun_out/r2_conv_grad_bench.txt
