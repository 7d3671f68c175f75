#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_pytest_final.log; tail -6 gpurun_out/r2_pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench4.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"], "launches", d["gpu_launches"])
print("mixed", (d.get("mixed_precision") or {}).get("img_per_s"))
t = d.get("ddp_train") or {}
print("ddp_train", {k: t.get(k) for k in ("img_per_s", "ms_per_step", "frac_of_peak", "error", "tensor_kernels")})
print("sampler", d["extra"]["config4_sampler"]["ms_per_diffusion_step"], "hier", d["extra"]["config3_hierarchy_forward_step"]["img_per_s"])
e = d.get("gpu_eager_baseline") or {}
print("eager", {k: (v.get("img_per_s") or v.get("ms_per_forward") or v.get("ms_per_diffusion_step")) if isinstance(v, dict) else v for k, v in e.items()})
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | head -c 400
