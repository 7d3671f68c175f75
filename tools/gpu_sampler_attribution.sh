#!/bin/bash
# where a sampler step's time goes: the tap-GEMM kernel with parts of its work switched off (T2H_DEBUG bits:
# 1 skip epilogue math/stores, 2 skip MMA issue, 4 skip TMA loads) -- results are garbage, timings are the point
for d in 0 1 2 4 7; do echo "--- T2H_DEBUG=$d"; T2H_DEBUG=$d timeout 300 python tools/bench_sampler.py fp32 32 2>&1 | tail -2; done
echo "--- T2H_PDL=0"; T2H_PDL=0 timeout 300 python tools/bench_sampler.py fp32 32 2>&1 | tail -2
