#!/bin/bash
# new-kernel hygiene: split-K unit tests, then compute-sanitizer memcheck over the fused attention / split-K /
# epilogue paths (small shapes only: the sanitizer slows kernels 10-50x)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_splitk.py -x -q 2>&1 | tail -3
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_attn_fused.py tests/test_gpu_splitk.py -x -q -k "not logits_equal" > gpurun_out/r02_sanitizer_memcheck_attn_splitk.log 2>&1; tail -6 gpurun_out/r02_sanitizer_memcheck_attn_splitk.log
