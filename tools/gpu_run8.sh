#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_gn.py -q -m gpu -s -x 2>&1 | grep -E "passed|failed|FAILED|Error|\[fused|assert|^E " | tail -20
for f in 1 0; do T2H_FUSE_GN=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-extra --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fuse_gn', $f, 'img/s', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['clocks'])"; done
T2H_FUSE_GN=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-extra --no-cpu-baseline --precision fp16 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fp16 fuse 1 img/s', d['value'])"
T2H_FUSE_GN=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-extra --no-cpu-baseline --precision fp16 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fp16 fuse 0 img/s', d['value'])"
