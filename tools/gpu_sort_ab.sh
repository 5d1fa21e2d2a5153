#!/bin/bash
# records sorted by loop length (KB_VERIFY_SORT, default 1) against the unsorted order, same library; verify parity tests first
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "verify or sig or eddsa or schnorr" 2>&1 | tail -3
run() { python bench.py --no-extras --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('verify', round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3), 'ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernels_ms'].items() if k.startswith('k_')})"; }
for s in 1 0 1 0; do echo "== KB_VERIFY_SORT=$s"; KB_VERIFY_SORT=$s run; done
