#!/bin/bash
# A/B of the split preparation (KB_VERIFY_SPLIT) on one box: parity tests with the split forced for every batch size, then timings
mkdir -p gpurun_out
KB_VERIFY_SPLIT=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "verify or sig or eddsa or schnorr" 2>&1 | tail -3
run() { python bench.py --no-extras --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('verify', round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3), {k: round(v,3) for k,v in d['roofline']['kernels_ms'].items() if k.startswith('k_')})"; }
echo "== split 0"; KB_VERIFY_SPLIT=0 run
for b in 1 2 3; do for pb in 5 4 6; do echo "== split 1 blocks $b pbound $pb"; KB_VERIFY_SPLIT=1 KB_VERIFY_SPLIT_BLOCKS=$b KB_VERIFY_SPLIT_PBOUND=$pb run; done; done
