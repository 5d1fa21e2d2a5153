#!/bin/bash
# A/B of library variants on one box: tools/gpu_ab.sh lib1.so lib2.so ...   (paths relative to the repo root)
# FULL=1: the first variant also runs the whole GPU test suite; otherwise only the verify parity tests.
first=1
for lib in "$@"; do
  export KYBER_B200_LIB=$PWD/$lib
  echo "== $lib"
  if [ $first = 1 ]; then
    if [ "$FULL" = 1 ]; then timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
    else timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "verify or sig or eddsa or schnorr" 2>&1 | tail -3; fi
    first=0
  fi
  for r in 1 2; do
  python bench.py --no-extras --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('verify', d['value'], 'e2e', d['e2e']['value'], d['roofline']['kernels_ms'])"
  done
  if [ "$ALL" = 1 ]; then
  python tools/bench_dkg.py --reps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dkg', d['round_ms'], d['verdicts_match_expected'])"
  python tools/msm_timing.py 2>/dev/null | tail -1
  fi
done
