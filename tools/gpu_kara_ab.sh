#!/bin/bash
export KYBER_B200_LIB=$PWD/kyber-rs_b200/libkyber_b200_kara.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for lib in kyber-rs_b200/libkyber_b200.so kyber-rs_b200/libkyber_b200_kara.so; do
  export KYBER_B200_LIB=$PWD/$lib
  echo "== $lib"
  python bench.py --no-extras --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('verify', d['value'], d['roofline']['kernels_ms'])"
  python tools/bench_dkg.py --reps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dkg', d['round_ms'], d['verdicts_match_expected'])"
  python tools/msm_timing.py 2>/dev/null | tail -1
done
