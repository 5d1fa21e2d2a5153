#!/bin/bash
# prints the DKG round times of a bench.py run (debug helper): tools/bench_cfg_rounds.sh [ENV=VALUE ...]
env "$@" python bench.py --steps 2 --warmup 3 --msm-max-log2 18 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['configs']
print('verify', round(d['value']/1e6,2), 'cfg3', c['cfg3']['round_ms'], c['cfg3']['e2e']['round_ms'], 'cfg4', c['cfg4']['round_ms'], c['cfg4']['whole_round']['round_ms'])"
