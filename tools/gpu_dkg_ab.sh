#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "dkg or cfg4 or cfg3 or proto" 2>&1 | tail -3
for so in 1 2 4 8; do for p in 0 1 2 3 4; do echo -n "cfg3 shard-of $so parts=$p: "; KB_FD_PARTS=$p timeout 120 python tools/bench_dkg.py --n 256 --t 171 --reps 5 --shard-of $so 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['round_ms'],3), d['verdicts_match_expected'])"; done; done
for so in 1 8; do for r in 1 2; do echo -n "cfg4 shard-of $so: "; timeout 120 python tools/bench_dkg.py --reps 3 --shard-of $so 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['round_ms'], d['verdicts_match_expected'])"; done; done
