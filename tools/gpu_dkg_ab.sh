#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "dkg or cfg4" 2>&1 | tail -2
for w in 0 4 1; do echo "full wide=$w"; KB_FD_STEPS_WIDE=$w python tools/bench_dkg.py --reps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['round_ms'], d['verdicts_match_expected'])"; done
for w in 0 4 1; do for so in 8 4 2; do echo "shard-of $so wide=$w"; KB_FD_STEPS_WIDE=$w python tools/bench_dkg.py --reps 3 --shard-of $so 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['round_ms'], d['verdicts_match_expected'])"; done; done
