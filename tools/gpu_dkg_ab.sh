#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "dkg or cfg4" 2>&1 | tail -2
for w in 0 1; do for p in 3 4; do echo "wide=$w parts=$p"; KB_FD_STEPS_WIDE=$w KB_FD_PARTS=$p python tools/bench_dkg.py --reps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['round_ms'], d['verdicts_match_expected'])"; done; done
for w in 0 1; do echo "shard8 wide=$w"; KB_FD_STEPS_WIDE=$w python tools/bench_dkg.py --reps 3 --shard-of 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['round_ms'], d['verdicts_match_expected'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_dkg_launches_b.csv python tools/bench_dkg.py --reps 1 > /dev/null 2>&1
python tools/launch_list.py gpurun_out/r2_dkg_launches_b.csv 2>/dev/null | head -8
