#!/bin/bash
# GPU box: forward-difference DKG round — parity tests, then the round at BASELINE configs 3/4 for each block count.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "dkg or cfg4 or cfg3 or pubpoly or abi" > gpurun_out/r2_dkg_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2_dkg_tests.log
for p in 2 3 4; do KB_FD_PARTS=$p python tools/bench_dkg.py --reps 3 > gpurun_out/r2_dkg_p$p.json 2> gpurun_out/r2_dkg_p$p.err; cat gpurun_out/r2_dkg_p$p.json; done
python tools/bench_dkg.py --reps 3 > gpurun_out/r2_dkg_auto.json 2> gpurun_out/r2_dkg_auto.err; cat gpurun_out/r2_dkg_auto.json
for p in 3 4; do KB_FD_PARTS=$p python tools/bench_dkg.py --reps 3 --shard-of 8 > gpurun_out/r2_dkg_shard8_p$p.json 2>/dev/null; cat gpurun_out/r2_dkg_shard8_p$p.json; done
for p in 1 2; do KB_FD_PARTS=$p python tools/bench_dkg.py --n 256 --t 171 --reps 5 > gpurun_out/r2_vss_p$p.json 2>/dev/null; cat gpurun_out/r2_vss_p$p.json; done
KB_DKG_FD=0 python tools/bench_dkg.py --n 256 --t 171 --reps 5 > gpurun_out/r2_vss_horner.json 2>/dev/null; cat gpurun_out/r2_vss_horner.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_dkg_launches.csv python tools/bench_dkg.py --reps 1 > /dev/null 2>&1
python tools/launch_list.py gpurun_out/r2_dkg_launches.csv 2>/dev/null | head -30
