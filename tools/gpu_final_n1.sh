#!/bin/bash
# GPU box (1 GPU): the whole GPU test suite and the files kept under profiles/ for round 2 (final build of the round)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_gpu_tests.log
python bench.py --impl reference > gpurun_out/r2_bench_n1_reference.json 2>/dev/null
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python tools/bench_dkg.py --reps 3 > gpurun_out/r2_dkg_round_n1.json 2>/dev/null
python tools/bench_dkg.py --reps 3 --shard-of 8 > gpurun_out/r2_dkg_shard_of_8.json 2>/dev/null
KB_DKG_FD=0 python tools/bench_dkg.py --reps 1 > gpurun_out/r2_dkg_round_n1_horner.json 2>/dev/null
python tools/bench_dkg.py --n 256 --t 171 --reps 5 > gpurun_out/r2_vss_cfg3_n1.json 2>/dev/null
python tools/quick_bench.py 20 > gpurun_out/r2_quick_bench_n1.json 2>/dev/null
python tools/msm_timing.py > gpurun_out/r2_msm_timing.txt 2>/dev/null
python tools/e2e_sweep.py > gpurun_out/r2_e2e_sweep.jsonl 2>/dev/null
for f in r2_bench_n1_reference r2_bench_n1 r2_dkg_round_n1 r2_dkg_shard_of_8 r2_dkg_round_n1_horner r2_vss_cfg3_n1; do echo $f; head -c 700 gpurun_out/$f.json; echo; done
cat gpurun_out/r2_msm_timing.txt
