#!/bin/bash
# GPU box: the round-2 ncu evidence (final build of the round).  Launch lists of the bench command and of the DKG / MSM
# targets, one `--set full` capture per hot kernel, exported as CSV (the .ncu-rep files stay on the box: gpurun_out/ is capped).
set -x
mkdir -p gpurun_out
T=/tmp/ncu_r2; mkdir -p $T
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2_prof_bench.json 2> gpurun_out/r2_prof_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_dkg.csv python tools/bench_dkg.py --reps 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_msm22.csv python tools/profile_msm.py 22 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_msm17.csv python tools/profile_msm.py 17 > /dev/null 2>&1
# --set full, one capture per kernel family
ncu --set full --clock-control none --import-source on -k regex:k_verify_half -s 2 -c 2 -f -o $T/verify python tools/profile_target.py 20 > /dev/null 2>&1
ncu -i $T/verify.ncu-rep --page raw --csv > gpurun_out/r2_ncu_verify_raw.csv
ncu -i $T/verify.ncu-rep --page source --csv --kernel-name regex:k_verify_half_main > gpurun_out/r2_ncu_verify_main_source.csv 2>/dev/null
ncu -i $T/verify.ncu-rep --page source --csv --kernel-name regex:k_verify_half_prep > gpurun_out/r2_ncu_verify_prep_source.csv 2>/dev/null
ncu --set full --clock-control none -k regex:"k_msm_accum|k_msm_prepare|k_msm_window_sums|k_msm_finish" -s 12 -c 4 -f -o $T/msm python tools/profile_msm.py 22 > /dev/null 2>&1
ncu -i $T/msm.ncu-rep --page raw --csv > gpurun_out/r2_ncu_msm_raw.csv
ls -la gpurun_out/r2_ncu_* gpurun_out/r2_launches_*
