#!/bin/bash
mkdir -p gpurun_out
python bench.py --impl reference > gpurun_out/r2_bench_n1_reference.json 2>/dev/null
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python tools/bench_dkg.py --n 256 --t 171 --reps 5 > gpurun_out/r2_vss_cfg3_n1.json 2>/dev/null
T=/tmp/ncu_r2; mkdir -p $T
ncu --set full --clock-control none -k regex:"k_msm_accum|k_msm_prepare|k_msm_window_sums|k_msm_finish" -s 4 -c 4 -f -o $T/msm python tools/profile_msm.py 22 > /dev/null 2>&1
ncu -i $T/msm.ncu-rep --page raw --csv > gpurun_out/r2_ncu_msm_raw.csv
ncu --set full --clock-control none -k regex:"k_fd_steps|k_fd_check" -s 2 -c 2 -f -o $T/fd python tools/profile_dkg.py 1024 683 > /dev/null 2>&1
ncu -i $T/fd.ncu-rep --page raw --csv > gpurun_out/r2_ncu_fd_raw.csv
ls -la gpurun_out/r2_ncu_msm_raw.csv gpurun_out/r2_ncu_fd_raw.csv
head -c 300 gpurun_out/r2_bench_n1.json
