#!/bin/bash
# GPU box (1 GPU): final build of round 2 — ncu evidence of the verify step first (bench.py reports the DRAM traffic measured
# here), then the whole GPU test suite and the files kept under profiles/
mkdir -p gpurun_out
T=/tmp/ncu_r2; mkdir -p $T
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2_prof_bench.json 2> gpurun_out/r2_prof_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_verify_half|k_half_sort" -s 5 -c 5 -f -o $T/verify python tools/profile_target.py 20 > /dev/null 2>&1
ncu -i $T/verify.ncu-rep --page raw --csv > gpurun_out/r2_ncu_verify_raw.csv
python tools/ncu_traffic.py gpurun_out/r2_ncu_verify_raw.csv > gpurun_out/r2_verify_traffic.json && cp gpurun_out/r2_verify_traffic.json profiles/r2_verify_traffic.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --impl reference > gpurun_out/r2_bench_n1_reference.json 2>/dev/null
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python tools/quick_bench.py 20 > gpurun_out/r2_quick_bench_n1.json 2>/dev/null
python tools/e2e_sweep.py > gpurun_out/r2_e2e_sweep.jsonl 2>/dev/null
head -c 400 gpurun_out/r2_bench_n1.json; echo; cat gpurun_out/r2_verify_traffic.json
