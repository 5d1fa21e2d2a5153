#!/bin/bash
mkdir -p gpurun_out
for lg in 17 22; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b_msm_launches_$lg.csv python tools/profile_msm.py $lg > /dev/null 2>&1
echo "== 2^$lg"; python tools/launch_list.py gpurun_out/r2b_msm_launches_$lg.csv | head -24
done
