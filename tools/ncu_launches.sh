#!/bin/bash
# usage: tools/ncu_launches.sh <tag> [workload]  -> gpurun_out/<tag>_launches.csv (+ dram bytes per launch)
tag=$1; wl=${2:-w32_ccpvdz}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled \
    --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_direct.py $wl 1 > gpurun_out/${tag}_launches.log 2>&1
