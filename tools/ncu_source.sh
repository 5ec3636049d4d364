#!/bin/bash
# usage: tools/ncu_source.sh "1, 1, 1, 0" tag  -> gpurun_out/src_<tag>.csv.gz (ncu source page of the DIGEST kernel of that class)
cls=$(echo "$1" | sed 's/\([0-9]\)/\\(int\\)\1/g')
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:eri_class_kernel<${cls}, \(int\)1>" -c 1 -o /tmp/prof_$2 python tools/profile_direct.py ${3:-w32_ccpvdz} 1 > gpurun_out/prof_$2.log 2>&1
ncu -i /tmp/prof_$2.ncu-rep --page source --csv 2>/dev/null | gzip -c > gpurun_out/src_$2.csv.gz
ncu -i /tmp/prof_$2.ncu-rep --page raw --csv > gpurun_out/raw_$2.csv 2>/dev/null
ls -la gpurun_out/src_$2.csv.gz
