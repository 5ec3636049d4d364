#!/bin/bash
# usage: tools/ncu_classes.sh <tag-prefix> "2, 1, 1, 0:dpps" "2, 1, 1, 1:dppp" ...
# one `ncu --set full --import-source on` capture of the block-digestion kernel of each listed class
# (first matching launch of one (H2O)32 direct build) -> gpurun_out/<prefix>_<name>.ncu-rep
pre=$1; shift
for spec in "$@"; do
  cls=${spec%%:*}; name=${spec##*:}
  pat=$(echo "$cls" | sed 's/\([0-9]\)/\\(int\\)\1/g')
  ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
      -k "regex:eri_(class|team)_kernel<${pat}, \(int\)1(, \(bool\)${FAR:-0})?>" -c 1 -o gpurun_out/${pre}_${name} \
      python tools/profile_direct.py ${WORKLOAD:-w32_ccpvdz} 1 > gpurun_out/${pre}_${name}.log 2>&1
  ls -la gpurun_out/${pre}_${name}.ncu-rep
done
