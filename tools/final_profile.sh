#!/bin/bash
# Round-end evidence on one B200: GPU test suite, the bench line (both arms), the ncu launch list + DRAM bytes of one
# build, and `ncu --set full` of the kernels named in profiles/r02_ncu_final_summary.txt.  Writes into gpurun_out/.
#   gpurun --timeout 1800 -- 'bash tools/final_profile.sh'
tag=${1:-r02F}
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${tag}_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
timeout 300 bash tools/ncu_launches.sh ${tag}
if [ "${FULL:-1}" = "1" ]; then
for spec in "1, 0, 0, 0:psss_plain" "1, 0, 3, 0:ps_S2s" "1, 0, 2, 1:psdp" "2, 1, 1, 1:dppp"; do
  cls=${spec%%:*}; name=${spec##*:}
  pat=$(echo "$cls" | sed 's/\([0-9]\)/\\(int\\)\1/g')
  timeout 240 ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
      -k "regex:eri_class_kernel<${pat}, \(int\)1" -c 1 -o gpurun_out/${tag}_${name} \
      python tools/profile_direct.py w32_ccpvdz 1 > gpurun_out/${tag}_${name}.log 2>&1
  ncu -i gpurun_out/${tag}_${name}.ncu-rep --page raw --csv > gpurun_out/raw_${tag}_${name}.csv 2>/dev/null
  rm -f gpurun_out/${tag}_${name}.ncu-rep
done
timeout 240 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:screen_kernel" -s 20 -c 3 -o gpurun_out/${tag}_screen python tools/profile_direct.py w32_ccpvdz 1 > gpurun_out/${tag}_screen.log 2>&1
ncu -i gpurun_out/${tag}_screen.ncu-rep --page raw --csv > gpurun_out/raw_${tag}_screen.csv 2>/dev/null; rm -f gpurun_out/${tag}_screen.ncu-rep
fi
tail -2 gpurun_out/${tag}_pytest.log
