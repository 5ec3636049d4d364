set -x
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02F_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r02F_bench_n1.json 2> gpurun_out/r02F_bench_n1.err
timeout 300 bash tools/ncu_launches.sh r02F
for spec in "1, 0, 0, 0:psss_plain" "1, 0, 3, 0:ps_S2s" "1, 0, 2, 1:psdp" "2, 1, 1, 1:dppp"; do
  cls=${spec%%:*}; name=${spec##*:}
  pat=$(echo "$cls" | sed 's/\([0-9]\)/\\(int\\)\1/g')
  timeout 240 ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
      -k "regex:eri_class_kernel<${pat}, \(int\)1" -c 1 -o gpurun_out/r02F_${name} \
      python tools/profile_direct.py w32_ccpvdz 1 > gpurun_out/r02F_${name}.log 2>&1
  ncu -i gpurun_out/r02F_${name}.ncu-rep --page raw --csv > gpurun_out/raw_r02F_${name}.csv 2>/dev/null
  rm -f gpurun_out/r02F_${name}.ncu-rep
done
timeout 240 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:screen_kernel" -s 20 -c 3 -o gpurun_out/r02F_screen python tools/profile_direct.py w32_ccpvdz 1 > gpurun_out/r02F_screen.log 2>&1
ncu -i gpurun_out/r02F_screen.ncu-rep --page raw --csv > gpurun_out/raw_r02F_screen.csv 2>/dev/null; rm -f gpurun_out/r02F_screen.ncu-rep
tail -2 gpurun_out/r02F_pytest.log
