mkdir -p gpurun_out
timeout 400 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:smpl_blend_tc_kernel" -s 2 -c 1 -f -o gpurun_out/prof_smpl_blend2 python bench.py --workload smpl --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_smpl_blend2.log 2>&1
ls -la gpurun_out/prof_smpl_blend2.ncu-rep
