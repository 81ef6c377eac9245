mkdir -p gpurun_out
timeout 200 python tools/smpl_blend_debug.py 2>&1 | grep -E "verts max|rror" | head -4
python bench.py --workload smpl --steps 5 --warmup 3 2>/dev/null > gpurun_out/r02_bench_g_smpl.json; python -c "import json;d=json.load(open('gpurun_out/r02_bench_g_smpl.json'));print('smpl', d['ms_per_step'], d['roofline']['frac'])"
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_smpl.csv python bench.py --workload smpl --steps 1 --warmup 3 > /dev/null 2>&1; python tools/launch_summary.py gpurun_out/launches_smpl.csv > gpurun_out/r02_launches_smpl_g.md; head -8 gpurun_out/r02_launches_smpl_g.md
# full-set captures of the dominant conv kernels (shared-memory roofline evidence)
B200ROMP_NO_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_tc2_kernel<.int.64, .int.64, .int.2' -s 40 -c 1 -f -o gpurun_out/prof_c64 python tools/ncu_step.py --precision bf16 --steps 1 > gpurun_out/ncu_c64.log 2>&1
B200ROMP_NO_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_tc2_kernel<.int.32, .int.32, .int.2' -s 40 -c 1 -f -o gpurun_out/prof_c32 python tools/ncu_step.py --precision bf16 --steps 1 > gpurun_out/ncu_c32.log 2>&1
ls -la gpurun_out/*.ncu-rep
