mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parse_smpl.py -q -x -m gpu 2>&1 | tail -15 > gpurun_out/r02_pytest_smpl_f.log
tail -15 gpurun_out/r02_pytest_smpl_f.log
timeout 300 python bench.py --workload smpl --steps 5 --warmup 3 > gpurun_out/r02_bench_f_smpl.json 2> gpurun_out/r02_bench_f_smpl.err
tail -3 gpurun_out/r02_bench_f_smpl.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_f_smpl.json'));print('smpl TC', d['ms_per_step'], d['roofline']['frac'])"
B200ROMP_SMPL_SIMT=1 timeout 300 python bench.py --workload smpl --steps 3 --warmup 3 > gpurun_out/r02_bench_f_smpl_simt.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/r02_bench_f_smpl_simt.json'));print('smpl SIMT', d['ms_per_step'])"
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu_f.log
tail -12 gpurun_out/r02_pytest_gpu_f.log
timeout 300 python bench.py --workload bev --steps 10 --warmup 3 > gpurun_out/r02_bench_f_bev.json 2> gpurun_out/r02_bench_f_bev.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_f_bev.json'));print('bev', d['value'], d['e2e']['value'], d['config']['persons_planted'], d['config']['persons_out'])"
timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_f.json'));print('bf16', d['value'], d['e2e']['value'], d['roofline']['achieved'])"
M=gpu__time_duration.sum
timeout 200 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/launches_f.csv python tools/ncu_step.py --precision bf16 --steps 1 > gpurun_out/ncu_f.log 2>&1
python tools/launch_summary.py gpurun_out/launches_f.csv 2>/dev/null | grep -E "smpl|parse|project|total"
