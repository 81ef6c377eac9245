mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu_d.log
tail -12 gpurun_out/r02_pytest_gpu_d.log
timeout 300 python bench.py --workload bev --steps 10 --warmup 3 > gpurun_out/r02_bench_d_bev.json 2> gpurun_out/r02_bench_d_bev.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_d_bev.json'));print('bev', d['value'], d['e2e']['value'], d['config'])"
B200ROMP_BEV_CENTER3D_2PASS=1 timeout 300 python bench.py --workload bev --steps 10 --warmup 3 > gpurun_out/r02_bench_d_bev_2pass.json 2> /dev/null
python -c "import json;d=json.load(open('gpurun_out/r02_bench_d_bev_2pass.json'));print('bev 2pass', d['value'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --precision tf32 > gpurun_out/r02_bench_d_tf32.json 2> gpurun_out/r02_bench_d_tf32.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_d_tf32.json'));print('tf32', d['value'], d['e2e']['value'], d['roofline']['achieved'])"
M=gpu__time_duration.sum
timeout 300 ncu --metrics $M --clock-control none -s 2000 -c 1500 --csv --log-file gpurun_out/launches_bev.csv python bench.py --workload bev --steps 2 --warmup 3 > gpurun_out/ncu_bev.log 2>&1
python tools/launch_summary.py gpurun_out/launches_bev.csv | head -24
