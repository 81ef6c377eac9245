mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parse_smpl.py tests/test_gpu_e2e.py -q -x -m gpu 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu_u.log
tail -3 gpurun_out/r02_pytest_gpu_u.log
timeout 300 python bench.py --workload smpl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_u_smpl.json 2> gpurun_out/r02_bench_u_smpl.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_u_smpl.json')); print('smpl', d['value'], d['ms_per_step'], d['roofline'])"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/launches_u_smpl.csv python bench.py --workload smpl --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_u_smpl.log 2>&1
python tools/launch_summary.py gpurun_out/launches_u_smpl.csv --steps 3 2>/dev/null | head -8
