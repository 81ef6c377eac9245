mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fuse_sum.py tests/test_gpu_e2e.py tests/test_gpu_bev.py -q -x -m gpu 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu_p.log
tail -4 gpurun_out/r02_pytest_gpu_p.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_p.json 2> gpurun_out/r02_bench_p.err
B200ROMP_NO_FUSE1X1_MERGE=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_p_nomerge.json 2>> gpurun_out/r02_bench_p.err
python -c "
import json
for n in ('p','p_nomerge'):
    d=json.load(open('gpurun_out/r02_bench_%s.json'%n)); print(n, round(d['value']), round(d['e2e']['value']), d['roofline']['achieved'], d.get('gpu_launches'))"
