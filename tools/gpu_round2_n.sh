mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fuse_sum.py tests/test_gpu_e2e.py -q -x -m gpu 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu_n.log
tail -4 gpurun_out/r02_pytest_gpu_n.log
timeout 300 python tools/op_profile.py --precision bf16 > gpurun_out/r02_op_profile_bf16_n.md 2> gpurun_out/op_n.err; grep "s2  128->64 \|s2  256->64" gpurun_out/r02_op_profile_bf16_n.md | cut -c1-250
timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_n.json 2> gpurun_out/r02_bench_n.err
B200ROMP_NO_S2_KSPLIT=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_n_nosplit.json 2>> gpurun_out/r02_bench_n.err
python -c "
import json
for n in ('n','n_nosplit'):
    d=json.load(open('gpurun_out/r02_bench_%s.json'%n)); print(n, round(d['value']), round(d['e2e']['value']), d['roofline']['achieved'])"
