mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_conv_tf32.py -q -x -m gpu 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_k.log
tail -6 gpurun_out/r02_pytest_gpu_k.log
timeout 400 python bench.py --steps 20 --warmup 3 --no-extra > gpurun_out/r02_bench_k_concat.json 2> gpurun_out/r02_bench_k.err
B200ROMP_NO_SKIP_CONCAT=1 timeout 400 python bench.py --steps 20 --warmup 3 --no-extra > gpurun_out/r02_bench_k_noconcat.json 2>> gpurun_out/r02_bench_k.err
python -c "
import json
for n in ('concat','noconcat'):
    d=json.load(open('gpurun_out/r02_bench_k_%s.json'%n)); print(n, round(d['value']), round(d['e2e']['value']), d['roofline']['achieved'])"
timeout 300 python tools/op_profile.py --precision bf16 > gpurun_out/r02_op_profile_bf16_k.md 2> gpurun_out/op_k.err; sed -n 1,12p gpurun_out/r02_op_profile_bf16_k.md
