mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fuse_sum.py tests/test_gpu_e2e.py -q -x -m gpu 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu_m.log
tail -4 gpurun_out/r02_pytest_gpu_m.log
timeout 300 python tools/op_profile.py --precision bf16 > gpurun_out/r02_op_profile_bf16_m.md 2> gpurun_out/op_m.err; grep "^| sum" gpurun_out/r02_op_profile_bf16_m.md
timeout 300 python tools/op_profile.py --precision tf32 > gpurun_out/r02_op_profile_tf32_m.md 2> gpurun_out/op_m2.err; grep "^| sum" gpurun_out/r02_op_profile_tf32_m.md
timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_m.json 2> gpurun_out/r02_bench_m.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_m.json')); print(round(d['value']), round(d['e2e']['value']), d['roofline']['achieved'])"
