# pixel-pair folded 32->32 convs: parity, per-op times, bench A/B against B200ROMP_TC_NO_FOLD=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_e2e.py -q -x -m gpu 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_h.log
tail -6 gpurun_out/r02_pytest_gpu_h.log
timeout 300 python tools/op_profile.py --precision bf16 > gpurun_out/r02_op_profile_bf16_h.md 2> gpurun_out/op_h.err; grep -n "pixel-pairs" gpurun_out/r02_op_profile_bf16_h.md | head -4; tail -5 gpurun_out/r02_op_profile_bf16_h.md
timeout 400 python bench.py --steps 20 --warmup 3 --no-extra > gpurun_out/r02_bench_h_fold.json 2> gpurun_out/r02_bench_h.err
B200ROMP_TC_NO_FOLD=1 timeout 400 python bench.py --steps 20 --warmup 3 --no-extra > gpurun_out/r02_bench_h_nofold.json 2>> gpurun_out/r02_bench_h.err
python -c "
import json
for n in ('fold','nofold'):
    d=json.load(open('gpurun_out/r02_bench_h_%s.json'%n)); print(n, round(d['value']), round(d['e2e']['value']), d['roofline']['achieved'])"
