mkdir -p gpurun_out
for pc in 64 128; do
B200ROMP_S2_KSPLIT_C=$pc timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_o_$pc.json 2> gpurun_out/r02_bench_o.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_o_$pc.json')); print($pc, round(d['value']), round(d['e2e']['value']), d['roofline']['achieved'])"
done
B200ROMP_S2_KSPLIT_C=64 timeout 300 python tools/op_profile.py --precision bf16 2> gpurun_out/op_o.err | grep "s2   64->64 .*t13\|s2  128->64" | cut -c1-200
