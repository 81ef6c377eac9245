# final N=1 bench line of the round (full: extras, comparator, cpu baseline) + the reference arm line
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err
tail -3 gpurun_out/r02_bench_final2.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_final2.json'))
print('bf16', round(d['value']), 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['achieved'], d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'launches', d.get('gpu_launches'))
print('vs_pytorch_cuda', d.get('vs_pytorch_cuda'), 'clocks', d['clocks'])
for k,v in d['extra'].items(): print(k, {kk: vv for kk, vv in v.items() if kk in ('frames_per_s','value','ms_per_step','steps','seconds','vs_pytorch_cuda')} if isinstance(v, dict) else v)
print('cpu', d.get('cpu_baseline'))"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_final2_reference.json 2> gpurun_out/r02_bench_final2_reference.err; cat gpurun_out/r02_bench_final2_reference.json | cut -c1-400
