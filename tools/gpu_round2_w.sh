# final validation of the round: GPU tests, smoke(), the full N=1 bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu_final.log
tail -4 gpurun_out/r02_pytest_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; tail -3 gpurun_out/r02_smoke_final.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -3 gpurun_out/r02_bench_final.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_final.json'))
print('bf16', round(d['value']), 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['achieved'], d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'launches', d.get('gpu_launches'))
print('vs_pytorch_cuda', d.get('vs_pytorch_cuda'), 'clocks', d['clocks'])
for k,v in d['extra'].items(): print(k, {kk: vv for kk, vv in v.items() if kk in ('frames_per_s','value','ms_per_step','steps','seconds','vs_pytorch_cuda')} if isinstance(v, dict) else v)
print('cpu', d.get('cpu_baseline'))"
