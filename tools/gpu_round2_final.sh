# final validation of the round: GPU tests, smoke(), the full N=1 bench line, ncu byte accounting of one step
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu_final.log
tail -6 gpurun_out/r02_pytest_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; tail -8 gpurun_out/r02_smoke_final.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -3 gpurun_out/r02_bench_final.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_final.json'))
print('bf16', round(d['value']), 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['achieved'], d['roofline']['frac'], 'traffic', d['roofline']['traffic'])
print('vs_pytorch_cuda', d.get('vs_pytorch_cuda'), 'clocks', d['clocks'])
for k,v in d['extra'].items(): print(k, {kk: vv for kk, vv in v.items() if kk in ('frames_per_s','value','ms_per_step','steps','seconds','vs_pytorch_cuda')} if isinstance(v, dict) else v)
print('cpu', d.get('cpu_baseline'))"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for prec in bf16 tf32; do
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/launches_final_$prec.csv python tools/ncu_step.py --precision $prec --steps 1 > gpurun_out/ncu_final_$prec.log 2>&1
done
