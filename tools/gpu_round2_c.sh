mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu_c.log
tail -12 gpurun_out/r02_pytest_gpu_c.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_c.json'));print('bf16', d['value'], d['e2e']['value'], d['roofline']['achieved'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --precision tf32 > gpurun_out/r02_bench_c_tf32.json 2> gpurun_out/r02_bench_c_tf32.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_c_tf32.json'));print('tf32', d['value'], d['e2e']['value'], d['roofline']['achieved'])"
