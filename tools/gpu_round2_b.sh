mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu_b.log
tail -4 gpurun_out/r02_pytest_gpu_b.log
for mode in lanes nolanes; do
  if [ $mode = nolanes ]; then export B200ROMP_NO_LANES=1; else unset B200ROMP_NO_LANES; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_b_$mode.json 2> gpurun_out/r02_bench_b_$mode.err
  python -c "import json;d=json.load(open('gpurun_out/r02_bench_b_$mode.json'));print('$mode bf16', d['value'], d['e2e']['value'], d['roofline']['achieved'])"
done
unset B200ROMP_NO_LANES
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --precision tf32 > gpurun_out/r02_bench_b_tf32.json 2> gpurun_out/r02_bench_b_tf32.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_b_tf32.json'));print('tf32', d['value'], d['e2e']['value'], d['roofline']['achieved'])"
timeout 300 python tools/op_profile.py --precision tf32 > gpurun_out/r02_op_profile_tf32_b.md 2> gpurun_out/op_tf32.err
head -20 gpurun_out/r02_op_profile_tf32_b.md
