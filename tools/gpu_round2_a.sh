set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu_a.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
timeout 300 python tools/op_profile.py --precision tf32 > gpurun_out/r02_op_profile_tf32_a.md 2> gpurun_out/op_tf32.err
timeout 300 python tools/op_profile.py --precision bf16 > gpurun_out/r02_op_profile_bf16_a.md 2> gpurun_out/op_bf16.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/launches_bf16.csv python tools/ncu_step.py --precision bf16 --steps 1 > gpurun_out/ncu_bf16.log 2>&1
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/launches_tf32.csv python tools/ncu_step.py --precision tf32 --steps 1 > gpurun_out/ncu_tf32.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu_a.log; cat gpurun_out/r02_bench_a.json; tail -5 gpurun_out/r02_bench_a.err
