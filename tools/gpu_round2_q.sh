# ncu byte accounting of one whole-path step for the current graph (bf16, tf32) + full GPU test suite + smoke
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for prec in bf16 tf32; do
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/launches_q_$prec.csv python tools/ncu_step.py --precision $prec --steps 1 > gpurun_out/ncu_q_$prec.log 2>&1
done
timeout 1200 python -m pytest tests -q -x -m gpu 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu_q.log
tail -6 gpurun_out/r02_pytest_gpu_q.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_q.log 2>&1; tail -8 gpurun_out/r02_smoke_q.log
