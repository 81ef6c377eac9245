# final-state checks: 2-GPU sharded bench (collective inside the timed region), reference arm line, N=1 full bench line
mkdir -p gpurun_out
N=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/nccl_gather_check.py > gpurun_out/nccl_gather_check_r_$N.log 2>&1
tail -3 gpurun_out/nccl_gather_check_r_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_r_${N}gpu.json 2> gpurun_out/r02_bench_r_${N}gpu.err
tail -3 gpurun_out/r02_bench_r_${N}gpu.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_r_${N}gpu.json'));print('N=$N', d['value'], d['e2e'], d['config'].get('collective'))"
