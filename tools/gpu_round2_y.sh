mkdir -p gpurun_out
B200ROMP_SUM_RING=1 timeout 300 python -m pytest tests/test_gpu_fuse_sum.py -q -x -m gpu 2>&1 | tail -2
B200ROMP_SUM_RING=1 B200ROMP_NO_FUSE1X1_MERGE=1 timeout 300 python -m pytest tests/test_gpu_e2e.py -q -x -m gpu 2>&1 | tail -2
run() { env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['value']), round(d['e2e']['value']), round(d['roofline']['achieved'],1))"; }
run A=0
run B200ROMP_SUM_RING=1 B200ROMP_NO_FUSE1X1_MERGE=1
run B200ROMP_SUM_RING=1
run A=0
