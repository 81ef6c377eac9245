# knock-out sweep of the pair kernels on the small-resolution layers (where do the 24-30 us go?)
mkdir -p gpurun_out
: > gpurun_out/r02_tc_debug_sweep_i.txt
for c in 23 24 21 22; do
  for dbg in 0 1 2 3 4 7; do
    B200ROMP_TC_DEBUG=$dbg timeout 120 python tools/tc_probe.py --case $c 2>&1 | grep PROBE | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l[6:]); print('debug=$dbg', d['case'], 'us/op=%.1f' % d.get('us_per_op', -1))" >> gpurun_out/r02_tc_debug_sweep_i.txt
  done
done
cat gpurun_out/r02_tc_debug_sweep_i.txt
