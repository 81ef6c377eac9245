mkdir -p gpurun_out
: > gpurun_out/r02_tc_timeline_j.txt
for cfg in "128 32" "64 64" "256 16" "32 128"; do
  set -- $cfg
  timeout 200 python tools/tc_timeline.py --cin $1 --hw $2 >> gpurun_out/r02_tc_timeline_j.txt 2>&1
done
cat gpurun_out/r02_tc_timeline_j.txt
: > gpurun_out/r02_tc_debug_sweep_j.txt
for c in 23 21; do
  for dbg in 0 1 2 4 6 7; do
    B200ROMP_TC_DEBUG=$dbg timeout 120 python tools/tc_probe.py --case $c 2>&1 | grep PROBE | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l[6:]); print('debug=$dbg', d['case'], 'us/op=%.1f' % d.get('us_per_op', -1))" >> gpurun_out/r02_tc_debug_sweep_j.txt
  done
done
cat gpurun_out/r02_tc_debug_sweep_j.txt
