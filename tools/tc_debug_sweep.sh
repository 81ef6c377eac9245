#!/bin/bash
# decomposition experiment: time the isolated conv ops with parts of the kernel disabled (B200ROMP_TC_DEBUG bits)
for dbg in 0 1 2 3 4 5 6 7; do
  for c in 21 22; do
    B200ROMP_TC_NO_2CTA=${NO2CTA:-1} B200ROMP_TC_DEBUG=$dbg timeout 120 python tools/tc_probe.py --case $c 2>&1 | grep PROBE | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l[6:]); print('debug=$dbg', d['case'], 'us/op=%.1f' % d.get('us_per_op', -1))"
  done
done
