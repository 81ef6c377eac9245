mkdir -p gpurun_out
M=gpu__time_duration.sum
: > gpurun_out/r02_smpl_blend_knockout.txt
for dbg in 0 1 3 4 5 7; do
B200ROMP_SMPL_DEBUG=$dbg timeout 200 ncu --metrics $M --clock-control none --csv -k regex:smpl_blend --log-file gpurun_out/kn_$dbg.csv python bench.py --workload smpl --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<PY >> gpurun_out/r02_smpl_blend_knockout.txt
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/kn_$dbg.csv') if l.startswith('"'))]
h=rows[0]; v=[float(r[h.index('Metric Value')].replace(',','')) for r in rows[1:]]
print('debug=$dbg blend kernel us:', [round(x/1000,1) if x>1e5 else x for x in v])
PY
done
cat gpurun_out/r02_smpl_blend_knockout.txt
