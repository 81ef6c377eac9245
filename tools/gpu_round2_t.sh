# ncu --set full of the two SMPL tensor-core kernels at the cfg5 size (65,536 persons)
mkdir -p gpurun_out
cap() {  # tag regex
  timeout 500 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on \
    --kernel-name-base demangled -k "regex:$2" -s 2 -c 1 -f -o "gpurun_out/prof_$1" \
    python bench.py --workload smpl --steps 2 --warmup 1 --no-cpu-baseline > "gpurun_out/ncu_$1.log" 2>&1
  python tools/ncu_summary.py gpurun_out/prof_$1.ncu-rep tc_wavefronts > gpurun_out/r02_ncu_$1.txt 2>&1
  grep -v "^smsp__average_warps_issue_stalled_\(barrier\|membar\|sleeping\|branch\|no_inst\|math\|tex\|lg_\|mio\|dispatch\|not_sel\|selected\)" gpurun_out/r02_ncu_$1.txt | head -40
}
cap smpl_blend 'smpl_blend_tc_kernel'
cap smpl_skin 'smpl_skin_tc_kernel'
