# ncu --set full of the pixel-pair folded 32->32 conv and of the fuse-sum kernel, in situ (one whole-path step)
mkdir -p gpurun_out
cap() {  # tag regex skip count
  timeout 500 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on --profile-from-start off \
    --kernel-name-base demangled -k "regex:$2" -s "$3" -c "$4" -f -o "gpurun_out/prof_$1" \
    python tools/ncu_step.py --steps 1 > "gpurun_out/ncu_$1.log" 2>&1
  python tools/ncu_summary.py gpurun_out/prof_$1.ncu-rep tc_wavefronts > gpurun_out/r02_ncu_$1.txt 2>&1
  head -50 gpurun_out/r02_ncu_$1.txt
}
cap fold_c32 'conv_tc2_kernel<.int.64, .int.64, .int.2' 5 2
cap fuse_sum 'fuse_sum_pipe_kernel' 0 2
