#!/bin/bash
# in-situ `ncu --set full` captures of selected conv kernels inside one bench step (graphs off so that every conv is
# a plain launch).  usage: tools/ncu_kernels.sh  -> gpurun_out/prof_<tag>.ncu-rep
mkdir -p gpurun_out
run() {  # tag regex skip
  B200ROMP_NO_GRAPH=1 timeout 500 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on \
    --kernel-name-base demangled -k "regex:$2" -s "$3" -c 1 -f -o "gpurun_out/prof_$1" \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > "gpurun_out/ncu_$1.log" 2>&1
}
run c256 'conv_tc_kernel<.int.3, .int.256' 30
run c32 'conv_tc_kernel<.int.3, .int.32' 80
run c64 'conv_tc_kernel<.int.3, .int.64' 100
ls -la gpurun_out/*.ncu-rep
