#!/bin/bash
# One `ncu --set full` capture of the dominant kernel of every BASELINE config (final code), after the same command has
# exited 0 without ncu, plus the launch list of the default bench.  Outputs under gpurun_out/ (summarise into profiles/).
mkdir -p gpurun_out
prof() {  # tag workload kernel-regex skip
  local CMD="python bench.py --workload $2 --steps 2 --warmup 3 --no-e2e --no-cpu --configs none"
  $CMD > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -o gpurun_out/prof_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "$1: $(tail -1 gpurun_out/ncu_$1.log)"
}
prof c2 c2 fft1024_warp 3
prof c1 c1 fir_umma_kernel 3
prof c3 c3 fir_umma_poly 3
prof c3chain c3chain fir_umma_c64 15
prof fir255_c64 fir255_c64 fir_umma_c64 3
prof c5_8192 c5_13 fft_reg2 3
CMD="python bench.py --steps 2 --warmup 3"
$CMD > gpurun_out/plain_default.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_default_bench.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list: $(wc -l < gpurun_out/launches_default_bench.csv) lines"
