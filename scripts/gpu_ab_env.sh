#!/bin/bash
# generic A/B inside ONE box: for each value of env var $VAR in $VALS ("-" = unset) run the bench workloads $WLS, $REPS times, interleaved
mkdir -p gpurun_out
out=gpurun_out/ab_env.txt
: > $out
for rep in $(seq 1 ${REPS:-2}); do
  for v in $VALS; do
    for wl in $WLS; do
      if [ "$v" = "-" ]; then unset $VAR; else export $VAR=$v; fi
      timeout 300 python bench.py --workload $wl --steps ${STEPS:-20} --warmup 3 --no-e2e --no-cpu --configs none 2>> gpurun_out/ab_env.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v %-36s %9.1f GS/s %8.4f ms frac %.3f host_ms %s sm %s' % (d['config']['workload'], d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('host_enqueue_ms_per_step'), d['clocks']['sm_mhz']))" >> $out
    done
  done
done
cat $out
