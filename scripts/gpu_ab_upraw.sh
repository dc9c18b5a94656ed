#!/bin/bash
# A/B of the polyphase FIR's raw-tile staging (SDR_UP_RAW = 0 cp.async per lane, 1 one bulk copy, 2 bulk per skew group)
mkdir -p gpurun_out
: > gpurun_out/ab_upraw.txt
for rm in ${MODES:-2 1 0}; do
  export SDR_UP_RAW=$rm
  echo "== SDR_UP_RAW=$rm" >> gpurun_out/ab_upraw.txt
  if [ "$rm" != "0" ]; then
    timeout 900 python -m pytest tests/test_gpu_fir.py -m gpu -q -x --timeout 300 -k "decimat or streaming or every_row or alignment or c3 or sweep or multichannel" > gpurun_out/pytest_upraw_$rm.log 2>&1
    echo "pytest exit $?" >> gpurun_out/ab_upraw.txt; tail -3 gpurun_out/pytest_upraw_$rm.log >> gpurun_out/ab_upraw.txt
  fi
  for wl in ${WLS:-c3 c3_2p28 c3chain}; do
    for rep in 1 2; do
      timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu 2>> gpurun_out/ab_upraw.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-36s %9.1f GS/s %8.4f ms frac %.3f sm %s' % (d['config']['workload'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz']))" >> gpurun_out/ab_upraw.txt
    done
  done
done
cat gpurun_out/ab_upraw.txt
