#!/bin/bash
# A/B of the L2-resident four-step kernels on one box (SDR_FFT_L2_MODE: 0 ticket/CTA items, 1 warp items, 2 TMA-fed CTA items)
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 150 python -m pytest tests/test_gpu_fft.py -m gpu -q -x --timeout 40 2>&1 | tail -5
  [ ${PIPESTATUS[0]} -eq 0 ] || { echo "TESTS FAILED - no bench"; exit 1; }
fi
for rep in 1 2; do
for v in ${MODES:-0 2}; do
  for wl in ${WLS:-c5_14 c5_15 c5_16}; do
    SDR_FFT_L2_MODE=$v timeout 40 python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MODE=$v LAG=$SDR_FFT_L2_LAG', d['config']['workload'], round(d['value'],1), round(d['roofline']['frac'],3), round(d['ms_per_step'],4), d['clocks']['sm_mhz'])" || { echo "bench failed"; exit 1; }
  done
done
done
