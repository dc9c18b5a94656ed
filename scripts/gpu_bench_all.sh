#!/bin/bash
# tests + bench sweep over the BASELINE configs + ncu launch list and one full capture. Outputs under gpurun_out/.
mkdir -p gpurun_out
R=${ROUND_TAG:-r01}
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
: > gpurun_out/bench_all.jsonl
for wl in c2 c1 c1c fir255_u8 c3 c3chain c3chain_fastest c3chain_linear c4 c4f64 c4_1024 fm c5_8 c5_9 c5_10 c5_11 c5_12 c5_13 c5_14 c5_15 c5_16; do
  extra="--no-e2e --no-cpu"
  if [ "$wl" = "c2" ]; then extra=""; fi
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 $extra >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_all.err
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_all.jsonl'):
    d=json.loads(l); r=d['roofline']
    print("%-28s %9.1f GS/s  %7.3f ms  %6.1f GB/s  frac %.3f  e2e %s" % (d['config']['workload'], d['value'], d['ms_per_step'], r['achieved'], r['frac'], d['e2e']['value'] if d.get('e2e') else None))
PY
# ncu: launch list of the default bench, then one full capture of the FFT kernel
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_${R}.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fft1024_warp -s 3 -c 2 -o gpurun_out/prof_fft_${R} -f $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -20
