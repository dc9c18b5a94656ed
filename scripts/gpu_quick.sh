#!/bin/bash
# quick iteration: selected tests + selected bench workloads (+ optional ncu of one kernel)
mkdir -p gpurun_out
T=${TESTS:-tests}
timeout 1200 python -m pytest $T -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -${TAILN:-12} gpurun_out/pytest_gpu.log
: > gpurun_out/bench_quick.jsonl
for wl in ${WLS:-c2}; do
  timeout 600 python bench.py --workload $wl --steps ${STEPS:-20} --warmup 3 ${BENCH_EXTRA:---no-e2e --no-cpu} >> gpurun_out/bench_quick.jsonl 2>> gpurun_out/bench_quick.err
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_quick.jsonl'):
    d=json.loads(l); r=d['roofline']
    print("%-28s %9.1f GS/s  %7.3f ms  %6.1f GB/s  frac %.3f  e2e %s clocks %s" % (d['config']['workload'], d['value'], d['ms_per_step'], r['achieved'], r['frac'], d['e2e']['value'] if d.get('e2e') else None, d['clocks']))
PY
tail -3 gpurun_out/bench_quick.err 2>/dev/null
if [ -n "$NCU_K" ]; then
  CMD="python bench.py --workload ${NCU_WL:-c2} --steps 2 --warmup 3 --no-e2e --no-cpu"
  $CMD > gpurun_out/plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$NCU_K -s 3 -c 2 -o gpurun_out/prof_${NCU_TAG:-k} -f $CMD > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
fi
