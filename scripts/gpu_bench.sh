#!/bin/bash
# default bench (headline + configs array) and the reference arm.  Outputs under gpurun_out/.
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 3 "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2> gpurun_out/bench.time; echo "bench exit $?"
tail -3 gpurun_out/bench.time; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print("C2 %.1f GS/s frac %.3f e2e %s" % (d['value'], d['roofline']['frac'], json.dumps(d['e2e'])[:600]))
    print("cpu", json.dumps(d['cpu_baseline']))
    print("clocks", d['clocks'])
    for c in d['configs']:
        if 'error' in c: print(c); continue
        print("%-40s %9.2f GS/s  %8.3f ms  frac %.3f  parity %s  setup %ss" % (c['workload'], c['value'], c['ms_per_step'], c['roofline']['frac'], c['parity_bit_exact'], c['setup_s']))
except Exception as e:
    print("parse failed", e)
PY
