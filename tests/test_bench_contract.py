"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU and prints the JSON line the driver reads.
The arm times the CPU oracle port of the path (the one other place bench.py may execute oracle/), so this also checks that
the line names the same workload / metric / unit as the GPU arm's config table."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in line, k
    assert line["unit"] == "Gsamples/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["steps"] == 1 and line["n_gpus"] == 1 and line["gpu_launches"] == 0
    assert line["config"]["workload"] == "c2_fft1024_u8iq_2p28"          # BASELINE.json configs[1], the headline
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["unit"] == line["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0

    # the GPU arm describes the same workload with the same config object
    sys.path.insert(0, ROOT)
    import bench
    assert bench.config_dict("c2") == line["config"]
