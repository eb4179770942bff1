"""bench.py contract checks that need no GPU: the reference arm (the reference's own viterbi.c /
trellis.c from oracle/_ref on the host cores) prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line(reference):  # skipped when oracle/_ref is not built
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-seconds", "1", "--profiles", "2000"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GCUPS" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["profiles"] == 2000 and "workload" in d["config"] and d["scaling"] == "weak"


def test_bench_never_routes_the_product_through_the_oracle():
    """Only the cpu_baseline leg and --impl reference may touch oracle/ (the checker is not the product)."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    uses = [i for i, ln in enumerate(src.splitlines(), 1) if "oracle" in ln and "import" in ln]
    body = src.splitlines()
    for i in uses:
        # every import of the oracle sits inside the CPU legs: cpu_scan_sample (shared by cpu_baseline and
        # --impl reference) and cpu_small_configs (the CPU side + parity check of configs 2 and 5)
        back = [ln for ln in body[:i] if ln.startswith("def ")]
        assert back and back[-1].startswith(("def cpu_scan_sample", "def cpu_small_configs")), (i, body[i - 1])
    # and the measured legs never call those functions
    b200 = src[src.index("def run_b200"):src.index("def _measured_peaks")]
    assert "cpu_scan_sample" not in b200 and "cpu_small_configs" not in b200 and "oracle" not in b200
