"""Run bench.py's config 2 / config 5 legs alone (dev tool; DCP_TIMING=1 for the host phases)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from deciphon_b200 import synth  # noqa: E402

args = argparse.Namespace(seed=20261018, tmp="")
t0 = time.perf_counter()
out = bench.small_config_legs(args, synth.NodePool())
print(json.dumps(out, indent=1), f"\n{time.perf_counter() - t0:.2f} s in all")
