"""Aggregate an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum
per launch, --csv) by kernel family (dev tool):  python scripts/launch_summary.py file.csv [steps]"""
import collections
import csv
import re
import sys


def family(name):
    n = re.sub(r"^void ", "", name).replace("dcp::", "").replace("(anonymous namespace)::", "")
    n = n.replace("(int)", "").replace("(bool)", "")
    m = re.match(r"score_row_kernel<(\d+), (\d+), (\d+), (\d)(?:, (\d))?>", n)
    if m:
        q, seg, mode, dump = (int(x) for x in m.groups()[:4])
        stage = " staged" if m.group(5) == "1" else ""
        if dump:
            return "score_row_kernel<Q,SEG,WHOLE,DUMP> (trace: value dump)"
        if mode == 0:
            return ("score_row_kernel<Q,32,WHOLE> (128 < K <= 256)" if seg == 32 else "score_row_kernel<Q,16/8/4,WHOLE> (K <= 128)") + stage
        return {1: "score_row_kernel<8,32,FIRST> (first 256-node segment)", 2: "score_row_kernel<8,32,MID> (later full segments)",
                3: "score_row_kernel<Q,SEG,LAST> (tail segments)"}[mode] + stage
    m = re.match(r"score_reg_kernel<(\d+), (\d+), (\d)>", n)
    if m:
        return "score_reg_kernel<Q,W,DUMP> (trace: value dump, K > 256)" if m.group(3) == "1" else "score_reg_kernel<Q,W> (exact redo of failed speculation)"
    return re.sub(r"[<(].*", "", n)


def main():
    lines = open(sys.argv[1]).read().splitlines()
    hi = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    agg = collections.defaultdict(lambda: [set(), 0.0, 0.0, 0.0])
    for r in csv.DictReader(lines[hi:]):
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        a = agg[family(r["Kernel Name"])]
        a[0].add(r["ID"])
        unit, metric = r["Metric Unit"], r["Metric Name"]
        if metric.startswith("gpu__time"):
            a[1] += v / 1e6 if unit.startswith("n") else v / 1e3 if unit.startswith("u") else v if unit.startswith("m") else v * 1e3
        else:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            a[2 if "read" in metric else 3] += v * scale
    tot = sum(a[1] for a in agg.values())
    print("| kernels | launches | ms | share | DRAM read GB/step | DRAM write GB/step |\n|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {len(a[0])} | {a[1]:.1f} | {100 * a[1] / tot:.1f} % | {a[2] / steps / 1e9:.2f} | {a[3] / steps / 1e9:.2f} |")
    print(f"total {tot:.1f} ms over {steps:g} steps")


main()
