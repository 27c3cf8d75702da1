"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table for ONE training step
(the launches between the last two stft_mel_log kernels = one forward+backward of the AV model).
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys


def ms(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    idx = [i for i, r in enumerate(rows) if "stft_mel_log" in r["Kernel Name"]]
    step = rows[idx[-2]:idx[-1]] if len(idx) >= 2 else rows
    agg = collections.defaultdict(lambda: [0, 0.0])
    mine = 0.0
    for r in step:
        if "at::" not in r["Kernel Name"] and "nccl" not in r["Kernel Name"].lower():
            mine += ms(r)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"void |<unnamed>::|at::native::|at::", "", name)[:90]
        agg[name][0] += 1
        agg[name][1] += ms(r)
    tot = sum(v[1] for v in agg.values())
    print(f"# one AV training step (B=64), ncu gpu__time_duration.sum per launch (cold cache, serialised)\n")
    print(f"launches: {len(step)}  total {tot:.2f} ms  (kernels of libavec_b200.so: {mine:.2f} ms = {100 * mine / tot:.1f} %)\n")
    print("| ms | share | launches | kernel |\n|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| {v[1]:.2f} | {100 * v[1] / tot:.1f} % | {v[0]} | `{k}` |")
    g = [r for r in step if "gemm_tc" in r["Kernel Name"]]
    ga = collections.defaultdict(lambda: [0, 0.0])
    for r in g:
        ga[r["Grid Size"]][0] += 1
        ga[r["Grid Size"]][1] += ms(r)
    print("\n## gemm_tc_kernel launches by grid\n\n| ms | launches | avg ms | grid |\n|---:|---:|---:|---|")
    for k, v in sorted(ga.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"| {v[1]:.3f} | {v[0]} | {v[1] / v[0]:.3f} | {k} |")


if __name__ == "__main__":
    main(sys.argv[1])
