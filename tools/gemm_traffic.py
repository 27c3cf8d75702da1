"""Sum dram__bytes_read.sum + dram__bytes_write.sum over the tcgen05 GEMM / conv launches of ONE training step from an
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` pass over tools/one_step.py.
usage: python tools/gemm_traffic.py gpurun_out/r02_gemm_traffic.csv > profiles/r02_gemm_traffic.json"""
import csv
import json
import sys

with open(sys.argv[1]) as f:
    rows = list(csv.DictReader([ln for ln in f if not ln.startswith("==")]))
ids = [r["ID"] for r in rows if "stft_mel_log" in r["Kernel Name"]]
uniq = sorted({int(i) for i in ids})
lo, hi = (uniq[-2], uniq[-1]) if len(uniq) >= 2 else (0, 1 << 60)
names = ("gemm_tc_kernel", "conv3x3_halo64", "wgrad_halo64", "wgrad_img", "stem3d_")
tot, n, per = 0.0, set(), {}
for r in rows:
    if not (lo <= int(r["ID"]) < hi) or not any(k in r["Kernel Name"] for k in names):
        continue
    v = float(r["Metric Value"].replace(",", ""))
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
    tot += v * mult
    n.add(r["ID"])
    key = next(k for k in names if k in r["Kernel Name"])
    per[key] = per.get(key, 0.0) + v * mult
print(json.dumps({"dram_bytes_per_step": tot, "launches": len(n), "per_kernel": per,
                  "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over tools/one_step.py (AV, B = 64, training graph), "
                         "launches between the last two stft_mel_log kernels"}, indent=1))
