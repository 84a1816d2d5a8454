"""Per-kernel / per-family time shares from an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file ...).
    python tools/launch_shares.py gpurun_out/r02e_launches.csv > profiles/r02e_launch_shares.txt"""
import csv, re, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
per, fam = defaultdict(lambda: [0, 0.0]), defaultdict(float)
for r in rows[1:]:
    try:
        us = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    except ValueError:
        continue
    name = re.sub(r"\(.*$", "", r[ki]).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").strip()
    per[name][0] += 1; per[name][1] += us
    fam[re.sub(r"<.*$", "", name).replace("void ", "")] += us
tot = sum(v[1] for v in per.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference`;")
print("# per-launch times are cold-cache and serialised: compare SHARES")
print(f"# total kernel time in capture: {tot / 1e3:.3f} ms over {sum(v[0] for v in per.values())} launches")
print("# by kernel family: " + ", ".join(f"{k} {v / tot:.3f}" for k, v in sorted(fam.items(), key=lambda kv: -kv[1])[:12]))
for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:90s} n={n:4d} total={us:10.1f} us share={us / tot:.3f}")
