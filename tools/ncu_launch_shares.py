"""Per-kernel shares of a `ncu --metrics gpu__time_duration.sum --csv` launch list (read here, on the CPU box).

    python tools/ncu_launch_shares.py gpurun_out/r02_launches_final.csv "header line" > profiles/r02_ncu_launch_shares.txt
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path, header = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "").replace("sceneego::", "")
        rows.append((name, ms))
    tot = sum(ms for _, ms in rows)
    agg = OrderedDict()
    for n, ms in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ms
    if header:
        print("# " + header)
    print("# cold-cache, serialised times: compare SHARES")
    print(f"# total {tot:.2f} ms over {len(rows)} launches")
    print("kernel, launches, total ms, share")
    for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n}, {c}, {ms:.3f}, {100 * ms / tot:.1f} %")


if __name__ == "__main__":
    main()
