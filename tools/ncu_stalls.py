"""Summarise an ncu source page: top stalled instructions of one kernel launch.
usage: python tools/ncu_stalls.py report.ncu-rep <launch-id> [topN]"""
import csv, subprocess, sys, io
rep, kid = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; idx = {n: i for i, n in enumerate(hdr)}
data = rows[h + 1:]
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
seen, uniq = set(), []
for r in data:
    if r[0] in seen: continue
    seen.add(r[0]); uniq.append(r)
tot = sum(f(r, "# Samples") for r in uniq)
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
print("total samples", tot)
for r in sorted(uniq, key=lambda r: -f(r, "# Samples"))[:top_n]:
    st = sorted([(f(r, s), s) for s in stalls], reverse=True)[:2]
    print("%7d %5.1f%% exec=%9d %-64s %s" % (f(r, "# Samples"), 100 * f(r, "# Samples") / tot, f(r, "Instructions Executed"), r[idx["Source"]].strip()[:64], [(int(a), b) for a, b in st]))
