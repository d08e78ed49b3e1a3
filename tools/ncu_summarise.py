"""Summarise an `ncu --set full` report (read here, on the CPU box, with `ncu -i ... --page raw --csv`).

    python tools/ncu_summarise.py gpurun_out/r02_prof_march_b64.ncu-rep --frames 64 --volume-size 64 \
        --out profiles/r02_ncu_march_b64

Writes <out>.txt (one block of key metrics per captured launch) and <out>_ncu_traffic.json, the file bench.py reads
the `roofline.traffic` figure from (dram__bytes_read.sum + dram__bytes_write.sum per launch; the entry is used only
when its frames_per_launch and volume_size match the bench run).
"""
import argparse
import csv
import io
import json
import os
import subprocess

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--frames", type=int, required=True, help="frames per launch of the captured run")
    ap.add_argument("--volume-size", type=int, default=64)
    ap.add_argument("--out", required=True)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {os.path.basename(a.report)}: ncu --set full --clock-control none, {a.frames} frames per launch, V = {a.volume_size}",
             "# (per-launch times under ncu are cold-cache and serialised at ~1.63 GHz: compare shares and ratios, not absolutes)"]
    if a.note:
        lines.append("# " + a.note)
    kernels = []
    for r in data:
        name = r[ix["Kernel Name"]]
        lines.append("")
        lines.append(name)
        vals = {}
        for key, label in KEYS:
            if key in ix:
                vals[key] = (r[ix[key]], units[ix[key]])
                lines.append(f"  {label:24s} {r[ix[key]]} {units[ix[key]]}")
        rd = float(vals["dram__bytes_read.sum"][0]) * SCALE.get(vals["dram__bytes_read.sum"][1], 1.0)
        wr = float(vals["dram__bytes_write.sum"][0]) * SCALE.get(vals["dram__bytes_write.sum"][1], 1.0)
        kernels.append({"kernel": name.replace("void ", "").replace("sceneego::", ""), "frames_per_launch": a.frames,
                        "volume_size": a.volume_size, "dram_bytes_read": rd, "dram_bytes_write": wr,
                        "duration_ms_under_ncu": float(vals["gpu__time_duration.sum"][0]) * (1e-3 if vals["gpu__time_duration.sum"][1] == "us" else 1.0),
                        "tensor_pipe_active_pct": float(vals["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0]),
                        "l2_hit_rate_pct": float(vals["lts__t_sector_hit_rate.pct"][0])})
    with open(a.out + ".txt", "w") as f:
        f.write("\n".join(lines) + "\n")
    json.dump({"report": os.path.basename(a.report), "kernels": kernels}, open(a.out + "_ncu_traffic.json", "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
