"""Summarise an `ncu --set full` report (one launch per hot kernel, tools/ncu_targets.py) into a
markdown table and a JSON that bench.py reads for `roofline.traffic`.

    python tools/ncu_summary.py gpurun_out/r1j_full.ncu-rep profiles/r1j_ncu_full
"""
import csv
import io
import json
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid")]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    recs = []
    for r in data:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("esmdiff::", "")
        d = {"kernel": name}
        for metric, key in COLS:
            if metric not in ix:
                continue
            v = float(r[ix[metric]].replace(",", "") or 0)
            d[key] = v * SCALE.get(units[ix[metric]], 1.0)
        d["dram_bytes"] = d.get("dram_read", 0) + d.get("dram_write", 0)
        d["dram_gbs"] = d["dram_bytes"] / (d["time"] * 1e-6) / 1e9 if d.get("time") else None
        recs.append(d)
    with open(out + ".json", "w") as f:
        json.dump({"source": rep, "note": "one launch per kernel at the config-2 shape (B=100, T=258), "
                   "ncu --set full --clock-control none; cold caches, serialised", "kernels": recs}, f, indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"source: {rep} (ncu --set full --clock-control none, tools/ncu_targets.py 100)\n\n")
        f.write("| kernel | grid | regs | time us | DRAM read MB | DRAM write MB | DRAM GB/s | DRAM % | tensor pipe % | XU % | issue % | L2 hit % |\n")
        f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for d in recs:
            f.write(f"| `{d['kernel']}` | {d.get('grid', 0):.0f} | {d.get('regs', 0):.0f} | {d['time']:.1f} | "
                    f"{d.get('dram_read', 0) / 1e6:.1f} | {d.get('dram_write', 0) / 1e6:.1f} | {d['dram_gbs']:.0f} | "
                    f"{d.get('dram_pct', 0):.1f} | {d.get('tensor_pct', 0):.1f} | {d.get('xu_pct', 0):.1f} | "
                    f"{d.get('issue_pct', 0):.1f} | {d.get('l2_hit_pct', 0):.1f} |\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
