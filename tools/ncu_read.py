"""Read an .ncu-rep here (no GPU): key raw metrics per captured launch + the hottest SASS lines.

    python tools/ncu_read.py gpurun_out/prof.ncu-rep [kernel-index]
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_active.avg",
        "launch__grid_size", "smsp__inst_executed.sum"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print("launches:", [r[ix["Kernel Name"]][:60] for r in data])
    for w in WANT:
        if w in ix:
            print(f"{w:75s} {units[ix[w]]:12s} {[r[ix[w]] for r in data]}")
    st = sorted(((float(data[which][i].replace(",", "") or 0), h.replace("smsp__pcsamp_warps_issue_stalled_", ""))
                 for h, i in ix.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_")
                 and not h.endswith("_not_issued")), reverse=True)[:10]
    print("stall samples (kernel", which, "):", st)
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    # the source page lists every sampled launch twice (two views): collapse the pairs
    dedup = []
    for blk in blocks:
        prev = dedup[-1] if dedup else None
        if prev and prev["name"] == blk["name"] and len(prev["rows"]) == len(blk["rows"]) and not prev.get("paired"):
            prev["paired"] = True
            continue
        dedup.append(blk)
    b = dedup[which]
    ix = {h: i for i, h in enumerate(b["hdr"])}
    S = ix["# Samples"] if "# Samples" in ix else ix["Warp Stall Sampling (All Samples)"]
    stalls = [h for h in b["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[S]) for r in b["rows"])
    agg = {h: sum(int(r[ix[h]]) for r in b["rows"]) for h in stalls}
    print(f"\n{b['name'][:80]}: {tot} samples, {len(b['rows'])} SASS instructions")
    print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
    top = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][S]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]
    for i in sorted(top):
        r = b["rows"][i]
        st = sorted(((int(r[ix[h]]), h) for h in stalls if int(r[ix[h]]) > 0), reverse=True)[:2]
        print(f"{i:5d} {int(r[S]):6d}  {r[1].strip()[:90]:90s} {st}")


if __name__ == "__main__":
    main()
