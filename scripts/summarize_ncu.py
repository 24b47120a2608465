"""Turns gpurun_out/*.ncu-rep and the launch list into the tracked summaries under profiles/.
Usage: python scripts/summarize_ncu.py <tag> <launches.csv> <rep> [<rep> ...]"""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    tag, launches = sys.argv[1], sys.argv[2]
    reps = sys.argv[3:]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    # launch list -> per-kernel totals and shares
    tot = defaultdict(lambda: [0, 0.0])
    with open(launches) as f:
        rows = [r for r in csv.reader(f) if len(r) > 14 and r[0].isdigit()]
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        tot[name][0] += 1
        tot[name][1] += float(r[-1]) / 1e3
    allus = sum(v[1] for v in tot.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {len(rows)} launches of `python bench.py --steps 2 --warmup 1`\n")
        f.write("# (serialised, cold-cache: compare shares, not absolutes)\n")
        f.write(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
        for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:60]:60s} {n:8d} {us:12.1f} {us / n:10.2f} {100 * us / allus:6.1f}%\n")
    traffic = {}
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full.txt"), "w") as f:
        for rep in reps:
            hdr, units, data = raw(rep)
            f.write(f"# ncu --set full --clock-control none --import-source on  ({os.path.basename(rep)})\n")
            stall = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
            for r in data:
                f.write(f"\n== {r[hdr.index('Kernel Name')][:90]}\n")
                for w in WANT:
                    if w in hdr:
                        f.write(f"   {w:70s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}\n")
                import re
                kname = re.sub(r"<[^<>]*>$", "", r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")).split("::")[-1]
                if "dram__bytes_read.sum" in hdr:
                    def to_bytes(col):
                        v, u = float(r[hdr.index(col)]), units[hdr.index(col)].lower()
                        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
                    traffic[kname] = int(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
                st = sorted(((float(r[hdr.index(h)] or 0), h) for h in stall), reverse=True)[:6]
                f.write("   top stalls (warps per issue): " + ", ".join(
                    f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}" for v, h in st) + "\n")
    import json
    traffic["source"] = (f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full (profiles/{tag}_ncu_full.txt); "
                         "the 200 KF window stays L2-resident between the iterations of a step")
    with open(os.path.join(ROOT, "profiles", "kernel_traffic.json"), "w") as jf:
        json.dump(traffic, jf, indent=1)
    print("wrote profiles/", tag)


if __name__ == "__main__":
    main()
