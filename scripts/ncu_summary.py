#!/usr/bin/env python
"""Summarise ncu outputs into profiles/ (tracked).  Usage:
    python scripts/ncu_summary.py <tag> <launches.csv> <prof.ncu-rep> [stream_bytes_of_the_profiled_launch]
Writes profiles/<tag>_launches.md, profiles/<tag>_kernel.md and profiles/traffic_<round>.json."""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
nbytes = int(sys.argv[4]) if len(sys.argv) > 4 else None
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

rows = list(csv.reader(open(launches)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 2:]:
    if len(r) > vi:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")) / 1e3)
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w") as f:
    f.write(f"# ncu launch list ({tag})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over the bench command; "
            "per-launch times are cold-cache and serialised: compare SHARES.\n\n| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|\n")
    for k, v in agg.items():
        f.write(f"| `{k[:90]}` | {len(v)} | {sum(v)/len(v):.1f} | {sum(v):.1f} | {100*sum(v)/tot:.1f}% |\n")
    # the timed step of bench.py = the device-resident scans: a scan-kernel launch of >= 1 ms and the deep / edge
    # launches that follow it (the host-path chunks and the auto-mode samples are the short launches)
    steps, cur = [], None
    for r in rows[hi + 2:]:
        if len(r) <= vi:
            continue
        name, us = r[ki], float(r[vi].replace(",", "")) / 1e3
        if "sfx_scan_kernel" in name:
            cur = {"scan": us, "deep": 0.0, "edge": 0.0} if us >= 1000 else None
            if cur:
                steps.append(cur)
        elif cur and "sfx_deep_kernel" in name:
            cur["deep"] += us
        elif cur and "sfx_edge_kernel" in name:
            cur["edge"] += us
        elif "sfx_" not in name:
            cur = None
    if steps:
        sc = sum(x["scan"] for x in steps); dp = sum(x["deep"] for x in steps); ed = sum(x["edge"] for x in steps)
        f.write(f"\nDevice-resident steps ({len(steps)} scans of the full shard): scan kernel {sc/len(steps):.1f} us, deep kernel "
                f"{dp/len(steps):.1f} us, edge kernel {ed/len(steps):.1f} us per step; **scan kernel share of the step "
                f"{sc/(sc+dp+ed):.3f}** (bench.py reports `roofline.kernel_share_of_step` from CUDA events).\n")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u, v = rr[0], rr[1], rr[2]
get = lambda name: v[h.index(name)] if name in h else None
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "smsp__warps_eligible.avg.per_cycle_active"]
stalls = sorted(((float(v[i]), n) for i, n in enumerate(h) if re.search(r"average_warps_issue_stalled.*_per_issue_active", n) and v[i]), reverse=True)
with open(os.path.join(ROOT, "profiles", f"{tag}_kernel.md"), "w") as f:
    f.write(f"# ncu --set full, dominant kernel ({tag})\n\nKernel: `{get('Kernel Name')}`\n\n| metric | value | unit |\n|---|---|---|\n")
    for k in keys:
        if k in h:
            f.write(f"| {k} | {v[h.index(k)]} | {u[h.index(k)]} |\n")
    f.write("\nTop warp stall reasons (warps stalled per issue-active cycle):\n\n")
    for val, n in stalls[:8]:
        f.write(f"- {n.split('stalled_')[1].split('_per_')[0]}: {val:.2f}\n")

def to_bytes(name):
    x, unit = float(get(name)), u[h.index(name)].lower()
    return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[unit]
if nbytes:
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    rnd = re.search(r"r(\d+)", tag).group(1)
    json.dump({"algo": "sfx", "kernel": get("Kernel Name"), "stream_bytes": nbytes, "dram_bytes_read": rd, "dram_bytes_write": wr,
               "dram_bytes_per_stream_byte": (rd + wr) / nbytes, "algorithmic_bytes_per_stream_byte": 3,
               "source": os.path.basename(rep)}, open(os.path.join(ROOT, "profiles", f"traffic_r{rnd}.json"), "w"), indent=1)
print(open(os.path.join(ROOT, "profiles", f"{tag}_launches.md")).read())
print(open(os.path.join(ROOT, "profiles", f"{tag}_kernel.md")).read())
