#!/usr/bin/env python
"""One kernel's `ncu --set full` capture as a tracked markdown summary.
    python scripts/ncu_kernel_md.py <prof.ncu-rep> <profiles/out.md> <stream bytes of the profiled launch> "<title>" ["note" ...]
Writes the headline metrics, per-stream-byte figures (L1 wavefronts, L2 sectors, DRAM bytes, warp instructions), the
pipe utilisation and the top stall reasons."""
import csv, re, subprocess, sys

rep, out, nbytes, title = sys.argv[1], sys.argv[2], int(float(sys.argv[3])), sys.argv[4]
notes = sys.argv[5:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u, v = rr[0], rr[1], rr[2]


def val(name):
    if name not in h:
        return None
    try:
        return float(v[h.index(name)].replace(",", ""))
    except ValueError:
        return None


def byt(name):
    x = val(name)
    if x is None:
        return None
    return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u[h.index(name)].lower(), 1)


ms = val("gpu__time_duration.sum") * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}[u[h.index("gpu__time_duration.sum")]]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
with open(out, "w") as f:
    f.write(f"# {title}\n\n`ncu --set full --clock-control none` ({rep.split('/')[-1]}); kernel `{v[h.index('Kernel Name')]}`; "
            f"profiled launch: {nbytes} stream bytes, {ms:.3f} ms = {nbytes / ms / 1e6:.1f} GB/s under the profiler "
            f"(bench values are never taken under ncu).\n\n")
    for n in notes:
        f.write(n + "\n\n")
    f.write("| metric | value | unit |\n|---|---|---|\n")
    for k in keys:
        if k in h:
            f.write(f"| {k} | {v[h.index(k)]} | {u[h.index(k)]} |\n")
    f.write("\nPer stream byte:\n\n| quantity | per byte |\n|---|---|\n")
    per = [("warp instructions", val("smsp__inst_executed.sum")), ("L1 LSU data-pipe wavefronts", val("l1tex__data_pipe_lsu_wavefronts.sum")),
           ("L2 sectors (32 B)", val("lts__t_sectors.sum")), ("DRAM bytes read", byt("dram__bytes_read.sum")),
           ("DRAM bytes written", byt("dram__bytes_write.sum"))]
    for name, x in per:
        if x is not None:
            f.write(f"| {name} | {x / nbytes:.4f} |\n")
    stalls = sorted(((float(v[i]), n) for i, n in enumerate(h) if re.search(r"average_warps_issue_stalled.*_per_issue_active", n) and v[i]), reverse=True)
    f.write("\nTop warp stall reasons (warps stalled per issue-active cycle):\n\n")
    for x, n in stalls[:8]:
        f.write(f"- {n.split('stalled_')[1].split('_per_')[0]}: {x:.2f}\n")
print(open(out).read())
