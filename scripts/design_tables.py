"""Print the markdown tables of DESIGN.md section 6 from the committed bench lines (profiles/bench_r02_*.json)."""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = {}
for n in (1, 2, 4, 8):
    p = os.path.join(ROOT, "profiles", f"bench_r02_n{n}.json")
    if os.path.exists(p):
        L[n] = json.loads(open(p).read().strip().splitlines()[-1])
ref = json.loads(open(os.path.join(ROOT, "profiles", "bench_r02_reference.json")).read().strip().splitlines()[-1])
print("| N GPUs | stream GB/s, weak (16 GiB per GPU) | ms/step | stream GB/s, strong (one 16 GiB stream) | e2e GB/s (`gpu_read_block`, 8-byte ids) | sparse mode: scan + flags + compaction GB/s | gather to rank 0 | host threads per rank |")
print("|---|---|---|---|---|---|---|---|")
for n, l in L.items():
    st = l.get("strong_scaling", {})
    g = l["sparse"].get("gather")
    gs = f"{g['bytes_into_rank0'] / 1e6:.0f} MB in {g['ms']:.2f} ms = {g['effective_GBps_into_rank0']:.0f} GB/s, {100 * g['share_of_scan_step']:.1f} % of the step" if g else "-"
    print(f"| {n} | **{l['value']:,.0f}** | {l['ms_per_step']:.2f} | {st.get('value', l['value'] if n == 1 else 0):,.0f} | {l['e2e']['value']:.1f} | {l['sparse']['value']:,.0f} | {gs} | {l['e2e']['host_threads']} |")
l = L[1]; r = l["roofline"]
print()
print(f"N = 1 roofline: dominant kernel `{r['kernel']}` {r['kernel_ms']:.2f} ms = {r['achieved'] / 1e3:.2f} TB/s algorithmic = **{r['frac']:.3f}** of the measured "
      f"{r['peak'] / 1e3:.2f} TB/s HBM copy peak; kernel share of the step {r['kernel_share_of_step']:.3f}; DRAM traffic {r['traffic'] / 1e9 if r['traffic'] else float('nan'):.1f} GB per launch "
      f"({(r['traffic'] or 0) / l['config']['bytes_per_gpu']:.2f} B per stream byte against 3 algorithmic).  Clocks {l['clocks']}.")
print(f"CPU reference in the same run: {l['cpu_baseline']['value']:.2f} GB/s on {l['cpu_baseline']['cores']} cores; `--impl reference` arm: {ref['value']:.2f} GB/s "
      f"({ref['config']['single_core_MBps']} MB/s on one core); LMAC {l['cpu_baseline']['other_algorithms']['lmac']['value'] * 1e3:.1f} MB/s (16 cores), MPBG {l['cpu_baseline']['other_algorithms']['mpbg']['bytes_per_second']} B/s (1 core).")
print()
print("| Config | bytes | GB/s | 3 B/B roofline fraction | kernel | equals the reference |")
print("|---|---|---|---|---|---|")
for k, v in l["configs"].items():
    print(f"| {k} | {v['bytes']:,} | {v['value']:.1f} | {v.get('roofline_frac', float('nan')):.3f} | {v['kernel']} | {v.get('matches_reference')} |")
c4 = l["configs"]["C4_merged_planted_kr"]
print()
print("C4 rates:", c4["counts_vs_exact"], c4["rates"])
print("e2e regimes N=1:", {k: v for k, v in l["e2e"]["regimes"].items() if k != "note"})
print("host link:", l["e2e"]["host_link"])
