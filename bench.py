#!/usr/bin/env python
"""bench.py -- stream GB/s of the bit-exact dictionary scan (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--gib G] [--stream KIND] [--algo A]
                    [--scaling weak|strong] [--no-configs] [--no-cpu-baseline]

Workload (config C3 of SURVEY.md 8d / BASELINE.json configs[2]): snort.dict + et.dict merged (55,580 patterns), seeded
synthetic stream "S-planted" (uniform bytes + one planted pattern per 4096-byte block).  Default: 16 GiB PER GPU (weak
scaling; rank r owns global offsets [r*16 GiB, (r+1)*16 GiB) and reads a max_pat_len-1 halo before it, so the union of
the ranks' results equals one continuous scan); `--scaling strong` cuts ONE 16 GiB stream into N shards instead, and
the weak run also reports that figure (`strong_scaling`).  No data-path collective: the shards are independent.  A
"step" = one scan of the rank's whole shard, dense uint16 longest-match id per position (what the reference's read_char
loop produces, measure.c:292-294).

One JSON line on stdout (rank 0):
  value      device-resident throughput (CUDA events, max over ranks)
  e2e        the same metric through the reference-facing plugin call gpu_read_block (8-byte pattern ids out, host buffers,
             H2D + D2H inside the timed region), with the other buffer regimes beside it
  roofline   dominant kernel of the chosen algorithm vs the measured HBM copy peak
  sparse     scan + in-kernel match flags + compaction + NCCL gather of the sorted record lists to rank 0
  configs    C2, C4 (Karp-Rabin, rates beside results.csv:4), C5a, C5b -- each with GB/s, roofline fraction, kernel and a
             digest check against the reference's own Aho-Corasick on a 64 MiB prefix (N = 1 only)
  cpu_baseline  the reference's own AC (oracle/_ref, unmodified sources) on the host cores, plus LMAC and MPBG rows
`--impl reference` times only the reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
DATA = os.path.join(ROOT, "oracle", "_ref", "data")
DICTS = [os.path.join(DATA, "snort.dict"), os.path.join(DATA, "et.dict")]
METRIC = "stream GB/s (bit-exact matches)"
HBM_FALLBACK_GBS = 6650.0
RESULTS_CSV_MPBG = {"false_pos": 0.0, "false_neg": 0.000293, "partial": 0.022754}   # results.csv:4 (reference MPBG)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as e:  # NVML missing: report that rather than inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def adversarial_dict_bytes():
    """C5a: a^k for k = 1..256 plus every string over {a,b} of length 1..12."""
    pats = [b"a" * k for k in range(1, 257)]
    for L in range(1, 13):
        for v in range(1 << L):
            pats.append(bytes(97 + ((v >> i) & 1) for i in range(L)))
    return b"\n".join(pats) + b"\n"


def cpu_rows(ref, gen, cores, stream_kind):
    """LMAC and MPBG of the reference, timed on bounded samples: LMAC on all cores, MPBG (O(#patterns)
    per byte, mpbg.c:132-145) single-threaded on the first bytes of the reference's own stream."""
    import numpy as np
    rows = {}
    try:
        s = gen.gen(stream_kind, 0, cores * (2 << 20))
        r = ref.scan_parallel(s, cores, algo=1)
        rows["lmac"] = {"value": s.size / r["max_loop_seconds"] / 1e9, "unit": "GB/s", "cores": cores,
                        "sample": f"{s.size >> 20} MiB, lmac_read_char loop (mplmac.c:382-397), fork per core"}
        c1 = np.fromfile(os.path.join(DATA, "dictionaries_generated.stream"), dtype=np.uint8)[:1024]
        secs, _, _ = ref.scan(c1, algo=2, want_ids=False)
        rows["mpbg"] = {"value": c1.size / secs / 1e9, "unit": "GB/s", "cores": 1, "bytes_per_second": round(c1.size / secs, 1),
                        "sample": "first 1,024 bytes of dictionaries_generated.stream (C1), mpbg_read_char loop (mpbg.c:132-145)"}
    except Exception as e:
        rows["error"] = str(e)[:200]
    return rows


def reference_arm(args, rank, world):
    """The reference's own CPU Aho-Corasick (unmodified sources, oracle/_ref) on all host cores."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from reflib import Reference
    from oracle_lib import Oracle
    cores = os.cpu_count() or 1
    t0 = time.time()
    ref = Reference(DICTS, algo_mask=1)          # patterns_tree_build + ac_compile of the reference
    build_s = time.time() - t0
    gen = Oracle()                               # only to regenerate the same seeded stream on the CPU
    for p in DICTS:
        gen.add_dict_file(p)
    gen.compile()
    sample = min(args.ref_mib << 20, cores * (32 << 20))
    stream = gen.gen(args.stream, 0, sample)
    for _ in range(args.warmup_ref):
        ref.scan_parallel(stream, cores)
    times = []
    for _ in range(args.steps_ref):
        r = ref.scan_parallel(stream, cores)
        times.append(r["max_loop_seconds"])     # slowest worker's read_char loop (what measure.c:290-297 times), fork excluded
    t = sum(times) / len(times)
    gbs = sample / t / 1e9
    one = ref.scan(stream[: 32 << 20], want_ids=False)[0]
    line = {"impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps_ref, "warmup": args.warmup_ref, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"C3: snort.dict+et.dict merged ({ref.n_patterns} patterns), S-{args.stream} stream; "
                                   f"each step = {sample >> 20} MiB sample of it on the host cores",
                       "reference": "ac_read_char loop (mpac.c:304-319, measure.c:292-294), gcc -O2, fork per core with halo",
                       "build_seconds": round(build_s, 2), "single_core_MBps": round((32 << 20) / one / 1e6, 2)},
            "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "reference",
                             "sample": f"{sample >> 20} MiB of the S-{args.stream} stream, all {cores} host threads"},
            "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def emit(line):
    """The one JSON line goes to the real stdout; everything else a library prints (NCCL's version banner
    on communicator creation, ...) was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def roofline_entry(algo_name, n, step_ms, main_ms, scan_ms, n_prof, peak, peak_src):
    """The dominant kernel of the chosen algorithm, algorithmic bytes = 1 B read + 2 B written per position."""
    if algo_name == "sfx" and n_prof:
        kernel, kernel_ms, share = "sfx_scan_kernel", main_ms / n_prof, (main_ms / scan_ms) if scan_ms else None
        how = "CUDA events around the kernel on its own stream (pm_engine_set_profiling)"
    else:
        kernel = {"dfa": "dfa_hot_kernel (forward DFA, hot rows in shared memory)", "kr": "sfx_scan_kernel + kr_scan_kernel",
                  "auto": "chosen per stream (see config.auto_choice)"}.get(algo_name, algo_name)
        kernel_ms, share, how = step_ms, 1.0, "whole step (the algorithm is one kernel chain without a separable dominant one)"
    achieved = 3.0 * n / (kernel_ms * 1e-3) / 1e9
    traffic, tsrc = None, None
    tp = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if os.path.exists(tp):
        try:
            tr = json.load(open(tp))
            if tr.get("algo") == algo_name:
                traffic = tr["dram_bytes_per_stream_byte"] * n
                tsrc = f"ncu --set full capture of this command ({tr.get('source')}), dram__bytes_read+write per launch scaled by stream bytes"
        except Exception:
            traffic = None
    # SURVEY 8d's second candidate: the dependent-lookup rate.  One root2 gather per position costs the LSU data pipe
    # 3.53 wavefronts per 32 positions (the expected maximum bank load of 32 random indices; scripts/microbench/lds_gather.cu,
    # profiles/r02_kernel.md), one wavefront per clock and SM: that alone allows n_sms * clock * 32 / 3.53 bytes of stream per
    # second.  The LOWER of the two stream ceilings is the binding roofline (HBM: peak / 3).
    lookup = None
    if algo_name == "sfx":
        sm_hz = 1.965e9
        ceiling = 148 * sm_hz * 32 / 3.53 / 1e9
        lookup = {"bound": "shared-memory gather (LSU data pipe)", "stream_ceiling_GBps": ceiling, "hbm_stream_ceiling_GBps": peak / 3.0,
                  "binding": "hbm" if peak / 3.0 < ceiling else "lookup", "achieved_stream_GBps": n / (kernel_ms * 1e-3) / 1e9,
                  "frac_of_lookup_ceiling": n / (kernel_ms * 1e-3) / 1e9 / ceiling,
                  "note": "148 SMs x 1.965 GHz x 32 positions / 3.53 wavefronts per gather; ncu: l1tex data pipe 88.6 % busy with ALL of the kernel's wavefronts"}
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "lookup_roofline": lookup,
            "traffic_source": tsrc, "kernel": kernel, "algorithmic_bytes_per_stream_byte": 3, "kernel_ms": kernel_ms,
            "kernel_share_of_step": share, "kernel_timing": how, "peak_source": peak_src}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gib", type=float, default=16.0, help="stream GiB per GPU (weak) or in total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--stream", default="planted", choices=["uniform", "planted", "almost", "ab", "ascii"])
    ap.add_argument("--algo", default="sfx", choices=["sfx", "dfa", "kr", "auto"])
    ap.add_argument("--e2e-mib", type=int, default=1024, help="host-buffer bytes per e2e step")
    ap.add_argument("--ref-mib", type=int, default=256, help="upper bound of the CPU sample (MiB)")
    ap.add_argument("--steps-ref", type=int, default=3)
    ap.add_argument("--warmup-ref", type=int, default=1)
    ap.add_argument("--min-len", type=int, default=4, help="sparse mode: report matches of at least this many pattern bytes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        args.steps_ref, args.warmup_ref = max(args.steps, 1), max(args.warmup, 0)
        reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import patternmatching_b200 as pm
    from patternmatching_b200 import multi

    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the engine has no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    d = pm.Dictionary()
    for p in DICTS:
        d.add_file(p)
    d.compile()
    eng = pm.Engine(d, device=local_rank)
    algo = pm.ALGOS[args.algo]
    peak, peak_src = peaks()
    W = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    st = torch.cuda.current_stream().cuda_stream
    total = int(args.gib * (1 << 30)) // 4096 * 4096
    plan_weak = multi.plan_shards(total * world, world)            # 16 GiB per rank
    plan_strong = multi.plan_shards(total, world)                   # one 16 GiB stream cut into `world` shards
    lead = 4096                                                     # generated before the shard so that the halo is real data
    n_max = plan_weak[rank].n
    buf = torch.empty(n_max + lead, dtype=torch.uint8, device=dev)
    out = torch.empty(n_max, dtype=torch.int16, device=dev)

    def load_shard(shard):
        """Generate the shard's bytes (plus the bytes before it, the halo) into buf; returns (device ptr, n, hist_valid)."""
        if shard.lo >= lead:
            eng.generate(args.stream, shard.lo - lead, shard.n + lead, buf)
            hist = lead
        else:
            eng.generate(args.stream, shard.lo, shard.n, buf.data_ptr() + lead)
            hist = 0
        torch.cuda.synchronize()
        return buf.data_ptr() + lead, shard.n, hist

    def timed_scans(ptr, n, hist, steps, profile=False):
        for _ in range(W):
            eng.scan_device(ptr, n, out, hist_valid=hist, algo=algo, cuda_stream=st)
        barrier()
        if profile:
            eng.set_profiling(True)
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            eng.scan_device(ptr, n, out, hist_valid=hist, algo=algo, cuda_stream=st)
        ev1.record()
        barrier()
        return reduce_max(ev0.elapsed_time(ev1)) / steps

    # ---- the headline: the scaling mode asked for --------------------------------------------------------------
    shard = (plan_weak if args.scaling == "weak" else plan_strong)[rank]
    d_stream, n, hist = load_shard(shard)
    launches0 = eng.launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_step = timed_scans(d_stream, n, hist, args.steps, profile=True)
    sampler.stop_flag = True
    sampler.join()
    launches = eng.launches - launches0 - 0
    n_prof, main_ms, scan_ms = eng.read_profile()
    eng.set_profiling(False)
    launches_timed = launches * args.steps // (args.steps + W) if (args.steps + W) else launches
    bytes_all = sum(s.n for s in (plan_weak if args.scaling == "weak" else plan_strong))
    value = bytes_all / (ms_step * 1e-3) / 1e9
    auto_choice = eng.auto_choice if args.algo == "auto" else None

    # correctness summary of the result that was timed: counts + digests, reduced over ranks with NCCL
    red = multi.reduce_summary(eng.summarize(out, n, pos_base=shard.lo), dist, dev)

    # ---- sparse mode + the multi-GPU gather: scan with in-kernel match flags, bitmap compaction, then the
    # per-rank sorted record lists to rank 0 over NCCL (counts all-gather + grouped send/recv from C++) ----------------
    cap = max(n // 256, 1 << 20)
    rec = torch.empty(cap, dtype=torch.int64, device=dev)
    sparse = {"min_pattern_len": args.min_len}
    try:
        cnt = eng.scan_device_records(d_stream, n, out, rec, cap, min_len=args.min_len, hist_valid=hist, pos_base=shard.lo, algo=algo)
        barrier()
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(3):
            cnt = eng.scan_device_records(d_stream, n, out, rec, cap, min_len=args.min_len, hist_valid=hist, pos_base=shard.lo,
                                          algo=algo, cuda_stream=st)
        ev1.record()
        barrier()
        sparse_ms = reduce_max(ev0.elapsed_time(ev1)) / 3
        sparse.update({"scan_flags_compact_ms": sparse_ms, "value": bytes_all / (sparse_ms * 1e-3) / 1e9, "unit": "GB/s",
                       "records_this_rank": int(cnt), "call": "pm_engine_scan_device_records"})
        if world > 1:
            comm = pm.Comm.from_torch(dist, local_rank)
            cnts = torch.tensor([cnt], dtype=torch.int64, device=dev)
            dist.all_reduce(cnts)
            all_cap = int(cnts.item()) + 16
            allrec = torch.empty(all_cap if rank == 0 else 1, dtype=torch.int64, device=dev)
            for _ in range(2):                                                                     # warm-up (connections, channel buffers)
                comm.gather_records(rec, min(cnt, cap), allrec, all_cap, root=0, cuda_stream=st)
            barrier()
            g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(5):
                counts, tot = comm.gather_records(rec, min(cnt, cap), allrec, all_cap, root=0, cuda_stream=st)
            g1.record()
            barrier()
            g_ms = reduce_max(g0.elapsed_time(g1)) / 5
            recv_bytes = 8 * (tot - counts[0])
            sparse["gather"] = {"ms": g_ms, "records_total": int(tot), "bytes_into_rank0": int(recv_bytes),
                                "effective_GBps_into_rank0": recv_bytes / (g_ms * 1e-3) / 1e9 if g_ms else None,
                                "share_of_scan_step": g_ms / ms_step, "call": "pm_comm_gather_records (ncclAllGather of counts, then one peer-to-peer copy per rank into the root's "
                                        "IPC-mapped staging buffer; grouped ncclSend/ncclRecv when the GPUs cannot map each other)"}
            if rank == 0:
                pos = allrec[:tot] >> 24
                sparse["gather"]["position_sorted"] = bool((pos[1:] > pos[:-1]).all().item()) if tot > 1 else True
            comm.free()
            del allrec
    except Exception as ex:
        sparse["error"] = str(ex)[:200]
    del rec

    # ---- the other scaling mode, as an extra figure ----------------------------------------------------------------
    other = None
    if world > 1:
        oshard = (plan_strong if args.scaling == "weak" else plan_weak)[rank]
        o_stream, o_n, o_hist = load_shard(oshard)
        o_ms = timed_scans(o_stream, o_n, o_hist, max(args.steps, 5))
        o_total = sum(s.n for s in (plan_strong if args.scaling == "weak" else plan_weak))
        other = {"scaling": "strong" if args.scaling == "weak" else "weak", "total_bytes": o_total, "bytes_per_gpu": o_n,
                 "ms_per_step": o_ms, "value": o_total / (o_ms * 1e-3) / 1e9, "unit": "GB/s"}
        d_stream, n, hist = load_shard(shard)       # back to the headline shard for what follows
        eng.scan_device(d_stream, n, out, hist_valid=hist, algo=algo, cuda_stream=st)
        torch.cuda.synchronize()

    # ---- end to end through the reference-facing plugin call ------------------------------------------------------
    # host buffers per rank: 1 B in + 8 B ids out per position, page-locked AND page-able copies of both -- the ranks of a
    # box share its RAM, so the per-rank piece shrinks with the world size (1 GiB at N = 1, 128 MiB at N = 8)
    ne = min(max(args.e2e_mib // world, 128) << 20, n)
    host = buf[lead:lead + ne].cpu().numpy()
    hin = pm.PinnedBuffer(ne); hids = pm.PinnedBuffer(8 * ne)
    hin.array(np.uint8)[:] = host
    plug = pm.MpsGpu("sfx")                                         # gpu_create / gpu_add_pattern / gpu_compile (mps.h:71-80)
    id_of_pid = np.zeros(d.n_patterns + 1, np.uint64)
    for pid in range(1, d.n_patterns + 1):
        id_of_pid[pid] = 0x7F0000000000 + 64 * pid                 # stand-ins for the PatternsTreeNode* the reference passes
        plug.add_pattern(d.pattern(pid)[4], int(id_of_pid[pid]))
    plug.compile()
    e2e_steps = max(3, min(args.steps, 5))

    def timed_host(fn, steps):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return reduce_max((time.perf_counter() - t0) / steps)

    def rb_pinned():
        plug.reset(); plug.read_block_ptr(hin.ptr, ne, hids.ptr)

    e2e_s = timed_host(rb_pinned, e2e_steps)
    e2e_gbs = world * ne / e2e_s / 1e9
    e2e_ok = None
    if shard.lo == 0:                                              # the ids must be the device result's pids, translated
        e2e_ok = bool(np.array_equal(hids.array(np.uint64)[:ne], id_of_pid[out[:ne].cpu().numpy().view(np.uint16)]))
    regimes = {}
    pageable_ids = np.zeros(ne, np.uint64)

    def chunked(step, limit):
        def run():
            plug.reset()
            for o in range(0, limit, step):
                k = min(step, limit - o)
                plug.read_block_ptr(host.ctypes.data + o, k, pageable_ids.ctypes.data + 8 * o)
        return run

    regimes["read_block_pageable_16MiB_calls"] = world * ne / timed_host(chunked(16 << 20, ne), 2) / 1e9
    small = min(ne, 64 << 20)
    regimes["read_block_pageable_100KiB_calls"] = world * small / timed_host(chunked(100 * 1024, small), 2) / 1e9
    hu16 = pm.PinnedBuffer(2 * ne)

    def scan_pinned():
        eng.reset(); eng.scan_host_ptr(hin.ptr, ne, hu16.ptr, algo=algo)

    regimes["scan_host_u16_pinned"] = world * ne / timed_host(scan_pinned, e2e_steps) / 1e9
    pageable_u16 = np.zeros(ne, np.uint16)

    def scan_pageable_16m():
        eng.reset()
        for o in range(0, ne, 16 << 20):
            k = min(16 << 20, ne - o)
            eng.scan_host_ptr(host.ctypes.data + o, k, pageable_u16.ctypes.data + 2 * o, algo=algo)

    regimes["scan_host_u16_pageable_16MiB_calls"] = world * ne / timed_host(scan_pageable_16m, 2) / 1e9
    rcap = max(ne // 64, 1 << 20)
    hrec = pm.PinnedBuffer(8 * rcap)
    n_rec = [0]

    def scan_records():
        eng.reset()
        _, n_rec[0] = eng.scan_host_records(None, min_len=args.min_len, cap=rcap, algo=algo, src_ptr=hin.ptr, n=ne, dst_ptr=hrec.ptr)

    regimes["scan_host_records_pinned"] = world * ne / timed_host(scan_records, e2e_steps) / 1e9
    regimes = {k: round(v, 3) for k, v in regimes.items()}
    regimes["records_per_step"] = int(n_rec[0])
    regimes["note"] = ("aggregate GB/s of stream over all ranks; read_block writes 8-byte pattern ids (the reference's read_char contract, "
                       "mps.h:41-42): bound by the host threads' store rate (scripts/microbench/host_mem.cpp: 107-131 GB/s of ids = "
                       "13-16 GB/s of stream with 8-16 threads on the 16-vCPU box); scan_host_u16 returns dense uint16 pids: bound by PCIe "
                       "(2 B per position)")

    # what the host link gives a plain pinned copy while EVERY rank copies at once (the regime of the e2e numbers above)
    link = None
    try:
        mb = 256 << 20
        h_a = torch.empty(mb, dtype=torch.uint8, pin_memory=True); h_b = torch.empty(mb, dtype=torch.uint8, pin_memory=True)
        d_a = buf[:mb]; d_b = out.view(torch.uint8)[:mb]
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        h_a.copy_(d_a); torch.cuda.synchronize()   # keep the stream bytes intact: d_a gets back what it held

        def timed_copy(fn, reps=4):
            fn(); barrier()
            t = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t) / reps

        def both():
            with torch.cuda.stream(s1):
                h_b.copy_(d_b, non_blocking=True)
            with torch.cuda.stream(s2):
                d_a.copy_(h_a, non_blocking=True)

        t_d2h = timed_copy(lambda: h_b.copy_(d_b, non_blocking=True))
        t_h2d = timed_copy(lambda: d_a.copy_(h_a, non_blocking=True))
        t_both = timed_copy(both)
        mine = torch.tensor([mb / t_d2h / 1e9, mb / t_h2d / 1e9, mb / t_both / 1e9], dtype=torch.float64, device=dev)
        lo_t, hi_t = mine.clone(), mine.clone()
        if world > 1:
            dist.all_reduce(lo_t, op=dist.ReduceOp.MIN); dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
        link = {"d2h_gbs_min_max": [round(float(lo_t[0]), 1), round(float(hi_t[0]), 1)],
                "h2d_gbs_min_max": [round(float(lo_t[1]), 1), round(float(hi_t[1]), 1)],
                "both_directions_gbs_each_min_max": [round(float(lo_t[2]), 1), round(float(hi_t[2]), 1)], "bytes": mb,
                "note": f"plain pinned cudaMemcpyAsync per rank, all {world} rank(s) copying at the same time (min and max over ranks)"}
        del h_a, h_b
    except Exception as ex:  # measurement aid only
        link = {"error": str(ex)[:120]}

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": f"C3: snort.dict+et.dict merged ({d.n_patterns} patterns, max_pat_len {d.max_pat_len}), "
                               f"S-{args.stream} stream, " + (f"{args.gib:g} GiB per GPU" if args.scaling == "weak" else f"{args.gib:g} GiB in total") +
                               f", shards with {pm.HALO}-byte halo",
                   "algo": args.algo, "auto_choice": auto_choice, "result": "dense uint16 longest-match pid per position",
                   "bytes_per_gpu": n, "l2": f"input {n / (1 << 30):g} GiB per step >> 126 MB L2 (no flush needed)",
                   "timing": "CUDA events on the launching stream, max over ranks"},
        "roofline": roofline_entry(args.algo, n, ms_step, main_ms, scan_ms, n_prof, peak, peak_src),
        "e2e": {"value": e2e_gbs, "unit": "GB/s", "h2d_bytes_per_step": ne, "d2h_bytes_per_step": 2 * ne,
                "call": "gpu_read_block (MpsElem plugin surface, mps_gpu_shim.c): page-locked stream in, 8-byte pattern ids out "
                        f"({8 * ne} bytes written on the host per step by {eng.host_threads} host threads from the 2-byte pids "
                        f"that cross PCIe), one call per {ne >> 20} MiB",
                "steps": e2e_steps, "matches_device_result": e2e_ok, "host_threads": eng.host_threads, "regimes": regimes,
                "host_link": link},
        "sparse": sparse,
        "gpu_launches": int(launches_timed),
        "clocks": sampler.result(),
        "result_check": {"positions_with_match": red["positions"], "matches_with_ancestors": red["matches"],
                         "digest_sum_longest": "%016x" % red["hsum_longest"], "digest_sum_all": "%016x" % red["hsum_all"],
                         "note": "sums over all ranks; equal to one continuous scan of the whole stream"},
    }
    if other:
        line["strong_scaling" if other["scaling"] == "strong" else "weak_scaling"] = other
    hin.free(); hids.free(); hu16.free(); hrec.free(); plug.free()
    del pageable_ids, pageable_u16

    ref = None
    if not args.no_cpu_baseline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from reflib import Reference
            from oracle_lib import Oracle
            cores = os.cpu_count() or 1
            sample = min(args.ref_mib << 20, cores * (32 << 20), n)
            host_s = buf[lead:lead + sample].cpu().numpy()
            with_rows = world == 1
            ref = Reference(DICTS, algo_mask=7 if with_rows else 1)
            r = ref.scan_parallel(host_s, cores)
            line["cpu_baseline"] = {"value": sample / r["max_loop_seconds"] / 1e9, "unit": "GB/s", "cores": cores,
                                    "kind": "reference",
                                    "sample": f"first {sample >> 20} MiB of this rank's stream, reference ac_read_char loop "
                                              f"(gcc -O2), fork per core with halo; time = slowest worker's loop"}
            # the GPU result on the same sample must carry the reference's digest
            sg = eng.summarize(out, sample, pos_base=shard.lo)
            line["cpu_baseline"]["gpu_matches_reference"] = bool(
                sg["positions"] == r["positions"] and sg["matches"] == r["matches"] and
                sg["hsum_longest"] == r["hsum_longest"] and sg["hsum_all"] == r["hsum_all"]) if shard.lo == 0 else None
            if with_rows:
                gen = Oracle()
                for p in DICTS:
                    gen.add_dict_file(p)
                gen.compile()
                line["cpu_baseline"]["other_algorithms"] = cpu_rows(ref, gen, cores, args.stream)
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}

    if world == 1 and not args.no_configs:
        try:
            line["configs"] = config_block(pm, torch, np, eng, d, buf, out, lead, n, ref, peak, args)
        except Exception as e:
            line["configs"] = {"error": str(e)[:300]}
    emit(line)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def config_block(pm, torch, np, eng, d, buf, out, lead, n16, ref, peak, args):
    """The other BASELINE.json configs on one GPU: GB/s (CUDA events, 3 scans after a warm-up), algorithmic roofline
    fraction (3 B per position), the kernel that ran, and a digest check of the first 64 MiB of the result against the
    reference's own Aho-Corasick on the same bytes (oracle/_ref; C5a's dictionary differs, so its reference runs in a
    child process -- the reference is not re-entrant)."""
    dev = out.device
    cores = os.cpu_count() or 1
    chk = 64 << 20
    res = {}

    def run(engine, kind, n, algo, hist=0, ptr=None, o=None):
        o = out if o is None else o
        ptr = buf.data_ptr() + lead if ptr is None else ptr
        engine.scan_device(ptr, n, o, hist_valid=hist, algo=algo); torch.cuda.synchronize()
        ms = engine.time_scan(ptr, n, o, hist_valid=hist, algo=algo, iters=3)
        return ms

    def entry(ms, n, kernel, extra=None):
        e = {"bytes": n, "ms": ms, "value": n / ms / 1e6, "unit": "GB/s", "roofline_frac": 3.0 * n / (ms * 1e-3) / 1e9 / peak,
             "kernel": kernel}
        e.update(extra or {})
        return e

    def digest(engine, n):
        s = engine.summarize(out, n)
        return {k: s[k] for k in ("positions", "matches", "hsum_longest", "hsum_all")}

    def same(a, b):
        return all(int(a[k]) == int(b[k]) for k in ("positions", "matches", "hsum_longest", "hsum_all"))

    # C1: snort.dict on the reference's own 10 KB stream, through the plugin call (BASELINE configs[0]); the results must carry
    # the digests of the reference's AC (tests/golden/ref_snort.json, generated from the unmodified reference), and the
    # PM_ALGO_MPBG mode must reproduce the reference MPBG's success counts (results.csv row of MPBG on this config)
    try:
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_snort.json")))
        s1 = np.fromfile(os.path.join(DATA, gold["stream"]), dtype=np.uint8)
        d1 = pm.Dictionary().add_file(DICTS[0]).compile()
        e1 = pm.Engine(d1, device=dev.index or 0)
        t_small = []
        for _ in range(20):
            e1.reset(); t0 = time.perf_counter(); got = e1.scan_host(s1, algo=pm.ALGO_AUTO); t_small.append(time.perf_counter() - t0)
        d_got = torch.from_numpy(got.view(np.int16).copy()).to(dev)
        sg = e1.summarize(d_got, s1.size)
        e1.reset(); mp = e1.scan_host(s1, algo=pm.ALGO_MPBG)
        c = e1.classify(torch.from_numpy(mp.view(np.int16).copy()).to(dev), d_got, s1.size)
        res["C1_snort_reference_stream"] = {
            "bytes": int(s1.size), "ms": sorted(t_small)[len(t_small) // 2] * 1e3, "value": s1.size / sorted(t_small)[len(t_small) // 2] / 1e9, "unit": "GB/s",
            "kernel": "sfx_edge_kernel (one launch: calls of <= 256 KiB are launch-latency bound)", "call": "pm_engine_scan_host, host buffers",
            "matches_reference": bool(sg["positions"] == gold["positions"] and sg["matches"] == gold["matches"] and
                                      "%016x" % sg["hsum_longest"] == gold["hsum_longest"] and "%016x" % sg["hsum_all"] == gold["hsum_all"]),
            "reference_ac_seconds_same_config_survey_8c": 0.002133,
            "mpbg_mode_counts_vs_exact": [c["success"], c["partial"], c["false_neg"], c["false_pos"]],
            "reference_mpbg_counts_vs_its_ac": gold["mpbg_vs_ac_counts"],
            "mpbg_mode_matches_reference": [c["success"], c["partial"], c["false_neg"], c["false_pos"]] == gold["mpbg_vs_ac_counts"]}
        del e1
    except Exception as ex:
        res["C1_snort_reference_stream"] = {"error": str(ex)[:200]}

    choice_name = {0: "sfx_scan_kernel (backward suffix-trie scan)", 1: "dfa_small_kernel / dfa_hot_kernel (forward DFA in shared memory)",
                   4: "deep_scan_kernel (compact goto+failure records)", -1: "undecided"}
    # C2: merged dictionary, 1 GiB planted stream (the first GiB of C3's stream: same generator, offset 0)
    n1 = min(1 << 30, n16)
    eng.generate("planted", 0, n1, buf.data_ptr() + lead); torch.cuda.synchronize()
    ms = run(eng, "planted", n1, pm.ALGO_SFX)
    c2 = entry(ms, n1, choice_name[0])
    if ref is not None:
        r = ref.scan_parallel(buf[lead:lead + chk].cpu().numpy(), cores)
        c2["matches_reference"] = same(digest(eng, chk), r)
        c2["reference_check"] = f"first {chk >> 20} MiB: positions, matches and both digest sums equal the reference AC's"
    res["C2_merged_planted_1GiB"] = c2
    # C5b: merged dictionary, 1 GiB "almost" stream (pattern-prefix soup), kernel chosen by PM_ALGO_AUTO
    eng.generate("almost", 0, n1, buf.data_ptr() + lead); torch.cuda.synchronize()
    ms = run(eng, "almost", n1, pm.ALGO_AUTO)
    c5b = entry(ms, n1, choice_name.get(eng.auto_choice, str(eng.auto_choice)), {"auto_choice": eng.auto_choice})
    if ref is not None:
        r = ref.scan_parallel(buf[lead:lead + chk].cpu().numpy(), cores)
        c5b["matches_reference"] = same(digest(eng, chk), r)
    res["C5b_merged_almost_1GiB"] = c5b
    # C4: Karp-Rabin variant on the 16 GiB planted stream, classified against the exact result like measure.c:174-190
    free, _ = torch.cuda.mem_get_info()
    n4 = n16 if free > 2 * n16 + (4 << 30) else min(n16, 1 << 30)
    eng.generate("planted", 0, n4, buf.data_ptr() + lead); torch.cuda.synchronize()
    exact = torch.empty(n4, dtype=torch.int16, device=dev)
    eng.scan_device(buf.data_ptr() + lead, n4, exact, algo=pm.ALGO_SFX)
    eng.set_kr_seed(0xF1A90003)
    ms = run(eng, "planted", n4, pm.ALGO_KR)
    c = eng.classify(out, exact, n4)
    tot = float(sum(c.values()))
    res["C4_merged_planted_kr"] = entry(ms, n4, "sfx_scan_kernel (patterns <= 8 bytes, exact) + kr_scan_kernel (fingerprints)", {
        "seed": "0xF1A90003", "counts_vs_exact": c,
        "rates": {"false_pos": c["false_pos"] / tot, "false_neg": c["false_neg"] / tot, "partial": c["partial"] / tot},
        "reference_mpbg_rates_results_csv_4": RESULTS_CSV_MPBG,
        "note": "classified on the device against the exact scan of the same bytes (pm_engine_classify = measure_success_rate); "
                "the reference's MPBG never reports a pattern of more than 8 bytes (SURVEY Q5), hence its FN / partial rates"})
    del exact
    # C5a: adversarial small-alphabet dictionary, 1 GiB {a,b} stream
    adv = adversarial_dict_bytes()
    d5 = pm.Dictionary().add_bytes(adv).compile()
    e5 = pm.Engine(d5, device=dev.index or 0)
    e5.generate("ab", 0, n1, buf.data_ptr() + lead); torch.cuda.synchronize()
    ms = run(e5, "ab", n1, pm.ALGO_AUTO)
    c5a = entry(ms, n1, choice_name.get(e5.auto_choice, str(e5.auto_choice)),
                {"auto_choice": e5.auto_choice, "patterns": d5.n_patterns, "dictionary": "a^k k=1..256 + all {a,b} strings of length 1..12"})
    try:
        s = e5.summarize(out, chk)
        path = os.path.join(ROOT, "gpurun_out", "adv_emit.dict")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        open(path, "wb").write(adv)
        r = json.loads(subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_digest.py"), "--dict", path, "--kind", "ab",
                                       "--bytes", str(chk), "--workers", str(cores)], capture_output=True, text=True, timeout=300).stdout.strip().splitlines()[-1])
        c5a["matches_reference"] = same({k: s[k] for k in ("positions", "matches", "hsum_longest", "hsum_all")}, r)
        c5a["matches_per_byte"] = r["matches"] / chk
    except Exception as ex:
        c5a["matches_reference"] = None
        c5a["reference_error"] = str(ex)[:200]
    res["C5a_adversarial_ab_1GiB"] = c5a
    return res


if __name__ == "__main__":
    main()
