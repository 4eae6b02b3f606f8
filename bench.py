#!/usr/bin/env python
"""bench.py -- stream GB/s of the bit-exact dictionary scan (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--gib G] [--stream KIND] [--algo A]

Workload (config C3 of SURVEY.md 8d / BASELINE.json configs[2]): snort.dict + et.dict merged
(55,580 patterns), seeded synthetic stream "S-planted" (uniform bytes + one planted pattern per
4096-byte block), 16 GiB PER GPU: rank r owns global offsets [r*16 GiB, (r+1)*16 GiB) and reads a
max_pat_len-1 halo before it, so the union of the ranks' results equals one continuous scan.  Weak
scaling, no data-path collective (independent shards); NCCL only reduces the per-rank match counts
and digests after the timed region.  A "step" = one scan of the rank's whole shard, dense uint16
longest-match id per position (what the reference's read_char loop produces, measure.c:292-294).

One JSON line on stdout (rank 0).  `value` = device-resident throughput, `e2e` = through the public
host-buffer call (pm_engine_scan_host, pinned buffers, H2D + D2H inside the timed region),
`roofline` = dominant kernel vs the measured HBM copy peak, `cpu_baseline` = the reference's own AC
(oracle/_ref, unmodified sources) on the host cores.  `--impl reference` times only that.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
DATA = os.path.join(ROOT, "oracle", "_ref", "data")
DICTS = [os.path.join(DATA, "snort.dict"), os.path.join(DATA, "et.dict")]
METRIC = "stream GB/s (bit-exact matches)"
HBM_FALLBACK_GBS = 6650.0


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


def reference_arm(args, rank, world):
    """The reference's own CPU Aho-Corasick (unmodified sources, oracle/_ref) on all host cores."""
    if rank != 0:
        return
    import numpy as np
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
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gib", type=float, default=16.0, help="stream GiB per GPU")
    ap.add_argument("--stream", default="planted", choices=["uniform", "planted", "almost", "ab", "ascii"])
    ap.add_argument("--algo", default="sfx", choices=["sfx", "dfa", "kr"])
    ap.add_argument("--e2e-mib", type=int, default=1024, help="host-buffer bytes per e2e step")
    ap.add_argument("--ref-mib", type=int, default=256, help="upper bound of the CPU sample (MiB)")
    ap.add_argument("--steps-ref", type=int, default=3)
    ap.add_argument("--warmup-ref", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    n = int(args.gib * (1 << 30)) // 4096 * 4096          # bytes per GPU
    lead = 4096                                            # generated before the shard so that the halo is real data
    off = rank * n                                         # global offset of this rank's shard
    have_lead = off >= lead
    buf = torch.empty(n + lead, dtype=torch.uint8, device=dev)
    out = torch.empty(n, dtype=torch.int16, device=dev)
    if have_lead:
        eng.generate(args.stream, off - lead, n + lead, buf)
    else:
        eng.generate(args.stream, off, n, buf.data_ptr() + lead)
    d_stream = buf.data_ptr() + lead
    hist = lead if have_lead else 0
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    st = torch.cuda.current_stream().cuda_stream
    for _ in range(max(args.warmup, 3)):
        eng.scan_device(d_stream, n, out, hist_valid=hist, algo=algo, cuda_stream=st)
    barrier()
    launches0 = eng.launches
    eng.set_profiling(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        eng.scan_device(d_stream, n, out, hist_valid=hist, algo=algo, cuda_stream=st)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    sampler.stop_flag = True
    sampler.join()
    launches = eng.launches - launches0
    n_prof, main_ms, scan_ms = eng.read_profile()
    eng.set_profiling(False)
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms.item()) / args.steps
    value = world * n / (ms_step * 1e-3) / 1e9

    # correctness summary of the result that was timed: counts + digests, reduced over ranks with NCCL
    s = eng.summarize(out, n, pos_base=off)
    from patternmatching_b200 import multi
    red = multi.reduce_summary(s, dist, dev)       # sums over ranks (NCCL all_reduce): the only collective

    # the multi-GPU gather of SURVEY 8(e), on a bounded slice (first 16 MiB of every rank's result): compact the
    # dense slice to position-sorted (pos << 24 | pid) records on the device, all-gather the counts, then the
    # variable-length record lists (NCCL when N > 1); rank order is position order.
    ng = min(16 << 20, n)
    cap = ng
    rec = torch.empty(cap, dtype=torch.int64, device=dev)
    cnt = eng.compact(out, ng, rec, cap, pos_base=off)
    torch.cuda.synchronize()
    tg0 = time.perf_counter()
    allrec = multi.gather_records(rec[:cnt], dist, dev)
    torch.cuda.synchronize()
    gather_ms = (time.perf_counter() - tg0) * 1e3
    pos_all = allrec >> 24
    gather_info = {"records": int(allrec.numel()), "slice_bytes_per_rank": ng, "ms": round(gather_ms, 3),
                   "position_sorted": bool((pos_all[1:] > pos_all[:-1]).all().item()) if allrec.numel() > 1 else True}
    del rec, allrec, pos_all

    # end to end through the public host-buffer call: pinned input, H2D, scan, D2H of the dense result
    ne = min(args.e2e_mib << 20, n)
    hin = pm.PinnedBuffer(ne); hout = pm.PinnedBuffer(2 * ne)
    a_in = hin.array(np.uint8); a_out = hout.array(np.uint16)
    torch.cuda.synchronize()
    tmp = buf[lead:lead + ne].cpu().numpy()
    a_in[:] = tmp
    del tmp
    e2e_steps = max(3, min(args.steps, 5))
    for _ in range(2):
        eng.reset(); eng.scan_host_ptr(hin.ptr, ne, hout.ptr, algo=algo)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.reset(); eng.scan_host_ptr(hin.ptr, ne, hout.ptr, algo=algo)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    # the e2e result must equal the device-resident one (same bytes, same kernel, through the host path)
    e2e_ok = bool(np.array_equal(a_out[:ne].view(np.int16), out[:ne].cpu().numpy())) if not have_lead else None
    # informational: the same host call with a SPARSE result (records of the matches with >= 4 pattern bytes):
    # the dense 2 B/position never crosses PCIe.  Not the headline -- the reference's contract is the dense result.
    rcap = max(ne // 64, 1 << 20)
    hrec = pm.PinnedBuffer(8 * rcap)
    for _ in range(2):
        eng.reset(); eng.scan_host_records(None, min_len=4, cap=rcap, algo=algo, src_ptr=hin.ptr, n=ne, dst_ptr=hrec.ptr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.reset(); _, n_rec = eng.scan_host_records(None, min_len=4, cap=rcap, algo=algo, src_ptr=hin.ptr, n=ne, dst_ptr=hrec.ptr)
    rec_s = (time.perf_counter() - t0) / e2e_steps
    t_r = torch.tensor([rec_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_r, op=dist.ReduceOp.MAX)
    e2e_records = {"value": world * ne / float(t_r.item()) / 1e9, "unit": "GB/s", "min_pattern_len": 4,
                   "records_per_step": int(n_rec), "h2d_bytes_per_step": ne, "d2h_bytes_per_step": int(min(n_rec, rcap)) * 8,
                   "call": "pm_engine_scan_host_records"}
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_gbs = world * ne / float(t_e.item()) / 1e9
    # informational: what the host link of THIS rank gives a plain pinned copy, so that e2e can be read against it
    # (the dense result is 2 B per stream byte: e2e <= d2h / 2)
    pcie = None
    if rank == 0:
        try:
            mb = 256 << 20
            h_a = torch.empty(mb, dtype=torch.uint8, pin_memory=True); h_b = torch.empty(mb, dtype=torch.uint8, pin_memory=True)
            d_a = buf[:mb]; d_b = out.view(torch.uint8)[:mb]
            s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

            def timed(fn, reps=4):
                fn(); torch.cuda.synchronize()
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

            h_a.copy_(d_a); torch.cuda.synchronize()   # keep the stream bytes intact: d_a gets back what it held
            t_d2h = timed(lambda: h_b.copy_(d_b, non_blocking=True))
            t_h2d = timed(lambda: d_a.copy_(h_a, non_blocking=True))
            t_both = timed(both)
            pcie = {"d2h_gbs": round(mb / t_d2h / 1e9, 1), "h2d_gbs": round(mb / t_h2d / 1e9, 1),
                    "both_directions_gbs_each": round(mb / t_both / 1e9, 1), "bytes": mb,
                    "note": "plain pinned cudaMemcpyAsync on rank 0, nothing else running"}
            del h_a, h_b
        except Exception as ex:  # measurement aid only
            pcie = {"error": str(ex)[:120]}

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    main_ms_avg = main_ms / max(n_prof, 1)
    alg_bytes = 3.0 * n                                   # 1 B stream read + 2 B dense result written per position
    achieved = alg_bytes / (main_ms_avg * 1e-3) / 1e9 if n_prof else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tp):
        try:
            tr = json.load(open(tp))
            traffic = tr["dram_bytes_per_stream_byte"] * n   # ncu --set full capture, scaled per stream byte
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": f"C3: snort.dict+et.dict merged ({d.n_patterns} patterns, max_pat_len {d.max_pat_len}), "
                               f"S-{args.stream} stream, {args.gib:g} GiB per GPU, shards with {pm.HALO}-byte halo",
                   "algo": args.algo, "result": "dense uint16 longest-match pid per position",
                   "bytes_per_gpu": n, "l2": "input 16 GiB per step >> 126 MB L2 (no flush needed)",
                   "timing": "CUDA events on the launching stream, max over ranks"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "kernel": "sfx_scan_kernel", "algorithmic_bytes_per_stream_byte": 3,
                     "kernel_ms": main_ms_avg, "kernel_share_of_step": (main_ms / scan_ms) if scan_ms else None,
                     "peak_source": peak_src},
        "e2e": {"value": e2e_gbs, "unit": "GB/s", "h2d_bytes_per_step": ne, "d2h_bytes_per_step": 2 * ne,
                "call": "pm_engine_scan_host (pinned host buffers, 16 MiB double-buffered chunks)", "steps": e2e_steps,
                "matches_device_result": e2e_ok, "host_link": pcie},
        "e2e_records": e2e_records,
        "gpu_launches": int(launches),
        "clocks": sampler.result(),
        "result_check": {"positions_with_match": red["positions"], "matches_with_ancestors": red["matches"],
                         "digest_sum_longest": "%016x" % red["hsum_longest"], "digest_sum_all": "%016x" % red["hsum_all"],
                         "note": "sums over all ranks; equal to one continuous scan of the whole stream"},
        "record_gather": gather_info,
    }
    if not args.no_cpu_baseline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            from reflib import Reference
            cores = os.cpu_count() or 1
            sample = min(args.ref_mib << 20, cores * (32 << 20), n)
            host = buf[lead:lead + sample].cpu().numpy()
            ref = Reference(DICTS, algo_mask=1)
            r = ref.scan_parallel(host, cores)
            line["cpu_baseline"] = {"value": sample / r["max_loop_seconds"] / 1e9, "unit": "GB/s", "cores": cores,
                                    "kind": "reference",
                                    "sample": f"first {sample >> 20} MiB of this rank's stream, reference ac_read_char loop "
                                              f"(gcc -O2), fork per core with halo; time = slowest worker's loop"}
            # the GPU result on the same sample must carry the reference's digest
            sg = eng.summarize(out, sample, pos_base=off)
            line["cpu_baseline"]["gpu_matches_reference"] = bool(
                sg["positions"] == r["positions"] and sg["matches"] == r["matches"] and
                sg["hsum_longest"] == r["hsum_longest"] and sg["hsum_all"] == r["hsum_all"]) if rank == 0 and off == 0 else None
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    emit(line)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
