"""Throughput of the host-facing calls in the buffer regimes the reference can present:
gpu_read_block (8-byte pattern ids out, what the patched measure.c calls) and pm_engine_scan_host (dense uint16 out),
with pageable / page-locked buffers and 100 KiB (measure.c:77) / 16 MiB / whole-stream calls.  Prints one JSON object.

    python scripts/host_regimes.py [total MiB] [threads,threads,...] [PM_HOST_IDS values, e.g. host,5,3,device]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import patternmatching_b200 as pm

DATA = os.path.join(ROOT, "oracle", "_ref", "data")
total = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
thread_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
ids_list = sys.argv[3].split(",") if len(sys.argv) > 3 else [None]


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    import torch
    d = pm.Dictionary().add_file(os.path.join(DATA, "snort.dict")).add_file(os.path.join(DATA, "et.dict")).compile()
    gen = pm.Engine(d)
    dev = torch.device("cuda:0")
    buf = torch.empty(total, dtype=torch.uint8, device=dev)
    gen.generate("planted", 0, total, buf)
    stream = buf.cpu().numpy()
    del buf
    P = d.n_patterns
    out = {"total_bytes": total, "runs": []}
    for threads, ids_mode in [(t, i) for t in thread_list for i in ids_list]:
        if threads:
            os.environ["PM_HOST_THREADS"] = str(threads)
        if ids_mode is not None:
            os.environ["PM_HOST_IDS"] = ids_mode
        m = pm.MpsGpu("sfx")
        for pid in range(1, P + 1):
            m.add_pattern(d.pattern(pid)[4], 0x7F0000000000 + 64 * pid)
        m.compile()
        eng = pm.Engine(d)
        hin = pm.PinnedBuffer(total); hout = pm.PinnedBuffer(8 * total)
        hin.array(np.uint8)[:] = stream
        ids = np.empty(total, np.uint64); u16 = np.empty(total, np.uint16)
        ids[:] = 0; u16[:] = 0                           # first touch outside the timed region
        res = {"host_threads": eng.host_threads, "PM_HOST_IDS": ids_mode}

        def blocks(call, step, n):
            def run():
                call_reset()
                for o in range(0, n, step):
                    call(o, min(step, n - o))
            return run

        # gpu_read_block, 8-byte ids out
        call_reset = m.reset
        n_small = min(total, 64 << 20)
        t = timed(blocks(lambda o, k: m.read_block_ptr(stream.ctypes.data + o, k, ids.ctypes.data + 8 * o), 100 * 1024, n_small), 2)
        res["read_block_pageable_100KiB"] = n_small / t / 1e9
        t = timed(blocks(lambda o, k: m.read_block_ptr(stream.ctypes.data + o, k, ids.ctypes.data + 8 * o), 16 << 20, total), 2)
        res["read_block_pageable_16MiB"] = total / t / 1e9
        t = timed(blocks(lambda o, k: m.read_block_ptr(stream.ctypes.data + o, k, ids.ctypes.data + 8 * o), total, total), 2)
        res["read_block_pageable_whole"] = total / t / 1e9
        t = timed(blocks(lambda o, k: m.read_block_ptr(hin.ptr + o, k, hout.ptr + 8 * o), total, total), 2)
        res["read_block_pinned_whole"] = total / t / 1e9
        t = timed(blocks(lambda o, k: m.read_block_ptr(hin.ptr + o, k, hout.ptr + 8 * o), 16 << 20, total), 2)
        res["read_block_pinned_16MiB"] = total / t / 1e9
        t = timed(blocks(lambda o, k: m.read_block_ptr(hin.ptr + o, k, hout.ptr + 8 * o), 100 * 1024, n_small), 2)
        res["read_block_pinned_100KiB"] = n_small / t / 1e9
        if ids_mode is not None and (threads, ids_mode) != (thread_list[0], ids_list[0]):   # the id sweep only needs the lines above
            out["runs"].append({k: (round(v, 3) if isinstance(v, float) else v) for k, v in res.items()})
            print(json.dumps(out["runs"][-1]), flush=True)
            m.free(); del eng, hin, hout
            continue
        t = timed(lambda: [m.read_char(int(c)) for c in stream[:2000]], 1)
        res["read_char_us"] = t / 2000 * 1e6
        # pm_engine_scan_host, dense uint16 out
        call_reset = eng.reset
        t = timed(blocks(lambda o, k: eng.scan_host_ptr(stream.ctypes.data + o, k, u16.ctypes.data + 2 * o), 100 * 1024, n_small), 2)
        res["scan_host_pageable_100KiB"] = n_small / t / 1e9
        t = timed(blocks(lambda o, k: eng.scan_host_ptr(stream.ctypes.data + o, k, u16.ctypes.data + 2 * o), 16 << 20, total), 2)
        res["scan_host_pageable_16MiB"] = total / t / 1e9
        t = timed(blocks(lambda o, k: eng.scan_host_ptr(stream.ctypes.data + o, k, u16.ctypes.data + 2 * o), total, total), 2)
        res["scan_host_pageable_whole"] = total / t / 1e9
        t = timed(blocks(lambda o, k: eng.scan_host_ptr(hin.ptr + o, k, hout.ptr + 2 * o), total, total), 2)
        res["scan_host_pinned_whole"] = total / t / 1e9
        # plain host-memory rates for reading the numbers above: a threaded copy and a threaded 8-byte fill
        out["runs"].append({k: (round(v, 3) if isinstance(v, float) else v) for k, v in res.items()})
        print(json.dumps(out["runs"][-1]), flush=True)
        m.free(); del eng, hin, hout
    print(json.dumps(out))


if __name__ == "__main__":
    main()
