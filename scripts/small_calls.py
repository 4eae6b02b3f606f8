"""Latency of the host-buffer entry point for small chunks (the reference driver feeds 100 KiB buffers,
measure.c:77,281-304).  Usage: python scripts/small_calls.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import patternmatching_b200 as pm

DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "data")
d = pm.Dictionary().add_file(os.path.join(DATA, "snort.dict")).add_file(os.path.join(DATA, "et.dict")).compile()
eng = pm.Engine(d)
rng = np.random.default_rng(1)
for algo_name in ("sfx", "auto"):
    algo = pm.ALGOS[algo_name]
    for size in (1, 4096, 100 * 1024, 1 << 20, 16 << 20):
        buf = rng.integers(0, 256, size, dtype=np.uint8)
        hin = pm.PinnedBuffer(size); hout = pm.PinnedBuffer(2 * size)
        hin.array(np.uint8)[:] = buf
        for pinned in (False, True):
            eng.reset()
            reps = 200 if size <= (1 << 20) else 20
            for _ in range(5):
                (eng.scan_host_ptr(hin.ptr, size, hout.ptr, algo=algo) if pinned else eng.scan_host(buf, algo=algo))
            t0 = time.perf_counter()
            for _ in range(reps):
                (eng.scan_host_ptr(hin.ptr, size, hout.ptr, algo=algo) if pinned else eng.scan_host(buf, algo=algo))
            dt = (time.perf_counter() - t0) / reps
            print(f"{algo_name:5s} {size:9d} B  {'pinned  ' if pinned else 'pageable'}  {dt * 1e6:9.1f} us/call  {size / dt / 1e9:8.3f} GB/s", flush=True)
