"""Quick device-resident timing of the scan kernels (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import patternmatching_b200 as pm

DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "data")
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 30
kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["planted", "uniform", "almost"]
algos = sys.argv[3].split(",") if len(sys.argv) > 3 else ["sfx", "dfa", "kr"]
t0 = time.time()
d = pm.Dictionary().add_file(os.path.join(DATA, "snort.dict")).add_file(os.path.join(DATA, "et.dict")).compile()
t1 = time.time()
eng = pm.Engine(d)
t2 = time.time()
print(f"dict compile {t1 - t0:.2f}s  engine upload {t2 - t1:.2f}s  tables {eng.total_mem / 1e6:.1f} MB", flush=True)
dev = torch.device("cuda:0")
buf = torch.empty(n, dtype=torch.uint8, device=dev)
out = torch.empty(n, dtype=torch.int16, device=dev)
for kind in kinds:
    eng.generate(kind, 0, n, buf); torch.cuda.synchronize()
    for a in algos:
        algo = pm.ALGOS[a]
        eng.scan_device(buf, n, out, algo=algo); torch.cuda.synchronize()   # warm-up (+ lazy table build)
        ms = eng.time_scan(buf, n, out, algo=algo, iters=3)
        if a == "sfx":
            eng.set_profiling(True)
            for _ in range(3):
                eng.scan_device(buf, n, out, algo=algo)
            np_, mainms, totms = eng.read_profile(); eng.set_profiling(False)
            print(f"   sfx main kernel {mainms / np_:.3f} ms, whole scan {totms / np_:.3f} ms, deferred queue slots {eng.last_deferred}")
        s = eng.summarize(out, n)
        print(f"{kind:8s} {a:4s} {ms:9.3f} ms  {n / ms / 1e6:9.1f} GB/s stream  {3 * n / ms / 1e6:9.1f} GB/s alg(3B/B)  pos={s['positions']} matches={s['matches']} h={s['hsum_all']:016x}", flush=True)
