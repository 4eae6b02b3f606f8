"""One device-resident scan of one stream kind with one algorithm (a short command line for ncu).
    python scripts/one_scan.py BYTES KIND ALGO [ENV=VALUE ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
n, kind, algo = int(float(sys.argv[1])), sys.argv[2], sys.argv[3]
for kv in sys.argv[4:]:
    k, v = kv.split("=", 1)
    os.environ[k] = v
import torch
import patternmatching_b200 as pm
DATA = os.path.join(ROOT, "oracle", "_ref", "data")
if kind == "ab":
    pats = [b"a" * k for k in range(1, 257)]
    for L in range(1, 13):
        for v in range(1 << L):
            pats.append(bytes(97 + ((v >> i) & 1) for i in range(L)))
    d = pm.Dictionary().add_bytes(b"\n".join(pats) + b"\n").compile()
else:
    d = pm.Dictionary().add_file(os.path.join(DATA, "snort.dict")).add_file(os.path.join(DATA, "et.dict")).compile()
eng = pm.Engine(d)
dev = torch.device("cuda:0")
buf = torch.empty(n, dtype=torch.uint8, device=dev)
out = torch.empty(n, dtype=torch.int16, device=dev)
eng.generate(kind, 0, n, buf)
for _ in range(3):
    eng.scan_device(buf, n, out, algo=pm.ALGOS[algo])
torch.cuda.synchronize()
ms = eng.time_scan(buf, n, out, algo=pm.ALGOS[algo], iters=3)
print(f"{kind} {algo} {n} bytes: {ms:.3f} ms {n / ms / 1e6:.1f} GB/s", flush=True)
