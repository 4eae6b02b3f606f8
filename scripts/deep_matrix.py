"""Throughput of the forward-walker variants on the deep-match configs (C5a adversarial {a,b}, C5b almost) --
development aid; every variant is compared with the backward scan's result for equality.
    python scripts/deep_matrix.py [bytes]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import patternmatching_b200 as pm

DATA = os.path.join(ROOT, "oracle", "_ref", "data")
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 30
dev = torch.device("cuda:0")
buf = torch.empty(n, dtype=torch.uint8, device=dev)
out = torch.empty(n, dtype=torch.int16, device=dev)
ref = torch.empty(n, dtype=torch.int16, device=dev)


def adv_dict():
    pats = [b"a" * k for k in range(1, 257)]
    for L in range(1, 13):
        for v in range(1 << L):
            pats.append(bytes(97 + ((v >> i) & 1) for i in range(L)))
    return pm.Dictionary().add_bytes(b"\n".join(pats) + b"\n").compile()


def run(tag, d, kinds, variants):
    print(f"== {tag}: {d.info.n_patterns} patterns, {d.info.n_ac_states} AC states", flush=True)
    base = pm.Engine(d)
    for kind in kinds:
        base.generate(kind, 0, n, buf)
        base.scan_device(buf, n, ref, algo=pm.ALGO_SFX if tag.startswith("snort") else pm.ALGO_DFA)
        torch.cuda.synchronize()
        for name, env, algo in variants:
            for k, v in env.items():
                os.environ[k] = v
            eng = pm.Engine(d)
            for k in env:
                del os.environ[k]
            eng.scan_device(buf, n, out, algo=algo); torch.cuda.synchronize()
            ms = eng.time_scan(buf, n, out, algo=algo, iters=3)
            same = "==ref" if bool(torch.equal(out, ref)) else "DIFFERS"
            print(f"{kind:8s} {name:14s} {ms:9.3f} ms {n / ms / 1e6:8.1f} GB/s  {same}  auto_choice={eng.auto_choice} tables={eng.total_mem / 1e6:.0f} MB", flush=True)
            del eng


which = sys.argv[2] if len(sys.argv) > 2 else "all"
if which in ("all", "merged"):
    d = pm.Dictionary().add_file(os.path.join(DATA, "snort.dict")).add_file(os.path.join(DATA, "et.dict")).compile()
    run("snort+et", d, ["almost", "ascii", "planted"],
        [("auto", {}, pm.ALGO_AUTO), ("deep-records", {"PM_DFA_DEEP": "1"}, pm.ALGO_DFA), ("flat-dense", {"PM_DFA_FLAT": "1"}, pm.ALGO_DFA),
         ("hot", {}, pm.ALGO_DFA)])
if which in ("all", "adv"):
    run("C5a adversarial {a,b}", adv_dict(), ["ab"],
        [("auto", {}, pm.ALGO_AUTO), ("fused-small", {}, pm.ALGO_DFA), ("hot-2gather", {"PM_DFA_NO_FUSED": "1"}, pm.ALGO_DFA)])
