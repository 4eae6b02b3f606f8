"""Throughput of every kernel on every BASELINE.json config (development aid; results go to DESIGN.md)."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import patternmatching_b200 as pm

DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "data")
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 30
dev = torch.device("cuda:0")
buf = torch.empty(n, dtype=torch.uint8, device=dev)
out = torch.empty(n, dtype=torch.int16, device=dev)
ref = torch.empty(n, dtype=torch.int16, device=dev)


def adv_dict():
    # C5a: a^k for k = 1..256 plus every string over {a,b} of length 1..12
    pats = [b"a" * k for k in range(1, 257)]
    for L in range(1, 13):
        for v in range(1 << L):
            pats.append(bytes(97 + ((v >> i) & 1) for i in range(L)))
    return pm.Dictionary().add_bytes(b"\n".join(pats) + b"\n").compile()


def run(tag, d, kinds, algos):
    eng = pm.Engine(d)
    i = d.info
    print(f"== {tag}: {i.n_patterns} patterns, {i.n_ac_states} AC states, {i.n_sfx_rows} suffix rows, {i.n_classes} classes, tables {eng.total_mem / 1e6:.1f} MB", flush=True)
    for kind in kinds:
        eng.generate(kind, 0, n, buf); torch.cuda.synchronize()
        base = None
        for a in algos:
            algo = pm.ALGOS[a]
            eng.scan_device(buf, n, out, algo=algo); torch.cuda.synchronize()
            ms = eng.time_scan(buf, n, out, algo=algo, iters=3)
            s = eng.summarize(out, n)
            if a == "sfx":
                ref.copy_(out); base = s
            same = "" if a == "kr" or base is None else (" ==sfx" if bool(torch.equal(out, ref)) else " DIFFERS")
            extra = ""
            if a == "kr" and base is not None:
                extra = f" kr-vs-exact: positions that differ {int((out != ref).sum().item())}"
            print(f"{kind:8s} {a:4s} {ms:9.3f} ms {n / ms / 1e6:8.1f} GB/s  pos={s['positions']} matches={s['matches']}{same}{extra}", flush=True)
    del eng


which = sys.argv[2] if len(sys.argv) > 2 else "all"
if which in ("all", "merged"):
    d = pm.Dictionary().add_file(os.path.join(DATA, "snort.dict")).add_file(os.path.join(DATA, "et.dict")).compile()
    run("snort+et", d, ["planted", "uniform", "ascii", "almost"], ["sfx", "dfa", "kr"])
if which in ("all", "adv"):
    run("C5a adversarial {a,b}", adv_dict(), ["ab"], ["sfx", "dfa"])
