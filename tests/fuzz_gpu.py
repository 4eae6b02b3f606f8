"""Differential fuzzing of the GPU kernels against the oracle: random dictionaries (alphabet size, pattern
lengths, shared suffixes / prefixes, nested patterns) x random streams with planted occurrences x every exact
kernel, and the randomized one against its restatement.  Test infrastructure (it drives the oracle).
Usage: python tests/fuzz_gpu.py [n_cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import patternmatching_b200 as pm
from oracle_lib import Oracle


def make_case(rng):
    alpha = int(rng.choice([2, 3, 4, 16, 64, 200, 256]))
    base = int(rng.integers(0, 257 - alpha))
    sym = np.array([b for b in range(base, base + alpha) if b != 10] or [97], dtype=np.uint8)
    n_pat = int(rng.choice([1, 5, 50, 500, 5000]))
    max_len = int(rng.choice([1, 2, 3, 4, 5, 8, 9, 17, 64, 353]))
    pats, pat_list = set(), []   # the list keeps insertion order: cases depend on the seed only, not on hash randomisation
    stems = [bytes(rng.choice(sym, int(rng.integers(1, max_len + 1)))) for _ in range(8)]
    tries = 0
    while len(pats) < n_pat and tries < 20 * n_pat:
        tries += 1
        L = int(rng.integers(1, max_len + 1))
        p = bytes(rng.choice(sym, L))
        mode = rng.integers(0, 4)
        if mode == 1:   # shared suffix
            st = stems[int(rng.integers(0, 8))]; p = (p + st)[-max_len:]
        elif mode == 2:  # shared prefix
            st = stems[int(rng.integers(0, 8))]; p = (st + p)[:max_len]
        elif mode == 3 and pats:  # nested: a suffix of an existing pattern
            q = pat_list[int(rng.integers(0, len(pat_list)))]; p = q[int(rng.integers(0, len(q))):]
        if p and p not in pats:
            pats.add(p); pat_list.append(p)
    pats = sorted(pats)
    n = int(rng.choice([1, 100, 511, 512, 513, 5000, 70000, 300001]))
    stream = rng.choice(sym, n).astype(np.uint8)
    if rng.integers(0, 2):   # some bytes outside the alphabet
        k = max(1, n // 50)
        stream[rng.integers(0, n, k)] = rng.integers(0, 256, k).astype(np.uint8)
    for _ in range(int(rng.integers(0, 1 + n // 20))):   # planted occurrences
        p = pats[int(rng.integers(0, len(pats)))]
        c = int(rng.integers(0, n))
        stream[c:c + len(p)] = np.frombuffer(p, np.uint8)[: n - c]
    hist = int(rng.choice([0, 0, 1, 3, 4, 15, 16, 17, 351, 352, 353, 1000]))
    return pats, stream, hist


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda:0")
    bad = 0
    for case in range(n_cases):
        pats, stream, hist = make_case(rng)
        d = pm.Dictionary(); o = Oracle()
        for i, p in enumerate(pats):
            d.add_pattern(p, 0, i + 1); o.add_pattern(p, 0, i + 1)
        d.compile(); o.compile()
        eng = pm.Engine(d)
        os.environ["PM_DFA_DEEP"] = "1"      # a second engine whose DFA scans take the compact-record walker (deep_scan.cu)
        eng_deep = pm.Engine(d)
        del os.environ["PM_DFA_DEEP"]
        hist = min(hist, stream.size - 1) if stream.size > 1 else 0
        body = stream[hist:]
        o.reset()
        want_all = (o.scan(stream) + 1).astype(np.uint16)
        want = want_all[hist:]
        pad = (-hist) % 16
        buf = np.concatenate([np.zeros(pad, np.uint8), stream])
        d_in = torch.from_numpy(buf).to(dev)
        for algo, name in ((pm.ALGO_SFX, "sfx"), (pm.ALGO_DFA, "dfa"), (pm.ALGO_AUTO, "auto"), (pm.ALGO_DFA, "deep")):
            d_out = torch.zeros(max(body.size, 8), dtype=torch.int16, device=dev)
            (eng_deep if name == "deep" else eng).scan_device(d_in.data_ptr() + pad + hist, body.size, d_out, hist_valid=hist, algo=algo)
            torch.cuda.synchronize()
            got = d_out.cpu().numpy().view(np.uint16)[:body.size]
            if not np.array_equal(got, want):
                bad += 1
                w = np.nonzero(got != want)[0]
                print(f"MISMATCH case {case} algo {name}: {len(pats)} patterns, n={body.size}, hist={hist}, first diff at {w[:5]} got {got[w[:5]]} want {want[w[:5]]}", flush=True)
        # 32-bit results and the reference MPBG's behaviour (longest pattern of <= 8 bytes on the answer's PatternsTree chain)
        d_out32 = torch.zeros(max(body.size, 8), dtype=torch.int32, device=dev)
        eng.scan_device32(d_in.data_ptr() + pad + hist, body.size, d_out32, hist_valid=hist, algo=pm.ALGO_AUTO)
        torch.cuda.synchronize()
        if not np.array_equal(d_out32.cpu().numpy().view(np.uint32)[:body.size], want.astype(np.uint32)):
            bad += 1
            print(f"MISMATCH case {case} scan_device32", flush=True)
        lens, par = o.lengths(), o.parents()
        short_of = np.arange(len(pats) + 1, dtype=np.int64)
        for q in range(1, len(pats) + 1):
            r = q - 1
            while r >= 0 and lens[r] > 8:
                r = int(par[r])
            short_of[q] = r + 1
        d_out = torch.zeros(max(body.size, 8), dtype=torch.int16, device=dev)
        eng.scan_device(d_in.data_ptr() + pad + hist, body.size, d_out, hist_valid=hist, algo=pm.ALGO_MPBG)
        torch.cuda.synchronize()
        if not np.array_equal(d_out.cpu().numpy().view(np.uint16)[:body.size], short_of[want].astype(np.uint16)):
            bad += 1
            print(f"MISMATCH case {case} mpbg mode", flush=True)
        # the randomized variant against the oracle's restatement of it (whole stream, no history)
        seed = 0xF1A90000 + case
        eng.set_kr_seed(seed)
        d_all = torch.from_numpy(stream.copy()).to(dev)
        d_out = torch.zeros(max(stream.size, 8), dtype=torch.int16, device=dev)
        eng.scan_device(d_all, stream.size, d_out, hist_valid=0, algo=pm.ALGO_KR)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().view(np.uint16)[:stream.size]
        want_kr = (o.kr_scan(stream, seed) + 1).astype(np.uint16)
        if not np.array_equal(got, want_kr):
            bad += 1
            w = np.nonzero(got != want_kr)[0]
            print(f"MISMATCH case {case} algo kr: {len(pats)} patterns, n={stream.size}, first diff at {w[:5]} got {got[w[:5]]} want {want_kr[w[:5]]}", flush=True)
        # host path with arbitrary cuts
        eng.reset()
        cuts = sorted(set([0, stream.size] + [int(x) for x in rng.integers(0, stream.size + 1, 3)]))
        outs = [eng.scan_host(stream[a:b], algo=pm.ALGO_SFX) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        got = np.concatenate(outs) if outs else np.zeros(0, np.uint16)
        if not np.array_equal(got, want_all):
            bad += 1
            print(f"MISMATCH case {case} scan_host cuts {cuts}", flush=True)
        print(f"case {case}: {len(pats)} patterns (max len {max(map(len, pats))}), classes {d.info.n_classes}, n={body.size}, hist={hist}: ok" if not bad else f"case {case} done", flush=True)
    print("FUZZ RESULT:", "all equal" if bad == 0 else f"{bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
