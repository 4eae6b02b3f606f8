#!/usr/bin/env python
"""Generate tests/golden/*.json from the REFERENCE ITSELF (oracle/_ref/libpmref.so, built from the
unmodified sources under /root/reference by oracle/Makefile) and cross-check our oracle restatement
(oracle/liboracle.so) against it.  Run in the build container only (needs /root/reference):

    python scripts/make_golden.py            # all configs (each in its own process: the
                                             # reference keeps global state and is not re-entrant)
"""
import ctypes as C, json, os, subprocess, sys, zlib
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "oracle", "_ref", "data")
GOLD = os.path.join(ROOT, "tests", "golden")
CONFIGS = {
    "snort": ["snort.dict"],
    "et": ["et.dict"],
    "merged": ["snort.dict", "et.dict"],
}


def load_ref():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libpmref.so"))
    lib.pmref_build.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int]
    lib.pmref_n_patterns.restype = C.c_size_t
    lib.pmref_max_pat_len.restype = C.c_size_t
    lib.pmref_total_mem.restype = C.c_size_t
    lib.pmref_total_mem.argtypes = [C.c_int]
    lib.pmref_scan.restype = C.c_double
    lib.pmref_scan.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.pmref_summary.restype = C.c_double
    lib.pmref_summary.argtypes = [C.c_int, C.c_void_p, C.c_size_t] + [C.POINTER(C.c_uint64)] * 3
    lib.pmref_last_hsum.argtypes = [C.POINTER(C.c_uint64)]
    lib.pmref_success.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
    lib.pmref_pattern.argtypes = [C.c_size_t] + [C.POINTER(C.c_uint32)] * 5 + [C.POINTER(C.POINTER(C.c_ubyte))]
    return lib


def run_config(name):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle  # our restatement (ctypes wrapper)
    dicts = [os.path.join(DATA, d) for d in CONFIGS[name]]
    stream = np.fromfile(os.path.join(DATA, "dictionaries_generated.stream"), dtype=np.uint8)
    ref = load_ref()
    arr = (C.c_char_p * len(dicts))(*[d.encode() for d in dicts])
    algo_mask = 0b011 if name != "snort" else 0b111   # MPBG (38 s) only on config #1
    assert ref.pmref_build(len(dicts), arr, algo_mask) == 0
    n = stream.size
    fo = np.empty(n, np.uint32); lo = np.empty(n, np.uint32)
    ref.pmref_reset(0)
    ref.pmref_scan(0, stream.ctypes.data, n, fo.ctypes.data, lo.ctypes.data)
    pos = C.c_uint64(); mat = C.c_uint64(); chk = C.c_uint64()
    ref.pmref_summary(0, stream.ctypes.data, n, C.byref(pos), C.byref(mat), C.byref(chk))
    hs = (C.c_uint64 * 2)(); ref.pmref_last_hsum(hs)
    # LMAC must equal AC (results.csv:3 -> 0/0/0)
    cnt = (C.c_uint64 * 4)()
    ref.pmref_success(1, stream.ctypes.data, n, cnt)
    lmac_counts = list(cnt)
    mpbg_counts = None
    mf = ml = None
    if algo_mask & 4:
        ref.pmref_success(2, stream.ctypes.data, n, cnt)
        mpbg_counts = list(cnt)
        # what the reference's MPBG reports at every position (mpbg.c:132-145 as shipped)
        mf = np.empty(n, np.uint32); ml = np.empty(n, np.uint32)
        ref.pmref_reset(2)
        ref.pmref_scan(2, stream.ctypes.data, n, mf.ctypes.data, ml.ctypes.data)
    # reference pattern list
    P = ref.pmref_n_patterns()
    pats = {}
    f = C.c_uint32(); l = C.c_uint32(); pf = C.c_uint32(); pl = C.c_uint32(); ln = C.c_uint32()
    bp = C.POINTER(C.c_ubyte)()
    for i in range(P):
        ref.pmref_pattern(i, C.byref(f), C.byref(l), C.byref(pf), C.byref(pl), C.byref(ln), C.byref(bp))
        pats[(f.value, l.value)] = (bytes(bytearray(bp[:ln.value])), (pf.value, pl.value))

    # ---- cross-check our restatement against the reference ----
    o = Oracle()
    for d in dicts:
        o.add_dict_file(d)
    o.compile()
    assert o.n_patterns == P, (o.n_patterns, P)
    n_states_ref = (ref.pmref_total_mem(0) - 24) // 2072
    assert o.n_states == n_states_ref, (o.n_states, n_states_ref)
    assert o.max_pat_len == ref.pmref_max_pat_len()
    crc = 0
    for i in range(P):
        file, line, parent, b = o.pattern(i)
        rb, rpar = pats[(file, line)]
        assert rb == b, (file, line)
        par_fl = (0xFFFFFFFF, 0xFFFFFFFF) if parent < 0 else o.pattern(parent)[:2]
        assert tuple(par_fl) == rpar, (file, line, par_fl, rpar)
        crc = zlib.crc32(b, zlib.crc32(np.array([file, line, len(b)], np.uint32).tobytes(), crc))
    longest = o.scan(stream)
    files, lines = o.id_arrays()
    of = np.where(longest >= 0, files[np.maximum(longest, 0)], 0xFFFFFFFF).astype(np.uint32)
    ol = np.where(longest >= 0, lines[np.maximum(longest, 0)], 0xFFFFFFFF).astype(np.uint32)
    assert np.array_equal(of, fo) and np.array_equal(ol, lo), "oracle longest-match differs from reference"
    if mf is not None:
        # The reference's MPBG never reports a pattern of more than 8 bytes (its fingerprint stages do not fire);
        # patterns of <= 8 bytes go through its exact KMP (bgps.c:459-464).  So its answer is the longest pattern of <= 8
        # bytes ending at the position = the first pattern of <= 8 bytes on the PatternsTree chain of the AC's answer.
        want_f = np.full(n, 0xFFFFFFFF, np.uint32); want_l = np.full(n, 0xFFFFFFFF, np.uint32)
        for i in range(n):
            q = int(longest[i])
            while q >= 0 and len(o.pattern(q)[3]) > 8:
                q = o.pattern(q)[2]
            if q >= 0:
                want_f[i], want_l[i] = o.pattern(q)[0], o.pattern(q)[1]
        assert np.array_equal(want_f, mf) and np.array_equal(want_l, ml), "reference MPBG differs from the <= 8-byte rule"
    s = o.summary(stream)
    assert (s.positions, s.matches, s.fnv) == (pos.value, mat.value, chk.value), (s.positions, s.matches, hex(s.fnv))
    assert (s.hsum_longest, s.hsum_all) == (hs[0], hs[1])

    gold = {
        "config": name, "dicts": CONFIGS[name], "stream": "dictionaries_generated.stream",
        "generated_by": "scripts/make_golden.py from oracle/_ref/libpmref.so (unmodified reference sources)",
        "n_patterns": P, "n_states": int(n_states_ref), "max_pat_len": int(ref.pmref_max_pat_len()),
        "ac_total_mem": int(ref.pmref_total_mem(0)), "lmac_total_mem": int(ref.pmref_total_mem(1)),
        "positions": pos.value, "matches": mat.value, "fnv": "%016x" % chk.value,
        "hsum_longest": "%016x" % hs[0], "hsum_all": "%016x" % hs[1],
        "lmac_vs_ac_counts": lmac_counts, "mpbg_vs_ac_counts": mpbg_counts,
        "patterns_crc32": "%08x" % crc,
        "n_lines": o.n_lines, "n_rejected": o.n_rejected, "n_duplicates": o.n_duplicates,
        # per-position longest match of the reference AC as (file,line); 0xFFFFFFFF = none
        "longest_file": [int(x) if x != 0xFFFFFFFF else -1 for x in fo],
        "longest_line": [int(x) if x != 0xFFFFFFFF else -1 for x in lo],
    }
    if algo_mask & 4:
        gold["mpbg_total_mem"] = int(ref.pmref_total_mem(2))
        # per-position answer of the reference's MPBG, same encoding as longest_*
        gold["mpbg_file"] = [int(x) if x != 0xFFFFFFFF else -1 for x in mf]
        gold["mpbg_line"] = [int(x) if x != 0xFFFFFFFF else -1 for x in ml]
    os.makedirs(GOLD, exist_ok=True)
    with open(os.path.join(GOLD, f"ref_{name}.json"), "w") as fh:
        json.dump(gold, fh, separators=(",", ":"))
    print(name, {k: v for k, v in gold.items() if not k.startswith("longest_")})


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_config(sys.argv[1])
    else:
        for name in CONFIGS:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), name])
