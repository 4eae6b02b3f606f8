"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol that
include/pm_b200.h declares, the host dictionary compiler reproduces the reference's ingest
(de-dup, ids, parents, state count), and there is no CPU fallback for scanning."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import patternmatching_b200 as pm
from conftest import GOLDEN, ROOT, TINY_DICT, dict_paths
from oracle_lib import Oracle, parse_line


def declared_functions():
    text = open(os.path.join(ROOT, "include", "pm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef struct \{.*?\} \w+;", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:pm|gpu|mps)_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 35
    L = C.CDLL(pm.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert pm.lib().pm_version() >= 1


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is exercised on the CPU box")
    d = pm.Dictionary().add_bytes(b"he\nshe\n").compile()
    with pytest.raises(pm.PmError, match="no CUDA device"):
        pm.Engine(d)


@pytest.mark.parametrize("name", ["snort", "et", "merged"])
def test_host_compiler_matches_reference_counts(name):
    gold = json.load(open(os.path.join(GOLDEN, f"ref_{name}.json")))
    d = pm.Dictionary()
    for p in dict_paths(name):
        d.add_file(p)
    d.compile()
    i = d.info
    assert i.n_patterns == gold["n_patterns"] and i.n_ac_states == gold["n_states"] and i.max_pat_len == gold["max_pat_len"]
    assert (i.n_lines, i.n_rejected, i.n_duplicates) == (gold["n_lines"], gold["n_rejected"], gold["n_duplicates"])


def test_host_compiler_patterns_and_parents_equal_oracle(dict_merged, oracle_merged):
    P = dict_merged.n_patterns
    assert P == oracle_merged.n_patterns
    of, ol = oracle_merged.id_arrays()
    par = oracle_merged.parents()
    rng = np.random.default_rng(1)
    for pid in list(range(1, 400)) + rng.integers(1, P + 1, 3000).tolist() + [P]:
        f, l, _, parent, b = dict_merged.pattern(pid)
        o = oracle_merged.pattern(pid - 1)
        assert (f, l, b) == (o[0], o[1], o[3])
        assert parent == par[pid - 1] + 1
    files, lines = dict_merged.id_arrays()
    assert np.array_equal(files[1:], of) and np.array_equal(lines[1:], ol)
    # is_pattern_suffix on pids == PatternsTree.c:485-494 on the oracle
    for a, b in rng.integers(1, P + 1, (2000, 2)).tolist():
        assert dict_merged.is_pattern_suffix(a, b) == bool(oracle_merged.L.pmo_is_pattern_suffix(oracle_merged.h, a - 1, b - 1))
    deep = int(np.argmax([0] + [dict_merged.pattern(p)[3] > 0 for p in range(1, 2000)]))
    assert dict_merged.is_pattern_suffix(dict_merged.pattern(deep)[3], deep)


def test_parse_line_same_as_oracle():
    cases = [b"abc", b"|41 42|CD", b"|41 |", b"|4|", b"|41", b"", b"||", b"a||b", b"| 2E 65 6D 66 |", b"|0a0D|", b"x|7C|y",
             b"|30 14 06 03 55 04 03 14 0D 2A|.dropbox.com", b"T|00|e|00|", b"|  41|", b"|4 1|", b"|zz|"]
    for c in cases:
        assert pm.parse_pattern_line(c) == parse_line(c), c


def test_add_pattern_dedup_returns_first_pid():
    d = pm.Dictionary()
    assert d.add_pattern(b"abc", 0, 1, 111) == 1
    assert d.add_pattern(b"xyz", 0, 2, 222) == 2
    assert d.add_pattern(b"abc", 0, 3, 333) == 1          # PatternsTree.c:193-196
    assert d.add_pattern(b"a\x00b", 0, 4, 444) == 3       # binary patterns with NUL bytes are kept whole
    d.compile()
    assert d.n_patterns == 3 and d.pattern(1)[2] == 111 and d.pattern(3)[4] == b"a\x00b"


def test_tiny_dictionary_shape():
    d = pm.Dictionary().add_bytes(TINY_DICT).compile()
    o = Oracle(); o.add_dict_bytes(TINY_DICT); o.compile()
    assert d.n_patterns == 11 and d.info.n_rejected == 2 and d.info.n_ac_states == o.n_states
    assert d.info.n_classes < 256                           # alphabet compression: unused bytes share class 0


def test_limits_are_reported_not_silently_truncated():
    """The reference's Aho-Corasick takes any dictionary (mpac.c:257-291).  Here: patterns of any length and any number of
    2-byte continuations compile (the engine serves dictionaries without backward-scan tables with the forward walkers);
    more than 65,535 unique patterns -- which dense uint16 results cannot name -- compile into parts that an engine scans
    one after the other (32-bit results; tests/test_gpu_parity.py::test_more_than_65535_patterns)."""
    d = pm.Dictionary()
    d.add_pattern(b"x" * 354, 0, 1)                                      # beyond the round-1 limit of 353 bytes
    d.add_pattern(b"y" * 5000, 0, 2)                                     # beyond what the backward scan's queue items encode
    d.compile()
    assert d.max_pat_len == 5000 and d.n_patterns == 2
    rng = np.random.default_rng(9)
    d = pm.Dictionary()                                                  # P + 2-byte continue codes > 65,535
    for i in range(60000):
        d.add_pattern(bytes(rng.integers(0, 256, 4, dtype=np.uint8)), 0, i + 1)
    d.compile()
    assert d.n_patterns + d.info.n_hot2_cont > 65535
    d = pm.Dictionary()                                                  # more patterns than dense uint16 results can name:
    for i in range(66000):                                               # compiled into parts (32-bit results, dict.hpp)
        d.add_pattern(b"%07d" % i, 0, i + 1)
    d.add_pattern(b"0001", 0, 70001)                                     # a suffix of b"0000001" and of nothing else
    d.compile()
    assert d.n_patterns == 66001 and d.max_pat_len == 7
    assert d.pattern(2)[3] == 66001 and d.pattern(66001)[3] == 0        # the PatternsTree relation spans the parts


def test_compiled_dictionary_cache_round_trip(tmp_path, dict_merged):
    """the compiled tables survive a save/load round trip and the content-keyed cache is reused."""
    import time
    f = tmp_path / "merged.bin"
    dict_merged.save(str(f))
    d2 = pm.Dictionary.load(str(f))
    a, b = dict_merged.info, d2.info
    for field, _ in pm.DictInfo._fields_:
        assert getattr(a, field) == getattr(b, field), field
    for pid in (1, 2, 777, 40000, a.n_patterns):
        assert dict_merged.pattern(pid) == d2.pattern(pid)
    paths = dict_paths("merged")
    t0 = time.time(); d3 = pm.Dictionary.from_files_cached(paths, str(tmp_path)); t_cold = time.time() - t0
    t0 = time.time(); d4 = pm.Dictionary.from_files_cached(paths, str(tmp_path)); t_warm = time.time() - t0
    assert d3.info.n_patterns == d4.info.n_patterns == a.n_patterns and d4.info.n_sfx_rows == a.n_sfx_rows
    assert len([x for x in os.listdir(tmp_path) if x.startswith("pmdict-")]) == 1
    assert d4.pattern(31337) == dict_merged.pattern(31337)
    # a different file order is a different dictionary (file numbers differ): different key
    pm.Dictionary.from_files_cached(paths[::-1], str(tmp_path))
    assert len([x for x in os.listdir(tmp_path) if x.startswith("pmdict-")]) == 2
    with pytest.raises(pm.PmError):
        bad = tmp_path / "bad.bin"; bad.write_bytes(b"not a dictionary"); pm.Dictionary.load(str(bad))
    print(f"cold {t_cold:.2f}s warm {t_warm:.2f}s")


def test_plugin_registration_fills_an_mpselem():
    """mps_gpu_register_into fills a struct laid out like MpsElem (Core/src/mps.h:71-80) with the seven callbacks."""
    L = pm.lib()
    for fn, name in ((L.mps_gpu_register_into, b"B200 exact dictionary scan"),
                     (L.mps_gpu_dfa_register_into, b"B200 Aho-Corasick DFA"),
                     (L.mps_gpu_kr_register_into, b"B200 Karp-Rabin stages")):
        e = pm.MpsElemStruct()
        fn(C.byref(e))
        assert e.name == name
        for field in ("create", "add_pattern", "compile", "read_char", "total_mem", "reset", "free"):
            assert C.cast(getattr(e, field), C.c_void_p).value, field
    # create / add_pattern work without a GPU (compile needs the device)
    e = pm.MpsElemStruct(); L.mps_gpu_register_into(C.byref(e))
    obj = e.create()
    e.add_pattern(obj, b"a\x00b", 3, 0x1234)
    assert e.total_mem(obj) == 0
    e.free(obj)


def emulate_deep_walk(d, stream):
    """Host emulation of deep_scan.cu's per-lane rounds over the compact goto + failure records (dict.hpp: DeepTables):
    per round at most one table load (a record, or a DENSE-row entry), the report of the byte consumed in the round before,
    and one transition.  The table builder is host logic and is checked here without a GPU."""
    recs = d.table("deep.recs", np.uint32).reshape(-1, 8)
    hot = d.table("deep.hot_rows", np.uint16).reshape(-1, 256)
    small_long = d.table("deep.hot_longest", np.uint16)
    dense = d.table("deep.dense_rows", np.uint32).reshape(-1, 256)
    n_hot, n_small = hot.shape[0], small_long.size
    dense_end = n_hot + dense.shape[0]
    out = np.zeros(stream.size, np.uint16)
    n = stream.size
    s, rel, orel, fetches, rounds = 0, 0, 0, 0, 0
    have = head = pend = pend_cold = False
    pend_val = 0
    w = None
    while rel < n or pend:
        rounds += 1
        active = rel < n
        c = int(stream[rel]) if active else 0
        is_hot = s < n_hot
        is_dense = (not is_hot) and s < dense_end
        if (not is_hot) and (not is_dense) and not have:
            w = [int(x) for x in recs[s]]; fetches += 1
            have = head = True
        if pend:
            out[orel] = w[1] if pend_cold else pend_val
            orel += 1; pend = False
        if not active:
            continue
        known, val = False, 0
        if is_hot:
            ns, consumed, have = int(hot[s, c]), True, False
        elif is_dense:
            ns, consumed, have = int(dense[s - n_hot, c]), True, False; fetches += 1
        else:
            kind, cnt, fail = (w[0] >> 24) & 3, w[0] >> 26, w[0] & 0xFFFFFF
            if kind == 1:
                if (w[2] & 0xFF) == c:
                    val, known, ns, consumed = w[4] & 0xFFFF, True, s + 1, True
                    labels = (w[2] | (w[3] << 32)) >> 8
                    longs = (w[4] | (w[5] << 32) | (w[6] << 64) | (w[7] << 96)) >> 16
                    w[2], w[3] = labels & 0xFFFFFFFF, labels >> 32
                    w[4], w[5], w[6], w[7] = (longs & 0xFFFFFFFF, (longs >> 32) & 0xFFFFFFFF, (longs >> 64) & 0xFFFFFFFF, longs >> 96)
                    w[0] -= 1 << 26
                    head = False
                    have = (w[0] >> 26) != 0
                else:
                    ns, consumed, have = (fail if head else s), False, False
            else:
                nxt = None
                for k in range(6):
                    if (w[2 + k] & 0xFF) == c:
                        nxt = w[2 + k] >> 8
                consumed = kind == 0 and nxt is not None
                ns, have = (nxt if consumed else fail), False
        if consumed:
            if (not known) and ns < n_small:
                val, known = int(small_long[ns]), True
            pend, pend_cold, pend_val = True, not known, val
            rel += 1
        s = ns
    return out, fetches, rounds


def test_deep_automaton_records_walk_like_the_oracle(dict_merged, oracle_merged):
    for kind, n in (("almost", 60000), ("planted", 30000), ("ascii", 20000)):
        stream = oracle_merged.gen(kind, 4096 * 3, n)
        got, fetches, rounds = emulate_deep_walk(dict_merged, stream)
        want = (oracle_merged.scan(stream) + 1).astype(np.uint16)
        assert np.array_equal(got, want), kind
        print(f"deep records on {kind}: {fetches / n:.3f} table loads and {rounds / n:.3f} rounds per byte")
    # a small-alphabet dictionary: long chains, heavy failure traffic
    pats = [b"a" * k for k in range(1, 41)] + [b"ab" * k + b"c" for k in range(1, 20)] + [b"bca", b"cab", b"abcabcabd"]
    d = pm.Dictionary(); o = Oracle()
    for i, p in enumerate(pats):
        d.add_pattern(p, 0, i + 1); o.add_pattern(p, 0, i + 1)
    d.compile(); o.compile()
    rng = np.random.default_rng(2)
    stream = rng.choice(np.frombuffer(b"aaabbc", np.uint8), 50000)
    got, _, _ = emulate_deep_walk(d, stream)
    assert np.array_equal(got, (o.scan(stream) + 1).astype(np.uint16))


def test_host_thread_pool_stress(tmp_path):
    """The staging pool of the host pipeline (host_pool.hpp: asynchronous jobs, caller participation, polling workers) is
    plain C++ without CUDA: built here with g++ and driven with random job mixes -- several jobs in flight, waited out of
    order, empty jobs, workers that fell asleep -- plus the pid -> id translation loop against a scalar loop."""
    import subprocess
    exe = str(tmp_path / "pool_stress")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "patternmatching_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host_pool_stress.cpp"), "-o", exe])
    for threads in ("1", "3", "8"):
        out = subprocess.run([exe, threads], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr
