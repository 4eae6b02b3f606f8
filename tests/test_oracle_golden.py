"""CPU tests: the oracle restatement (oracle/pm_oracle.c) against the golden fixtures generated from
the reference itself (tests/golden/ref_*.json, scripts/make_golden.py) and against the known-answer
cases the reference's sources contain."""
import json
import os

import numpy as np
import pytest

from conftest import DATA, GOLDEN, TINY_DICT, TINY_STREAM, dict_paths
from oracle_lib import Oracle, lib, parse_line


@pytest.mark.parametrize("name", ["snort", "et", "merged"])
def test_oracle_matches_reference_golden(name):
    gold = json.load(open(os.path.join(GOLDEN, f"ref_{name}.json")))
    o = Oracle()
    for p in dict_paths(name):
        o.add_dict_file(p)
    o.compile()
    assert o.n_patterns == gold["n_patterns"]
    assert o.n_states == gold["n_states"]                     # reference: (ac_total_mem - 24) / 2072
    assert o.max_pat_len == gold["max_pat_len"]
    assert (o.n_lines, o.n_rejected, o.n_duplicates) == (gold["n_lines"], gold["n_rejected"], gold["n_duplicates"])
    assert 24 + 2072 * o.n_states == gold["ac_total_mem"]     # results.csv:2 for merged: 1,485,093,592
    stream = np.fromfile(os.path.join(DATA, gold["stream"]), dtype=np.uint8)
    longest = o.scan(stream)
    files, lines = o.id_arrays()
    f = np.where(longest >= 0, files[np.maximum(longest, 0)].astype(np.int64), -1)
    l = np.where(longest >= 0, lines[np.maximum(longest, 0)].astype(np.int64), -1)
    assert np.array_equal(f, np.array(gold["longest_file"])) and np.array_equal(l, np.array(gold["longest_line"]))
    s = o.summary(stream)
    assert (s.positions, s.matches, "%016x" % s.fnv) == (gold["positions"], gold["matches"], gold["fnv"])
    assert ("%016x" % s.hsum_longest, "%016x" % s.hsum_all) == (gold["hsum_longest"], gold["hsum_all"])
    if name == "merged":
        assert gold["ac_total_mem"] == 1485093592 and gold["lmac_total_mem"] == 40137664    # results.csv:2-3
    assert gold["lmac_vs_ac_counts"] == [10240, 0, 0, 0]      # LMAC == AC, results.csv:3
    # the reference MPBG only ever reports patterns of <= 8 bytes: the rule reproduces its rates
    if gold.get("mpbg_vs_ac_counts"):
        lens = o.lengths(); par = o.parents()
        demoted = longest.copy()
        for i in np.nonzero(longest >= 0)[0]:
            q = int(longest[i])
            while q >= 0 and lens[q] > 8:
                q = int(par[q])
            demoted[i] = q
        c = o.classify(demoted, longest)
        assert [c["success"], c["partial"], c["false_neg"], c["false_pos"]] == gold["mpbg_vs_ac_counts"]
        # ... and, position for position, the reference MPBG's own output (pmref_scan over mpbg_read_char, mpbg.c:132-145)
        mf = np.where(demoted >= 0, files[np.maximum(demoted, 0)].astype(np.int64), -1)
        ml = np.where(demoted >= 0, lines[np.maximum(demoted, 0)].astype(np.int64), -1)
        assert np.array_equal(mf, np.array(gold["mpbg_file"])) and np.array_equal(ml, np.array(gold["mpbg_line"]))


def test_parser_quirk_q1():
    """Core/src/parser.c:63-99: hex sections, rejection of a space before the closing bar, odd nibbles,
    unterminated sections; raw bytes otherwise."""
    assert parse_line(b"abc") == b"abc"
    assert parse_line(b"|41 42|CD") == b"ABCD"
    assert parse_line(b"|2829|") == b"\x28\x29"               # pairs need no separator
    assert parse_line(b"|0a 0D|x|ff|") == b"\x0a\x0dx\xff"    # upper / lower case digits
    assert parse_line(b"| 50 4B 03 04|") == b"PK\x03\x04"     # spaces before a nibble are skipped
    assert parse_line(b"|4 1|") == b"A"                       # ... also between the nibbles
    assert parse_line(b"|41 |") is None                       # space directly before the closing bar
    assert parse_line(b"| 3C |") is None
    assert parse_line(b"|4|") is None                         # odd nibble count
    assert parse_line(b"|41") is None                         # unterminated
    assert parse_line(b"|4G|") is None                        # not hex
    assert parse_line(b"") is None and parse_line(b"||") is None   # empty result is skipped
    assert parse_line(b"a||b") == b"ab"
    assert parse_line(b"530 ") == b"530 "                     # trailing blanks are pattern bytes


def test_dedup_and_ids_quirk_q2():
    """First occurrence wins; line numbers count every line incl. rejected / empty ones; file = -d index."""
    o = Oracle()
    o.add_dict_bytes(b"abc\n\n|41 |\nabc\nxyz\n")            # line 2 empty, line 3 rejected, line 4 duplicate
    o.add_dict_bytes(b"xyz\nnew\n")
    o.compile()
    got = [o.pattern(i)[:2] + (o.pattern(i)[3],) for i in range(o.n_patterns)]
    assert got == [(0, 1, b"abc"), (0, 5, b"xyz"), (1, 2, b"new")]
    assert o.n_duplicates == 2 and o.n_lines == 7


def test_readme_suffix_tree_example():
    """Core/src/README.md:16-47: {abcdefg, cdefg, efg, afg, fg} -> root->fg->{efg->cdefg->abcdefg, afg}."""
    o = Oracle()
    o.add_dict_bytes(b"abcdefg\ncdefg\nefg\nafg\nfg\n")
    o.compile()
    idx = {o.pattern(i)[3]: i for i in range(5)}
    par = {o.pattern(i)[3]: o.pattern(i)[2] for i in range(5)}
    assert par[b"fg"] == -1
    assert par[b"efg"] == idx[b"fg"] and par[b"afg"] == idx[b"fg"]
    assert par[b"cdefg"] == idx[b"efg"] and par[b"abcdefg"] == idx[b"cdefg"]
    L = lib()
    assert L.pmo_is_pattern_suffix(o.h, idx[b"fg"], idx[b"abcdefg"]) == 1
    assert L.pmo_is_pattern_suffix(o.h, idx[b"afg"], idx[b"abcdefg"]) == 0
    assert L.pmo_is_pattern_suffix(o.h, -1, idx[b"fg"]) == 0


def test_appendix_a_tiny_selfcheck():
    """11 unique patterns, 7 positions with a match, 12 matches."""
    o = Oracle(); o.add_dict_bytes(TINY_DICT); o.compile()
    assert o.n_patterns == 11
    stream = np.frombuffer(TINY_STREAM, np.uint8)
    longest = o.scan(stream)
    lines = o.id_arrays()[1]
    par = o.parents()
    got = {}
    for i in np.nonzero(longest >= 0)[0]:
        q, chain = int(longest[i]), []
        while q >= 0:
            chain.append(int(lines[q])); q = int(par[q])
        got[int(i)] = chain
    assert got == {3: [7, 6], 5: [9], 13: [1, 2, 3, 5], 18: [4, 5], 23: [10], 28: [10], 37: [13]}
    s = o.summary(stream)
    assert (s.positions, s.matches) == (7, 12)


def test_kmp_known_answer():
    """Core/src/kmprt.c:303-327: the commented-out main's pattern / text report matches ending at 17 and 42."""
    L = lib()
    pat = np.frombuffer(b"AAAAAAAAAAAAAAAAAB", np.uint8)
    text = np.frombuffer(b"AAAAAAAAAAAAAAAAABAAAAAABAAAAAAAAAAAAAAAAABAAAAAAA", np.uint8)
    ends = np.zeros(8, np.uint64)
    n = L.pmo_kmp_search(pat.ctypes.data, pat.size, text.ctypes.data, text.size, ends.ctypes.data, 8)
    assert n == 2 and ends[:2].tolist() == [17, 42]


def test_bg_intended_answer_with_our_kr_variant():
    """Core/src/bgps.c:624-655: pattern ABCDABDABC in the example text must be reported at 13, 20, 34 (the
    shipped reference prints nothing; our seeded variant does)."""
    o = Oracle(); o.add_dict_bytes(b"ABCDABDABC\n"); o.compile()
    text = np.frombuffer(b"ABCDABCDABDABCDABDABCDABBABCDABDABCDABDBADFSG", np.uint8)
    for seed in (1, 2, 0xF1A90003):
        got = o.kr_scan(text, seed)
        assert np.nonzero(got >= 0)[0].tolist() == [13, 20, 34]
    assert np.nonzero(o.scan(text) >= 0)[0].tolist() == [13, 20, 34]


def test_field_and_fingerprint_identities():
    """Fingerprint.h:66-89 identities in GF(2^31-1): prefix + suffix * r^|prefix| = whole; inverse from field.c."""
    L = lib()
    p = 2147483647
    rng = np.random.default_rng(3)
    for _ in range(20):
        r = int(L.pmo_kr_seed_r(int(rng.integers(1, 1 << 60))))
        assert 1 <= r < p
        assert L.pmo_mulmod(r, L.pmo_invmod(r)) == 1
        s = rng.integers(0, 256, 40, dtype=np.uint8)
        k = int(rng.integers(1, 39))
        whole = L.pmo_fp(s.ctypes.data, 40, r)
        pre = L.pmo_fp(s.ctypes.data, k, r)
        suf = L.pmo_fp(s[k:].copy().ctypes.data, 40 - k, r)
        assert (pre + L.pmo_mulmod(suf, L.pmo_powmod(r, k))) % p == whole
        assert L.pmo_mulmod((whole - pre) % p, L.pmo_invmod(L.pmo_powmod(r, k))) == suf


def test_fingerprints_and_inverses_equal_the_reference():
    """The Karp-Rabin arithmetic pinned against the reference's own code (compiled unchanged into oracle/_ref/libpmref.so):
    calc_fp / calc_fp_with_prefix (Core/src/Fingerprint.c:29-42, 57-77) and calculate_inverse (Core/src/field.c:26-72),
    p = 2^31-1 (Core/src/mpbg.c:83).  Bytes < 0x80 only: for larger bytes the reference's pattern side sign-extends
    `char` and lets the sum wrap, which the unsigned restatement deliberately does not reproduce -- the
    last block shows the two really differ there."""
    import ctypes as C
    import os
    from conftest import ROOT
    path = os.path.join(ROOT, "oracle", "_ref", "libpmref.so")
    if not os.path.exists(path):
        pytest.fail("oracle/_ref/libpmref.so missing: run `make -C oracle` in the build container")
    R = C.CDLL(path)

    class FieldVal(C.Structure):
        _fields_ = [("val", C.c_ulonglong), ("inv", C.c_ulonglong)]

    R.calc_fp.restype = C.c_ulonglong
    R.calc_fp.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(FieldVal), C.POINTER(FieldVal), C.c_ulonglong]
    R.calc_fp_with_prefix.restype = C.c_ulonglong
    R.calc_fp_with_prefix.argtypes = [C.c_char_p, C.c_size_t, C.c_ulonglong, C.c_size_t, C.POINTER(FieldVal), C.POINTER(FieldVal), C.c_ulonglong]
    R.calculate_inverse.restype = C.c_ulonglong
    R.calculate_inverse.argtypes = [C.c_ulonglong, C.c_ulonglong]
    L = lib()
    p = 2147483647
    rng = np.random.default_rng(11)
    for a in [1, 2, 3, 7, p - 1, p - 2, 65537, 1 << 30] + rng.integers(1, p, 200).tolist():
        inv = R.calculate_inverse(a, p)
        assert inv == L.pmo_invmod(a) and (inv * a) % p == 1
    for trial in range(200):
        r = int(L.pmo_kr_seed_r(int(rng.integers(1, 1 << 60))))
        rv = FieldVal(r, R.calculate_inverse(r, p))
        n = int(rng.integers(1, 348))
        s = rng.integers(0, 128, n, dtype=np.uint8)               # the reference's sign-extension does not bite below 0x80
        rn = FieldVal()
        want = R.calc_fp(s.tobytes(), n, C.byref(rn), C.byref(rv), p)
        assert want == L.pmo_fp(s.ctypes.data, n, r)
        assert rn.val == L.pmo_powmod(r, n) and rn.inv == L.pmo_invmod(L.pmo_powmod(r, n))
        # extending a prefix fingerprint: Fingerprint.c:57-77 == our identity fp(whole) = fp(pre) + r^k fp(suf)
        k = int(rng.integers(0, n))
        rk = FieldVal()
        pre = R.calc_fp(s[:k].tobytes(), k, C.byref(rk), C.byref(rv), p) if k else 0
        if not k:
            rk = FieldVal(1, 1)
        whole = R.calc_fp_with_prefix(s.tobytes(), n, pre, k, C.byref(rk), C.byref(rv), p)   # takes the WHOLE sequence and the prefix length
        assert whole == want
    # Q6: with bytes >= 0x80 the reference's own two sides disagree with the unsigned definition in a measurable share
    differ = 0
    for trial in range(300):
        r = int(L.pmo_kr_seed_r(1000 + trial))
        rv = FieldVal(r, R.calculate_inverse(r, p)); rn = FieldVal()
        s = rng.integers(128, 256, 12, dtype=np.uint8)
        differ += R.calc_fp(s.tobytes(), 12, C.byref(rn), C.byref(rv), p) != L.pmo_fp(s.ctypes.data, 12, r)
    assert differ > 0


def test_sharded_scan_with_halo_equals_continuous(oracle_merged):
    """Quirk Q8 on the CPU: re-scanning shards from reset with a max_pat_len-1 halo reproduces the scan."""
    o = oracle_merged
    stream = o.gen("almost", 0, 1 << 18)
    full = o.summary(stream)
    halo = o.max_pat_len - 1
    acc = dict(positions=0, matches=0, h0=0, h1=0)
    k = 7
    for w in range(k):
        lo, hi = stream.size * w // k, stream.size * (w + 1) // k
        start = max(lo - halo, 0)
        s = o.summary(stream[start:hi], skip=lo - start, pos_base=lo)
        acc["positions"] += s.positions; acc["matches"] += s.matches
        acc["h0"] = (acc["h0"] + s.hsum_longest) % (1 << 64); acc["h1"] = (acc["h1"] + s.hsum_all) % (1 << 64)
    assert (acc["positions"], acc["matches"], acc["h0"], acc["h1"]) == (full.positions, full.matches, full.hsum_longest, full.hsum_all)


def test_generators_are_offset_pure(oracle_merged):
    for kind in ("uniform", "planted", "almost", "ab"):
        whole = oracle_merged.gen(kind, 4096 * 3, 4096 * 5)
        for off, n in ((4096 * 3 + 17, 1000), (4096 * 4 - 5, 4106), (4096 * 7, 4096)):
            part = oracle_merged.gen(kind, off, n)
            assert np.array_equal(part, whole[off - 4096 * 3: off - 4096 * 3 + n]), kind
