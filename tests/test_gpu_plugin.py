"""GPU tests of the drop-in boundary: the reference's OWN program with the B200 matcher registered in its
mps_table (oracle/_ref/exe_gpu, built by oracle/make_gpu_exe.py from a patched scratch copy of Core/src), the
measurement driver hosting the reference's CPU algorithms beside the GPU rows, and the plugin surface fed with the
full merged dictionary in the reference's add order."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import patternmatching_b200 as pm
from conftest import DATA, ROOT, dict_paths
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def read_csv(path):
    rows = [l.split(",") for l in open(path).read().splitlines()]
    return rows[0], {r[0]: r for r in rows[1:]}


def rates(row):
    return [float(x) for x in row[3:6]]          # false positive, false negative, partial success


def test_reference_program_runs_with_the_gpu_matcher_inside(tmp_path):
    """main.c / measure.c / mps.c of the reference, patched exactly as INTEGRATION.md says, linked against libpm_b200.so:
    `exe -d snort.dict -d et.dict -s dictionaries_generated.stream -o out.csv`.  The reference classifies every row
    against ITS reliable Aho-Corasick (mps.c:52-53, measure.c:174-190): the GPU exact row must score 0/0/0 like AC and
    LMAC (results.csv:2-3), the reference's MPBG row must reproduce results.csv:4."""
    exe = os.path.join(REF_DIR, "exe_gpu")
    if not os.path.exists(exe):
        pytest.fail("oracle/_ref/exe_gpu missing: run `make -C oracle` in the build container")
    out = tmp_path / "out.csv"
    out.write_text("")                                # the reference opens with O_CREAT and no mode
    args = [exe, "-v", "-o", str(out), "-s", os.path.join(DATA, "dictionaries_generated.stream")]
    for p in dict_paths("merged"):
        args += ["-d", p]
    r = subprocess.run(args, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    head, rows = read_csv(out)
    assert head[:6] == ["Algorithm", "Time (in secs)", "Total Memory Used", "False Positive Rate", "False Negative Rate",
                        "Partial Success Rate"]
    assert len(rows) == 5, list(rows)
    gpu, kr = rows["B200 exact dictionary scan"], rows["B200 Karp-Rabin stages"]
    assert rates(gpu) == [0.0, 0.0, 0.0] and int(gpu[2]) > 0
    others = [r for name, r in rows.items() if not name.startswith("B200")]
    exact_cpu = [r for r in others if rates(r) == [0.0, 0.0, 0.0]]
    assert len(exact_cpu) == 2 and int(exact_cpu[0][2]) in (1485093592, 40137664)       # AC and LMAC, results.csv:2-3
    mpbg = [r for r in others if rates(r) != [0.0, 0.0, 0.0]][0]
    assert rates(mpbg) == [0.0, 0.000293, 0.022754] and int(mpbg[2]) == 26423230         # results.csv:4
    fp, fn, part = rates(kr)
    assert fn == 0.0 and part == 0.0 and fp <= 1e-3
    print("reference exe with GPU rows:", {k: (v[1], rates(v)) for k, v in rows.items()})


def test_driver_hosts_reference_cpu_algorithms_beside_gpu_rows(tmp_path):
    """pm_driver -p: every entry of the reference's mps_table (AC, MPBG, LMAC compiled unchanged into
    oracle/_ref/libpmref.so) is driven through its seven MpsElem callbacks with one read_char per byte, the GPU matchers
    through the same struct plus read_block; all rows are classified against a second instance of the reference's AC."""
    exe = os.path.join(ROOT, "patternmatching_b200", "pm_driver")
    out = tmp_path / "results.csv"
    args = [exe, "-v", "-o", str(out), "-p", os.path.join(REF_DIR, "libpmref.so"),
            "-s", os.path.join(DATA, "dictionaries_generated.stream"), "-d", dict_paths("snort")[0]]
    r = subprocess.run(args, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    head, rows = read_csv(out)
    assert len(rows) == 7
    gpu_rows = [v for k, v in rows.items() if k.startswith("B200")]
    cpu_rows = [v for k, v in rows.items() if not k.startswith("B200")]
    assert len(gpu_rows) == 4 and len(cpu_rows) == 3
    for v in gpu_rows[:2]:
        assert rates(v) == [0.0, 0.0, 0.0] and int(v[6]) == 10240
    exact_cpu = [v for v in cpu_rows if rates(v) == [0.0, 0.0, 0.0]]
    assert len(exact_cpu) == 2 and 511307 * 2072 + 24 in [int(v[2]) for v in exact_cpu]      # AC and LMAC; 511,307 states
    mpbg = [v for v in cpu_rows if rates(v) != [0.0, 0.0, 0.0]][0]
    assert rates(mpbg) == [0.0, 0.000391, 0.018262]            # snort only: 4 FN + 187 partial of 10240
    assert rates(rows["B200 MPBG (as shipped)"]) == rates(mpbg)  # PM_ALGO_MPBG: the same rates as the reference's MPBG row
    # -r gpu: the same run classified against the B200 DFA instead; needs no plugin
    out2 = tmp_path / "r2.csv"
    r = subprocess.run([exe, "-o", str(out2), "-r", "gpu", "-b", "102400", "-s", os.path.join(DATA, "dictionaries_generated.stream"),
                        "-d", dict_paths("snort")[0]], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    _, rows2 = read_csv(out2)
    assert len(rows2) == 4 and all(rates(v)[1] == 0.0 for k, v in rows2.items() if "MPBG" not in k)


@pytest.fixture(scope="module")
def reference_patterns():
    """The merged dictionary's unique patterns in the reference's add_pattern order, from the reference itself."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from reflib import Reference
    ref = Reference(dict_paths("merged"), algo_mask=0)        # PatternsTree only: no matcher is built
    return list(ref.patterns())


def test_plugin_with_the_merged_dictionary_in_reference_add_order(reference_patterns, oracle_merged):
    """55,580 patterns through gpu_add_pattern in the order and with the kind of ids the reference passes
    (post-order of its PatternsTree, PatternsTree.c:390-401; ids are distinct non-null pointers), then read_block /
    read_char in every buffer regime: pageable multi-piece, 100 KiB chunks (measure.c:77), single bytes."""
    pats = reference_patterns
    assert len(pats) == oracle_merged.n_patterns
    of, ol = oracle_merged.id_arrays()
    canon = {(int(f), int(l)): i for i, (f, l) in enumerate(zip(of, ol))}
    id_of_canon = np.zeros(len(pats) + 1, np.uint64)          # oracle index + 1 -> the id given to add_pattern
    m = pm.MpsGpu("sfx")
    for k, (f, l, _, _, b) in enumerate(pats):
        ident = 0x7F0000001000 + 48 * k
        id_of_canon[canon[(f, l)] + 1] = ident
        m.add_pattern(b, ident)
    m.compile()
    n = (9 << 20) + 4321
    stream = np.concatenate([oracle_merged.gen("planted", 0, 6 << 20), oracle_merged.gen("almost", 4096, 4 << 20)])[:n]
    want = id_of_canon[oracle_merged.scan(stream) + 1]
    m.reset()
    assert np.array_equal(m.read_block(stream), want)                       # pageable, several 4 MiB pieces
    m.reset()
    got = np.concatenate([m.read_block(stream[o:o + 102400]) for o in range(0, 3 << 20, 102400)])
    assert np.array_equal(got, want[:got.size])                             # the reference's chunking, state carried
    rest = np.array([m.read_char(int(c)) for c in stream[got.size:got.size + 300]], np.uint64)
    assert np.array_equal(rest, want[got.size:got.size + 300])              # read_char continues the same stream
    hin = pm.PinnedBuffer(n); hout = pm.PinnedBuffer(8 * n)                 # page-locked buffers
    hin.array(np.uint8)[:] = stream
    m.reset()
    m.read_block_ptr(hin.ptr, n, hout.ptr)
    assert np.array_equal(hout.array(np.uint64)[:n], want)
    m.free()


def test_scan_host_buffer_regimes_agree(oracle_merged, engine_merged):
    """pm_engine_scan_host: pageable vs page-locked, small-call path vs pipeline, odd sizes around the thresholds."""
    rng = np.random.default_rng(3)
    total = oracle_merged.gen("planted", 8192, 13 << 20)
    want = (oracle_merged.scan(total) + 1).astype(np.uint16)
    for n in (1, 100, 102400, (256 << 10), (256 << 10) + 1, (4 << 20) + 5, total.size):
        engine_merged.reset()
        assert np.array_equal(engine_merged.scan_host(total[:n]), want[:n]), n
    # mixed call sizes on one stream: the carried history is the only state
    engine_merged.reset()
    cuts = [0, 5, 4096, 150000, 150001, 700000, 5 << 20, total.size]
    got = np.concatenate([engine_merged.scan_host(total[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
    assert np.array_equal(got, want)
    hin = pm.PinnedBuffer(total.size); hout = pm.PinnedBuffer(2 * total.size)
    hin.array(np.uint8)[:] = total
    engine_merged.reset()
    engine_merged.scan_host_ptr(hin.ptr, total.size, hout.ptr)
    assert np.array_equal(hout.array(np.uint16)[:total.size], want)
    table = rng.integers(1, 1 << 62, oracle_merged.n_patterns + 1, dtype=np.uint64)
    table[0] = 0
    engine_merged.reset()
    assert np.array_equal(engine_merged.scan_host_ids(total, table), table[want])
    engine_merged.reset()
    assert engine_merged.host_threads >= 1 and engine_merged.scratch_mem > 0


def test_device_scans_on_alternating_streams_are_ordered(oracle_merged, engine_merged):
    """pm_engine_scan_device shares one set of scan scratch per engine; successive calls on DIFFERENT CUDA streams
    without any host synchronisation in between must still each produce the exact result."""
    import torch
    dev = torch.device("cuda:0")
    n = 2 << 20
    streams = [oracle_merged.gen(kind, 4096 * 3, n) for kind in ("planted", "almost", "ascii", "planted")]
    wants = [(oracle_merged.scan(s) + 1).astype(np.uint16) for s in streams]
    d_in = [torch.from_numpy(s).to(dev) for s in streams]
    d_out = [torch.zeros(n, dtype=torch.int16, device=dev) for _ in streams]
    cs = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(3):
        for k in range(4):
            engine_merged.scan_device(d_in[k], n, d_out[k], algo=pm.ALGO_SFX, cuda_stream=cs[k & 1].cuda_stream)
    torch.cuda.synchronize()
    for k in range(4):
        assert np.array_equal(d_out[k].cpu().numpy().view(np.uint16), wants[k]), k


def test_engines_on_two_threads_share_one_dictionary(oracle_merged, dict_merged):
    """Two engines created on two threads from one compiled dictionary, both asking for the lazily built DFA tables and
    different KR seeds at the same time (the dictionary builds the DFA once under its own lock; KR tables are per engine)."""
    d = pm.Dictionary()
    for p in dict_paths("snort"):
        d.add_file(p)
    d.compile()
    o = Oracle()
    o.add_dict_file(dict_paths("snort")[0]); o.compile()
    stream = o.gen("planted", 0, 1 << 20)
    want = (o.scan(stream) + 1).astype(np.uint16)
    results, errors = {}, []

    def work(i):
        try:
            eng = pm.Engine(d)
            eng.set_kr_seed(1000 + i)
            a = eng.scan_host(stream, algo=pm.ALGO_DFA)
            eng.reset()
            b = eng.scan_host(stream, algo=pm.ALGO_KR)
            results[i] = (a, b)
        except Exception as ex:          # surfaced below
            errors.append(ex)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errors, errors
    for i in range(2):
        assert np.array_equal(results[i][0], want)
        assert np.array_equal(results[i][1], (o.kr_scan(stream, 1000 + i) + 1).astype(np.uint16))
