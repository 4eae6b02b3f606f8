"""ctypes wrapper around oracle/_ref/libpmref.so -- the UNMODIFIED reference sources driven by
oracle/ref_harness.c.  TEST / BASELINE INFRASTRUCTURE ONLY: used by tests/ and by bench.py's
cpu_baseline leg and `--impl reference` arm, never by the product path.  One build per process (the
reference keeps global state and is not re-entrant)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "_ref", "data")
AC, LMAC, MPBG = 0, 1, 2


class Reference:
    def __init__(self, dict_paths, algo_mask=0b001, opt="O2"):
        name = "libpmref.so" if opt == "O2" else "libpmref_O0.so"
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle` where /root/reference exists")
        L = C.CDLL(path)
        L.pmref_build.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int]
        L.pmref_n_patterns.restype = C.c_size_t
        L.pmref_max_pat_len.restype = C.c_size_t
        L.pmref_total_mem.restype = C.c_size_t
        L.pmref_total_mem.argtypes = [C.c_int]
        L.pmref_reset.argtypes = [C.c_int]
        L.pmref_scan.restype = C.c_double
        L.pmref_scan.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.pmref_summary.restype = C.c_double
        L.pmref_summary.argtypes = [C.c_int, C.c_void_p, C.c_size_t] + [C.POINTER(C.c_uint64)] * 3
        L.pmref_last_hsum.argtypes = [C.POINTER(C.c_uint64)]
        L.pmref_success.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
        L.pmref_scan_parallel.restype = C.c_double
        L.pmref_scan_parallel.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int] + [C.POINTER(C.c_uint64)] * 3 + [C.POINTER(C.c_double)]
        L.pmref_pattern.argtypes = [C.c_size_t] + [C.POINTER(C.c_uint32)] * 5 + [C.POINTER(C.POINTER(C.c_ubyte))]
        self.L = L
        arr = (C.c_char_p * len(dict_paths))(*[os.fsencode(p) for p in dict_paths])
        if L.pmref_build(len(dict_paths), arr, algo_mask) != 0:
            raise RuntimeError("pmref_build failed (already built in this process?)")

    n_patterns = property(lambda s: s.L.pmref_n_patterns())
    max_pat_len = property(lambda s: s.L.pmref_max_pat_len())

    def patterns(self):
        """Unique patterns in the reference's add_pattern order (post-order of its PatternsTree, PatternsTree.c:390-401):
        (file, line, parent_file, parent_line, bytes); parent 0xFFFFFFFF = root."""
        f = C.c_uint32(); l = C.c_uint32(); pf = C.c_uint32(); pl = C.c_uint32(); n = C.c_uint32()
        b = C.POINTER(C.c_ubyte)()
        for i in range(self.n_patterns):
            self.L.pmref_pattern(i, C.byref(f), C.byref(l), C.byref(pf), C.byref(pl), C.byref(n), C.byref(b))
            yield f.value, l.value, pf.value, pl.value, bytes(bytearray(b[:n.value]))

    def total_mem(self, algo=AC):
        return self.L.pmref_total_mem(algo)

    def scan(self, buf, algo=AC, reset=True, want_ids=True):
        a = np.ascontiguousarray(buf, np.uint8)
        if reset:
            self.L.pmref_reset(algo)
        if want_ids:
            fo = np.empty(a.size, np.uint32); lo = np.empty(a.size, np.uint32)
            secs = self.L.pmref_scan(algo, a.ctypes.data, a.size, fo.ctypes.data, lo.ctypes.data)
            return secs, fo, lo
        return self.L.pmref_scan(algo, a.ctypes.data, a.size, None, None), None, None

    def summary(self, buf, algo=AC):
        a = np.ascontiguousarray(buf, np.uint8)
        p = C.c_uint64(); m = C.c_uint64(); h = C.c_uint64()
        secs = self.L.pmref_summary(algo, a.ctypes.data, a.size, C.byref(p), C.byref(m), C.byref(h))
        hs = (C.c_uint64 * 2)(); self.L.pmref_last_hsum(hs)
        return dict(seconds=secs, positions=p.value, matches=m.value, fnv=h.value, hsum_longest=hs[0], hsum_all=hs[1])

    def scan_parallel(self, buf, workers, algo=AC):
        """All-host-cores scan (fork after compile, contiguous shards with a max_pat_len-1 halo)."""
        a = np.ascontiguousarray(buf, np.uint8)
        p = C.c_uint64(); m = C.c_uint64(); h = C.c_uint64(); tmax = C.c_double()
        wall = self.L.pmref_scan_parallel(algo, a.ctypes.data, a.size, workers, C.byref(p), C.byref(m), C.byref(h), C.byref(tmax))
        hs = (C.c_uint64 * 2)(); self.L.pmref_last_hsum(hs)
        return dict(wall_seconds=wall, max_loop_seconds=tmax.value, positions=p.value, matches=m.value,
                    hsum_longest=hs[0], hsum_all=hs[1])
