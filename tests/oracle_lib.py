"""ctypes wrapper around oracle/liboracle.so (the CPU restatement) -- test infrastructure only."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
DATA_DIR = os.path.join(ORACLE_DIR, "_ref", "data")


class PmoSummary(C.Structure):
    _fields_ = [("positions", C.c_uint64), ("matches", C.c_uint64), ("fnv", C.c_uint64),
                ("hsum_longest", C.c_uint64), ("hsum_all", C.c_uint64)]


_lib = None


def build_oracle():
    """Build liboracle.so (and oracle/_ref when /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        vp, sz, u8p = C.c_void_p, C.c_size_t, C.c_void_p
        L.pmo_create.restype = vp
        L.pmo_free.argtypes = [vp]
        L.pmo_add_dict_file.argtypes = [vp, C.c_char_p]
        L.pmo_add_dict_mem.argtypes = [vp, u8p, sz]
        L.pmo_add_pattern.argtypes = [vp, u8p, sz, C.c_uint32, C.c_uint32]
        L.pmo_compile.argtypes = [vp]
        for f in ("pmo_n_patterns", "pmo_n_states", "pmo_max_pat_len", "pmo_n_lines", "pmo_n_rejected", "pmo_n_duplicates"):
            getattr(L, f).restype = sz
            getattr(L, f).argtypes = [vp]
        L.pmo_pattern.argtypes = [vp, sz, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_uint32), C.POINTER(C.POINTER(C.c_ubyte))]
        L.pmo_reset.argtypes = [vp]
        L.pmo_scan.argtypes = [vp, u8p, sz, vp]
        L.pmo_summary.argtypes = [vp, u8p, sz, sz, C.c_uint64, C.POINTER(PmoSummary)]
        L.pmo_classify.argtypes = [vp, vp, vp, sz, C.POINTER(C.c_uint64)]
        L.pmo_is_pattern_suffix.argtypes = [vp, C.c_int32, C.c_int32]
        L.pmo_parse_line.argtypes = [u8p, sz, u8p, C.POINTER(sz)]
        L.pmo_kmp_search.restype = sz
        L.pmo_kmp_search.argtypes = [u8p, sz, u8p, sz, vp, sz]
        for f in ("pmo_mulmod", "pmo_powmod"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_uint64, C.c_uint64]
        L.pmo_invmod.restype = C.c_uint64
        L.pmo_invmod.argtypes = [C.c_uint64]
        L.pmo_fp.restype = C.c_uint64
        L.pmo_fp.argtypes = [u8p, sz, C.c_uint64]
        L.pmo_kr_seed_r.restype = C.c_uint64
        L.pmo_kr_seed_r.argtypes = [C.c_uint64]
        L.pmo_kr_scan.argtypes = [vp, C.c_uint64, u8p, sz, sz, vp]
        L.pmo_splitmix64.restype = C.c_uint64
        L.pmo_splitmix64.argtypes = [C.c_uint64]
        L.pmo_gen_uniform.argtypes = [C.c_uint64, sz, u8p]
        L.pmo_gen_planted.argtypes = [vp, C.c_uint64, sz, u8p]
        L.pmo_gen_almost.argtypes = [vp, C.c_uint64, sz, u8p]
        L.pmo_gen_ab.argtypes = [C.c_uint64, sz, u8p]
        L.pmo_gen_ascii.argtypes = [C.c_uint64, sz, u8p]
        _lib = L
    return _lib


def _buf(b):
    a = np.frombuffer(b, dtype=np.uint8) if isinstance(b, (bytes, bytearray)) else np.ascontiguousarray(b, dtype=np.uint8)
    return a


def parse_line(line: bytes):
    """Return the pattern bytes, or None if the reference grammar rejects the line."""
    L = lib()
    a = _buf(line)
    out = np.empty(max(len(line), 1), np.uint8)
    n = C.c_size_t()
    ok = L.pmo_parse_line(a.ctypes.data if len(line) else None, len(line), out.ctypes.data, C.byref(n))
    return out[:n.value].tobytes() if ok else None


class Oracle:
    def __init__(self):
        self.L = lib()
        self.h = self.L.pmo_create()

    def __del__(self):
        try:
            self.L.pmo_free(self.h)
        except Exception:
            pass

    def add_dict_file(self, path):
        assert self.L.pmo_add_dict_file(self.h, path.encode()) == 0, path

    def add_dict_bytes(self, data: bytes):
        a = _buf(data)
        self.L.pmo_add_dict_mem(self.h, a.ctypes.data if a.size else None, a.size)

    def add_pattern(self, pat: bytes, file: int, line: int):
        a = _buf(pat)
        return self.L.pmo_add_pattern(self.h, a.ctypes.data, a.size, file, line)

    def compile(self):
        assert self.L.pmo_compile(self.h) == 0

    n_patterns = property(lambda s: s.L.pmo_n_patterns(s.h))
    n_states = property(lambda s: s.L.pmo_n_states(s.h))
    max_pat_len = property(lambda s: s.L.pmo_max_pat_len(s.h))
    n_lines = property(lambda s: s.L.pmo_n_lines(s.h))
    n_rejected = property(lambda s: s.L.pmo_n_rejected(s.h))
    n_duplicates = property(lambda s: s.L.pmo_n_duplicates(s.h))

    def pattern(self, i):
        f = C.c_uint32(); l = C.c_uint32(); p = C.c_int32(); n = C.c_uint32(); b = C.POINTER(C.c_ubyte)()
        assert self.L.pmo_pattern(self.h, i, C.byref(f), C.byref(l), C.byref(p), C.byref(n), C.byref(b)) == 0
        return f.value, l.value, p.value, bytes(bytearray(b[:n.value]))

    def id_arrays(self):
        P = self.n_patterns
        files = np.empty(P, np.uint32); lines = np.empty(P, np.uint32)
        f = C.c_uint32(); l = C.c_uint32()
        for i in range(P):
            self.L.pmo_pattern(self.h, i, C.byref(f), C.byref(l), None, None, None)
            files[i] = f.value; lines[i] = l.value
        return files, lines

    def parents(self):
        P = self.n_patterns
        par = np.empty(P, np.int32); p = C.c_int32()
        for i in range(P):
            self.L.pmo_pattern(self.h, i, None, None, C.byref(p), None, None)
            par[i] = p.value
        return par

    def lengths(self):
        P = self.n_patterns
        ln = np.empty(P, np.uint32); n = C.c_uint32()
        for i in range(P):
            self.L.pmo_pattern(self.h, i, None, None, None, C.byref(n), None)
            ln[i] = n.value
        return ln

    def reset(self):
        self.L.pmo_reset(self.h)

    def scan(self, buf, reset=True):
        a = _buf(buf)
        out = np.empty(a.size, np.int32)
        if reset:
            self.reset()
        self.L.pmo_scan(self.h, a.ctypes.data, a.size, out.ctypes.data)
        return out

    def summary(self, buf, skip=0, pos_base=0, reset=True):
        a = _buf(buf)
        s = PmoSummary()
        if reset:
            self.reset()
        self.L.pmo_summary(self.h, a.ctypes.data, a.size, skip, pos_base, C.byref(s))
        return s

    def classify(self, algo, real):
        algo = np.ascontiguousarray(algo, np.int32); real = np.ascontiguousarray(real, np.int32)
        cnt = (C.c_uint64 * 4)()
        self.L.pmo_classify(self.h, algo.ctypes.data, real.ctypes.data, algo.size, cnt)
        return dict(success=cnt[0], partial=cnt[1], false_neg=cnt[2], false_pos=cnt[3])

    def kr_scan(self, buf, seed):
        a = _buf(buf)
        out = np.empty(a.size, np.int32)
        self.L.pmo_kr_scan(self.h, seed, a.ctypes.data, a.size, 0, out.ctypes.data)
        return out

    def gen(self, kind, off, n):
        out = np.empty(n, np.uint8)
        if kind == "uniform":
            self.L.pmo_gen_uniform(off, n, out.ctypes.data)
        elif kind == "planted":
            self.L.pmo_gen_planted(self.h, off, n, out.ctypes.data)
        elif kind == "almost":
            self.L.pmo_gen_almost(self.h, off, n, out.ctypes.data)
        elif kind == "ab":
            self.L.pmo_gen_ab(off, n, out.ctypes.data)
        elif kind == "ascii":
            self.L.pmo_gen_ascii(off, n, out.ctypes.data)
        else:
            raise ValueError(kind)
        return out
