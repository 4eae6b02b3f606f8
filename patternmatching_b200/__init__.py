"""patternmatching_b200 -- host-side mirror of the reference's algorithm-plugin interface over the
C-ABI of ``libpm_b200.so`` (include/pm_b200.h).

The reference (yehonatan145/PatternMatching) exposes a matcher as seven C callbacks on an opaque
object -- ``MpsElem`` in Core/src/mps.h:71-80: create / add_pattern / compile / read_char /
total_mem / reset / free.  :class:`MpsGpu` has the same seven operations (plus the batched
``read_block``), :class:`Dictionary` is the host dictionary compiler and :class:`Engine` the
device-resident scanner they are built on.  Everything here is ctypes plumbing: the work happens in
the CUDA library, and there is NO CPU fallback -- without the library or without a GPU the calls
raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpm_b200.so")

ALGO_SFX, ALGO_DFA, ALGO_KR, ALGO_AUTO = 0, 1, 2, 3
ALGO_MPBG = 5   # the reference's MPBG as shipped, position for position (pm_b200.h: PM_ALGO_MPBG)
ALGOS = {"sfx": ALGO_SFX, "dfa": ALGO_DFA, "kr": ALGO_KR, "auto": ALGO_AUTO, "mpbg": ALGO_MPBG}
STREAM_UNIFORM, STREAM_PLANTED, STREAM_ALMOST, STREAM_AB = 0, 1, 2, 3
STREAMS = {"uniform": 0, "planted": 1, "almost": 2, "ab": 3, "ascii": 4}
HALO = 352  # bytes of history that make a shard scan identical to the continuous scan (>= max_pat_len-1)


class PmError(RuntimeError):
    pass


class DictInfo(C.Structure):
    _fields_ = [("n_lines", C.c_uint64), ("n_rejected", C.c_uint64), ("n_duplicates", C.c_uint64),
                ("n_patterns", C.c_uint32), ("max_pat_len", C.c_uint32), ("total_pat_bytes", C.c_uint64),
                ("n_ac_states", C.c_uint32), ("n_sfx_nodes", C.c_uint32), ("n_sfx_rows", C.c_uint32),
                ("n_classes", C.c_uint32), ("n_hot2_cont", C.c_uint32), ("table_bytes", C.c_uint64)]


class MpsElemStruct(C.Structure):
    """Layout of MpsElem (Core/src/mps.h:71-80)."""
    _fields_ = [("name", C.c_char_p),
                ("create", C.CFUNCTYPE(C.c_void_p)),
                ("add_pattern", C.CFUNCTYPE(None, C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p)),
                ("compile", C.CFUNCTYPE(None, C.c_void_p)),
                ("read_char", C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_char)),
                ("total_mem", C.CFUNCTYPE(C.c_size_t, C.c_void_p)),
                ("reset", C.CFUNCTYPE(None, C.c_void_p)),
                ("free", C.CFUNCTYPE(None, C.c_void_p))]


_lib = None


def lib():
    """Load libpm_b200.so; fail loudly if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PmError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(or `make -C patternmatching_b200/csrc`); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, sz, u64, u32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32
    sig = {
        "pm_last_error": (C.c_char_p, []),
        "pm_version": (C.c_int, []),
        "pm_dict_create": (vp, []),
        "pm_dict_free": (None, [vp]),
        "pm_parse_pattern_line": (C.c_int, [vp, sz, vp, C.POINTER(sz)]),
        "pm_dict_add_file": (C.c_int, [vp, C.c_char_p]),
        "pm_dict_add_mem": (C.c_int, [vp, vp, sz]),
        "pm_dict_add_pattern": (u32, [vp, vp, sz, u32, u32, u64]),
        "pm_dict_compile": (C.c_int, [vp]),
        "pm_dict_get_info": (C.c_int, [vp, C.POINTER(DictInfo)]),
        "pm_dict_save": (C.c_int, [vp, C.c_char_p]),
        "pm_dict_load": (vp, [C.c_char_p]),
        "pm_dict_compile_files_cached": (vp, [C.POINTER(C.c_char_p), C.c_int, C.c_char_p]),
        "pm_dict_pattern": (C.c_int, [vp, u32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u64), C.POINTER(u32),
                                      C.POINTER(u32), C.POINTER(C.POINTER(C.c_ubyte))]),
        "pm_dict_table": (C.c_int, [vp, C.c_char_p, C.POINTER(vp), C.POINTER(sz)]),
        "pm_dict_is_pattern_suffix": (C.c_int, [vp, u32, u32]),
        "pm_engine_create": (vp, [vp, C.c_int]),
        "pm_engine_free": (None, [vp]),
        "pm_engine_total_mem": (sz, [vp]),
        "pm_engine_set_kr_seed": (C.c_int, [vp, u64]),
        "pm_engine_scan_device": (C.c_int, [vp, C.c_int, vp, sz, sz, vp, vp]),
        "pm_engine_scan_device32": (C.c_int, [vp, C.c_int, vp, sz, sz, vp, vp]),
        "pm_engine_prepare_host": (C.c_int, [vp]),
        "pm_engine_generate_host": (C.c_int, [vp, C.c_int, u64, sz, vp]),
        "pm_engine_scan_host": (C.c_int, [vp, C.c_int, vp, sz, vp]),
        "pm_engine_scan_device_records": (C.c_int, [vp, C.c_int, vp, sz, sz, u64, u32, vp, vp, sz, C.POINTER(u64), vp]),
        "pm_engine_scan_host_ids": (C.c_int, [vp, C.c_int, vp, sz, vp, sz, vp]),
        "pm_engine_scratch_mem": (sz, [vp]),
        "pm_engine_host_threads": (C.c_int, [vp]),
        "pm_engine_scan_host_records": (C.c_int, [vp, C.c_int, vp, sz, u32, vp, sz, C.POINTER(u64)]),
        "pm_engine_reset": (None, [vp]),
        "pm_engine_summarize": (C.c_int, [vp, vp, sz, u64, C.POINTER(u64), vp]),
        "pm_engine_classify": (C.c_int, [vp, vp, vp, sz, C.POINTER(u64), vp]),
        "pm_engine_compact": (C.c_int, [vp, vp, sz, u64, C.c_int, vp, sz, C.POINTER(u64), vp]),
        "pm_engine_generate": (C.c_int, [vp, C.c_int, u64, sz, vp, vp]),
        "pm_engine_time_scan": (C.c_int, [vp, C.c_int, vp, sz, sz, vp, C.c_int, C.POINTER(C.c_float), vp]),
        "pm_engine_launch_count": (u64, [vp]),
        "pm_engine_set_profiling": (C.c_int, [vp, C.c_int]),
        "pm_engine_read_profile": (C.c_int, [vp, C.POINTER(u32), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "pm_engine_last_deferred": (u64, [vp]),
        "pm_engine_auto_choice": (C.c_int, [vp]),
        "pm_comm_last_error": (C.c_char_p, []),
        "pm_comm_unique_id": (C.c_int, [vp]),
        "pm_comm_create": (vp, [vp, C.c_int, C.c_int, C.c_int]),
        "pm_comm_free": (None, [vp]),
        "pm_comm_gather_records": (C.c_int, [vp, vp, u64, vp, u64, C.POINTER(u64), C.POINTER(u64), C.c_int, vp]),
        "pm_host_alloc": (vp, [sz]),
        "pm_host_free": (None, [vp]),
        "pm_host_register": (C.c_int, [vp, sz]),
        "pm_host_unregister": (C.c_int, [vp]),
        "gpu_create": (vp, []), "gpu_dfa_create": (vp, []), "gpu_kr_create": (vp, []), "gpu_mpbg_create": (vp, []),
        "gpu_add_pattern": (None, [vp, C.c_char_p, sz, vp]),
        "gpu_compile": (None, [vp]),
        "gpu_read_char": (vp, [vp, C.c_char]),
        "gpu_read_block": (sz, [vp, vp, sz, vp]),
        "gpu_total_mem": (sz, [vp]),
        "gpu_reset": (None, [vp]),
        "gpu_free": (None, [vp]),
        "mps_gpu_register_into": (None, [C.POINTER(MpsElemStruct)]),
        "mps_gpu_dfa_register_into": (None, [C.POINTER(MpsElemStruct)]),
        "mps_gpu_kr_register_into": (None, [C.POINTER(MpsElemStruct)]),
        "mps_gpu_mpbg_register_into": (None, [C.POINTER(MpsElemStruct)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _err(L, what):
    return PmError(f"{what}: {L.pm_last_error().decode(errors='replace')}")


def _u8(b):
    if isinstance(b, (bytes, bytearray, memoryview)):
        return np.frombuffer(b, dtype=np.uint8)
    return np.ascontiguousarray(b, dtype=np.uint8)


def parse_pattern_line(line: bytes):
    """One .dict line -> pattern bytes, or None when the reference grammar rejects the line
    (Core/src/parser.c:63-99)."""
    L = lib()
    a = _u8(line)
    out = np.empty(max(a.size, 1), np.uint8)
    n = C.c_size_t()
    ok = L.pm_parse_pattern_line(a.ctypes.data if a.size else None, a.size, out.ctypes.data, C.byref(n))
    return out[:n.value].tobytes() if ok else None


class Dictionary:
    """Host dictionary compiler: merged .dict files -> flat device tables.
    Mirrors patterns_tree_build + ac_add_pattern + ac_compile (PatternsTree.c:469-475, mpac.c:257-291)."""

    def __init__(self):
        self.L = lib()
        self.h = self.L.pm_dict_create()
        self.compiled = False

    def __del__(self):
        try:
            if self.h:
                self.L.pm_dict_free(self.h)
                self.h = None
        except Exception:
            pass

    def save(self, path):
        if self.L.pm_dict_save(self.h, os.fsencode(path)) != 0:
            raise _err(self.L, "pm_dict_save")

    @classmethod
    def load(cls, path):
        self = cls.__new__(cls)
        self.L = lib()
        self.h = self.L.pm_dict_load(os.fsencode(path))
        if not self.h:
            raise _err(self.L, "pm_dict_load")
        self.compiled = True
        return self

    @classmethod
    def from_files_cached(cls, paths, cache_dir):
        """Load the compiled tables from <cache_dir>/pmdict-<hash of the files>.bin, or compile and store them."""
        self = cls.__new__(cls)
        self.L = lib()
        arr = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
        self.h = self.L.pm_dict_compile_files_cached(arr, len(paths), os.fsencode(cache_dir))
        if not self.h:
            raise _err(self.L, "pm_dict_compile_files_cached")
        self.compiled = True
        return self

    def add_file(self, path):
        if self.L.pm_dict_add_file(self.h, os.fsencode(path)) != 0:
            raise _err(self.L, "pm_dict_add_file")
        return self

    def add_bytes(self, data: bytes):
        a = _u8(data)
        if self.L.pm_dict_add_mem(self.h, a.ctypes.data if a.size else None, a.size) != 0:
            raise _err(self.L, "pm_dict_add_mem")
        return self

    def add_pattern(self, pat: bytes, file=0, line=0, user_id=0):
        a = _u8(pat)
        return self.L.pm_dict_add_pattern(self.h, a.ctypes.data if a.size else None, a.size, file, line, user_id)

    def compile(self):
        if self.L.pm_dict_compile(self.h) != 0:
            raise _err(self.L, "pm_dict_compile")
        self.compiled = True
        return self

    @property
    def info(self):
        i = DictInfo()
        self.L.pm_dict_get_info(self.h, C.byref(i))
        return i

    @property
    def n_patterns(self):
        return self.info.n_patterns

    @property
    def max_pat_len(self):
        return self.info.max_pat_len

    def pattern(self, pid):
        """pid (1..P) -> (file, line, user_id, parent_pid, bytes)"""
        f = C.c_uint32(); l = C.c_uint32(); u = C.c_uint64(); p = C.c_uint32(); n = C.c_uint32()
        b = C.POINTER(C.c_ubyte)()
        if self.L.pm_dict_pattern(self.h, pid, C.byref(f), C.byref(l), C.byref(u), C.byref(p), C.byref(n), C.byref(b)) != 0:
            raise _err(self.L, "pm_dict_pattern")
        return f.value, l.value, u.value, p.value, bytes(bytearray(b[:n.value]))

    def id_arrays(self):
        """(file[pid], line[pid]) as uint32 arrays of length P+1; entry 0 = 0xFFFFFFFF (no pattern)."""
        P = self.n_patterns
        files = np.full(P + 1, 0xFFFFFFFF, np.uint32)
        lines = np.full(P + 1, 0xFFFFFFFF, np.uint32)
        f = C.c_uint32(); l = C.c_uint32()
        for pid in range(1, P + 1):
            self.L.pm_dict_pattern(self.h, pid, C.byref(f), C.byref(l), None, None, None, None)
            files[pid] = f.value
            lines[pid] = l.value
        return files, lines

    def table(self, name, dtype):
        """Read-only numpy view of a compiled table (see pm_dict_table in include/pm_b200.h)."""
        ptr = C.c_void_p(); n = C.c_size_t()
        if self.L.pm_dict_table(self.h, name.encode(), C.byref(ptr), C.byref(n)) != 0:
            raise _err(self.L, "pm_dict_table")
        if n.value == 0:
            return np.zeros(0, dtype)
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_ubyte)), shape=(n.value,)).view(dtype)
        a.flags.writeable = False
        return a

    def is_pattern_suffix(self, first_pid, second_pid):
        return bool(self.L.pm_dict_is_pattern_suffix(self.h, first_pid, second_pid))


def _ptr(x):
    """device pointer of a torch tensor / raw int"""
    if isinstance(x, int):
        return x
    return x.data_ptr()


class Engine:
    """Device-resident scanner.  Raises when no CUDA device is usable (no CPU fallback)."""

    def __init__(self, dictionary: Dictionary, device=0):
        self.L = lib()
        self.dict = dictionary
        if not dictionary.compiled:
            dictionary.compile()
        self.device = device
        self.h = self.L.pm_engine_create(dictionary.h, device)
        if not self.h:
            raise _err(self.L, "pm_engine_create")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.pm_engine_free(self.h)
                self.h = None
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise _err(self.L, what)

    @property
    def total_mem(self):
        return self.L.pm_engine_total_mem(self.h)

    @property
    def launches(self):
        return self.L.pm_engine_launch_count(self.h)

    def set_kr_seed(self, seed):
        self.L.pm_engine_set_kr_seed(self.h, seed)

    def reset(self):
        self.L.pm_engine_reset(self.h)

    def scan_device(self, d_stream, n, d_out, hist_valid=0, algo=ALGO_SFX, cuda_stream=0):
        self._check(self.L.pm_engine_scan_device(self.h, algo, _ptr(d_stream), n, hist_valid, _ptr(d_out), cuda_stream),
                    "pm_engine_scan_device")

    def scan_device32(self, d_stream, n, d_out32, hist_valid=0, algo=ALGO_SFX, cuda_stream=0):
        """32-bit results (any number of patterns): d_out32[i] = global pid of the longest pattern ending at i."""
        self._check(self.L.pm_engine_scan_device32(self.h, algo, _ptr(d_stream), n, hist_valid, _ptr(d_out32), cuda_stream),
                    "pm_engine_scan_device32")

    def scan_device_records(self, d_stream, n, d_out, d_records, cap, min_len=4, hist_valid=0, pos_base=0, algo=ALGO_SFX,
                            cuda_stream=0):
        """Sparse mode: dense result in d_out, position-sorted (pos << 24 | pid) records of the matches with >= min_len
        pattern bytes in d_records; returns the number of such matches."""
        cnt = C.c_uint64()
        self._check(self.L.pm_engine_scan_device_records(self.h, algo, _ptr(d_stream), n, hist_valid, pos_base, min_len,
                                                          _ptr(d_out), _ptr(d_records), cap, C.byref(cnt), cuda_stream),
                    "pm_engine_scan_device_records")
        return cnt.value

    def scan_host(self, buf, algo=ALGO_SFX, out=None):
        a = _u8(buf)
        if out is None:
            out = np.empty(a.size, np.uint16)
        self._check(self.L.pm_engine_scan_host(self.h, algo, a.ctypes.data if a.size else None, a.size,
                                                out.ctypes.data if a.size else None), "pm_engine_scan_host")
        return out

    def scan_host_records(self, buf, min_len=1, cap=None, algo=ALGO_SFX, src_ptr=None, n=None, dst_ptr=None):
        """Sparse result: (pos << 24 | pid) records of the positions whose longest match has >= min_len bytes."""
        cnt = C.c_uint64()
        if src_ptr is None:
            a = _u8(buf)
            n, src_ptr = a.size, (a.ctypes.data if a.size else None)
        if cap is None:
            cap = n
        out = None
        if dst_ptr is None:
            out = np.empty(max(cap, 1), np.uint64)
            dst_ptr = out.ctypes.data
        self._check(self.L.pm_engine_scan_host_records(self.h, algo, src_ptr, n, min_len, dst_ptr, cap, C.byref(cnt)),
                    "pm_engine_scan_host_records")
        return (out[:min(cnt.value, cap)] if out is not None else None), cnt.value

    def scan_host_ids(self, buf, id_of_pid, algo=ALGO_SFX, out=None):
        """Dense result translated through a pid -> 64-bit id table (what the plugin's read_block returns)."""
        a = _u8(buf)
        t = np.ascontiguousarray(id_of_pid, dtype=np.uint64)
        if out is None:
            out = np.empty(a.size, np.uint64)
        self._check(self.L.pm_engine_scan_host_ids(self.h, algo, a.ctypes.data if a.size else None, a.size, t.ctypes.data,
                                                    t.size, out.ctypes.data if a.size else None), "pm_engine_scan_host_ids")
        return out

    @property
    def scratch_mem(self):
        return self.L.pm_engine_scratch_mem(self.h)

    @property
    def host_threads(self):
        return self.L.pm_engine_host_threads(self.h)

    def scan_host_ptr(self, src_ptr, n, dst_ptr, algo=ALGO_SFX):
        self._check(self.L.pm_engine_scan_host(self.h, algo, src_ptr, n, dst_ptr), "pm_engine_scan_host")

    def summarize(self, d_out, n, pos_base=0, cuda_stream=0):
        o = (C.c_uint64 * 4)()
        self._check(self.L.pm_engine_summarize(self.h, _ptr(d_out), n, pos_base, o, cuda_stream), "pm_engine_summarize")
        return dict(positions=o[0], matches=o[1], hsum_longest=o[2], hsum_all=o[3])

    def classify(self, d_algo, d_real, n, cuda_stream=0):
        """measure_success_rate (measure.c:174-190) on the device: counts of success / partial / false_neg / false_pos."""
        o = (C.c_uint64 * 4)()
        self._check(self.L.pm_engine_classify(self.h, _ptr(d_algo), _ptr(d_real), n, o, cuda_stream), "pm_engine_classify")
        return dict(success=o[0], partial=o[1], false_neg=o[2], false_pos=o[3])

    def compact(self, d_out, n, d_records, cap, pos_base=0, expand_ancestors=False, cuda_stream=0):
        cnt = C.c_uint64()
        self._check(self.L.pm_engine_compact(self.h, _ptr(d_out), n, pos_base, int(expand_ancestors), _ptr(d_records), cap,
                                             C.byref(cnt), cuda_stream), "pm_engine_compact")
        return cnt.value

    def generate(self, kind, off, n, d_dst, cuda_stream=0):
        k = STREAMS[kind] if isinstance(kind, str) else kind
        self._check(self.L.pm_engine_generate(self.h, k, off, n, _ptr(d_dst), cuda_stream), "pm_engine_generate")

    @property
    def auto_choice(self):
        return self.L.pm_engine_auto_choice(self.h)

    @property
    def last_deferred(self):
        return self.L.pm_engine_last_deferred(self.h)

    def set_profiling(self, on=True):
        self.L.pm_engine_set_profiling(self.h, int(on))

    def read_profile(self):
        """-> (n_scans, ms summed over the dominant kernel, ms summed over whole scans)"""
        n = C.c_uint32(); a = C.c_float(); b = C.c_float()
        self._check(self.L.pm_engine_read_profile(self.h, C.byref(n), C.byref(a), C.byref(b)), "pm_engine_read_profile")
        return n.value, a.value, b.value

    def time_scan(self, d_stream, n, d_out, hist_valid=0, algo=ALGO_SFX, iters=1, cuda_stream=0):
        ms = C.c_float()
        self._check(self.L.pm_engine_time_scan(self.h, algo, _ptr(d_stream), n, hist_valid, _ptr(d_out), iters,
                                               C.byref(ms), cuda_stream), "pm_engine_time_scan")
        return ms.value


class Comm:
    """The library's NCCL communicator for the record gather (include/pm_b200.h: pm_comm_*).  One per rank."""

    def __init__(self, unique_id: bytes, rank: int, world: int, device: int):
        self.L = lib()
        self.rank, self.world = rank, world
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self.h = self.L.pm_comm_create(buf, rank, world, device)
        if not self.h:
            raise PmError("pm_comm_create: " + self.L.pm_comm_last_error().decode(errors="replace"))

    @staticmethod
    def unique_id() -> bytes:
        L = lib()
        buf = (C.c_ubyte * 128)()
        if L.pm_comm_unique_id(buf) != 0:
            raise PmError("pm_comm_unique_id: " + L.pm_comm_last_error().decode(errors="replace"))
        return bytes(buf)

    @classmethod
    def from_torch(cls, dist, device_index):
        """Create the communicator inside a torch.distributed job: rank 0's unique id is broadcast with torch."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        t = torch.zeros(128, dtype=torch.uint8, device=torch.device("cuda", device_index))
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return cls(bytes(t.cpu().numpy().tobytes()), rank, world, device_index)

    def gather_records(self, d_local, n_local, d_all, cap, root=0, cuda_stream=0):
        """-> (per-rank counts, total); on `root` d_all holds the concatenation once cuda_stream has been synchronised."""
        counts = (C.c_uint64 * self.world)()
        total = C.c_uint64()
        rc = self.L.pm_comm_gather_records(self.h, _ptr(d_local) if d_local is not None else None, n_local,
                                           _ptr(d_all) if d_all is not None else None, cap, counts, C.byref(total), root, cuda_stream)
        if rc != 0:
            raise PmError("pm_comm_gather_records: " + self.L.pm_comm_last_error().decode(errors="replace"))
        return list(counts), total.value

    def free(self):
        if self.h:
            self.L.pm_comm_free(self.h)
            self.h = None


class PinnedBuffer:
    """Page-locked host memory as numpy arrays (input / result buffers of Engine.scan_host)."""

    def __init__(self, nbytes):
        self.L = lib()
        self.nbytes = nbytes
        self.ptr = self.L.pm_host_alloc(nbytes)
        if not self.ptr:
            raise _err(self.L, "pm_host_alloc")

    def array(self, dtype=np.uint8):
        n = self.nbytes // np.dtype(dtype).itemsize
        return np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_ubyte)), shape=(self.nbytes,)).view(dtype)[:n]

    def free(self):
        if self.ptr:
            self.L.pm_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MpsGpu:
    """The reference's plugin operations (MpsElem, Core/src/mps.h:71-80) on the GPU object, in the
    order the reference driver calls them: create -> add_pattern* -> compile -> (reset -> read_char*)*.
    ``pattern ids`` are opaque non-zero integers (pattern_id_t is a pointer in the reference;
    0 / None is null_pattern_id)."""

    _CREATE = {"sfx": "gpu_create", "dfa": "gpu_dfa_create", "kr": "gpu_kr_create", "mpbg": "gpu_mpbg_create"}

    def __init__(self, algo="sfx"):
        self.L = lib()
        self.obj = getattr(self.L, self._CREATE[algo])()          # MpsElem.create

    def add_pattern(self, pat: bytes, pattern_id: int):            # MpsElem.add_pattern
        self.L.gpu_add_pattern(self.obj, bytes(pat), len(pat), pattern_id)

    def compile(self):                                              # MpsElem.compile
        self.L.gpu_compile(self.obj)

    def read_char(self, c: int):                                    # MpsElem.read_char
        r = self.L.gpu_read_char(self.obj, C.c_char(bytes([c & 0xFF])))
        return r or 0

    def read_block(self, buf, out=None):                            # batched extension
        a = _u8(buf)
        if out is None:
            out = np.zeros(a.size, np.uint64)
        self.L.gpu_read_block(self.obj, a.ctypes.data if a.size else None, a.size, out.ctypes.data)
        return out

    def read_block_ptr(self, src_ptr, n, dst_ptr):                  # same call on raw host pointers (bench)
        self.L.gpu_read_block(self.obj, src_ptr, n, dst_ptr)

    def total_mem(self):                                            # MpsElem.total_mem
        return self.L.gpu_total_mem(self.obj)

    def reset(self):                                                # MpsElem.reset
        self.L.gpu_reset(self.obj)

    def free(self):                                                 # MpsElem.free
        if self.obj:
            self.L.gpu_free(self.obj)
            self.obj = None
