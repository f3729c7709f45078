"""ctypes front end of oracle/liboracle.so (the plain-C restatement of SAGE2 steps 1-3).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg, never by sage2_b200 (the product).  Also drives the compiled UNMODIFIED reference in
oracle/_ref/ (when it was built in the build container and travelled with the repo).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_SAGE2 = os.path.join(HERE, "_ref", "SAGE2")
REF_STEPS = os.path.join(HERE, "_ref", "ref_steps123")


class Ext(C.Structure):
    _fields_ = [("id", C.c_uint64), ("type", C.c_uint32), ("length", C.c_uint32)]


class Edge(C.Structure):
    _fields_ = [("frm", C.c_uint64), ("to", C.c_uint64), ("type", C.c_uint32),
                ("delta", C.c_uint32), ("delta_twin", C.c_uint32)]


class Result(C.Structure):
    _fields_ = [
        ("total_reads", C.c_uint64), ("good_reads", C.c_uint64), ("unique_reads", C.c_uint64),
        ("total_bp", C.c_uint64), ("avg_len", C.c_uint64),
        ("length", C.POINTER(C.c_uint16)), ("frequency", C.POINTER(C.c_uint16)),
        ("byte_off", C.POINTER(C.c_uint64)), ("fwd", C.POINTER(C.c_uint8)), ("rc", C.POINTER(C.c_uint8)),
        ("hash_len", C.c_uint64), ("distinct_keys", C.c_uint64), ("keys_over_threshold", C.c_uint64),
        ("right_ext", C.POINTER(Ext)), ("left_ext", C.POINTER(Ext)),
        ("explored_a", C.POINTER(C.c_uint8)), ("explored_b", C.POINTER(C.c_uint8)),
        ("compare_calls", C.c_uint64),
        ("contained_ext", C.c_uint64), ("contained_size", C.c_uint64), ("left_to_explore", C.c_uint64),
        ("edges_inserted_c", C.c_uint64), ("transitive_removed", C.c_uint64),
        ("n_edges", C.c_uint64), ("edges", C.POINTER(Edge)),
    ]


EXT_DT = np.dtype([("id", "<u8"), ("type", "<u4"), ("length", "<u4")])
EDGE_DT = np.dtype([("from", "<u8"), ("to", "<u8"), ("type", "<u4"), ("delta", "<u4"),
                    ("delta_twin", "<u4"), ("_pad", "<u4")])

_lib = None


def build(force: bool = False) -> None:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "sage2_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.sgo_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(Result)]
        _lib.sgo_run.restype = C.c_int
        _lib.sgo_free.argtypes = [C.POINTER(Result)]
        _lib.sgo_write_reads.argtypes = [C.POINTER(Result), C.c_char_p]
        _lib.sgo_write_graph3.argtypes = [C.POINTER(Result), C.c_char_p]
        _lib.sgo_chars_to_bytes.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
        _lib.sgo_get64.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _lib.sgo_get64.restype = C.c_uint64
        _lib.sgo_string_compare.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _lib.sgo_string_compare.restype = C.c_int
        _lib.sgo_map_reads.argtypes = [C.POINTER(Result), C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        _lib.sgo_map_reads.restype = C.c_int
    return _lib


class OracleRun:
    """Owns one sgo_result; numpy views are copies so they outlive free()."""

    def __init__(self, bases: np.ndarray, offsets: np.ndarray, k: int, threads: int = 0):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self._r = Result()
        rc = lib().sgo_run(bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1, k, threads, C.byref(self._r))
        if rc != 0:
            raise RuntimeError("sgo_run failed")
        r = self._r
        U = self.U = int(r.unique_reads)
        self.N = int(r.good_reads)
        self.total_reads = int(r.total_reads)
        self.avg_len = int(r.avg_len)
        self.hash_len = int(r.hash_len)
        self.distinct_keys = int(r.distinct_keys)
        self.keys_over_threshold = int(r.keys_over_threshold)
        self.compare_calls = int(r.compare_calls)
        self.contained_ext = int(r.contained_ext)
        self.contained_size = int(r.contained_size)
        self.left_to_explore = int(r.left_to_explore)
        self.edges_inserted_c = int(r.edges_inserted_c)
        self.transitive_removed = int(r.transitive_removed)
        self.length = np.ctypeslib.as_array(r.length, shape=(U + 1,)).copy()
        self.frequency = np.ctypeslib.as_array(r.frequency, shape=(U + 1,)).copy()
        self.byte_off = np.ctypeslib.as_array(r.byte_off, shape=(U + 2,)).copy()
        nb = int(self.byte_off[U + 1])
        self.fwd = np.ctypeslib.as_array(r.fwd, shape=(max(nb, 1),)).copy()[:nb]
        self.rc = np.ctypeslib.as_array(r.rc, shape=(max(nb, 1),)).copy()[:nb]

        def ext(p):
            a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=((U + 1) * 16,)).copy()
            return a.view(EXT_DT)
        self.right_ext = ext(r.right_ext)
        self.left_ext = ext(r.left_ext)
        self.explored_a = np.ctypeslib.as_array(r.explored_a, shape=(U + 1,)).copy()
        self.explored_b = np.ctypeslib.as_array(r.explored_b, shape=(U + 1,)).copy()
        E = self.n_edges = int(r.n_edges)
        if E:
            a = np.ctypeslib.as_array(C.cast(r.edges, C.POINTER(C.c_uint8)), shape=(E * 32,)).copy()
            self.edges = a.view(EDGE_DT)
        else:
            self.edges = np.zeros(0, dtype=EDGE_DT)

    def map_reads(self, bases: np.ndarray, offsets: np.ndarray, k: int):
        """getIdOfRead of every read (readLoader.cpp:319-353): (signed ids int64, isGoodRead uint8)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        ids, good = np.zeros(n, np.int64), np.zeros(n, np.uint8)
        if lib().sgo_map_reads(C.byref(self._r), bases.ctypes.data, offsets.ctypes.data, n, k, ids.ctypes.data, good.ctypes.data) != 0:
            raise RuntimeError("sgo_map_reads failed")
        return ids, good

    def write_reads(self, path: str) -> None:
        if lib().sgo_write_reads(C.byref(self._r), path.encode()) != 0:
            raise OSError(path)

    def write_graph3(self, path: str) -> None:
        if lib().sgo_write_graph3(C.byref(self._r), path.encode()) != 0:
            raise OSError(path)

    def close(self):
        if self._r is not None:
            lib().sgo_free(C.byref(self._r))
            self._r = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def have_reference() -> bool:
    return os.access(REF_SAGE2, os.X_OK)


def run_reference(fastq: str, k: int, outdir: str, prefix: str, max_step: int = 3, threads: int | None = None,
                  save: bool = True, min_step: int = 1, input_prefix: str | None = None,
                  timeout: float = 600.0) -> str:
    """Run the compiled UNMODIFIED reference; returns the path prefix of its outputs."""
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    os.makedirs(outdir, exist_ok=True)
    cmd = [REF_SAGE2, "-f", fastq, "-k", str(k), "-o", outdir.rstrip("/") + "/", "-p", prefix,
           "-M", str(max_step), "-m", str(min_step)]
    if save:
        cmd.append("-s")
    if input_prefix:
        cmd += ["-i", input_prefix]
    subprocess.check_call(cmd, env=env, stdout=subprocess.DEVNULL, timeout=timeout)
    return os.path.join(outdir, prefix)
