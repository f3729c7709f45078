"""ctypes loader for tests/host_emul.cpp (TEST INFRASTRUCTURE: CPU emulation that drives the product's
core.cuh primitives + host_phase_c.cpp; see the header of host_emul.cpp)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_build", "host_emul.so")
SRCS = [os.path.join(HERE, "host_emul.cpp"), os.path.join(ROOT, "sage2_b200", "csrc", "host_phase_c.cpp")]
DEPS = SRCS + [os.path.join(ROOT, "sage2_b200", "csrc", "core.cuh"), os.path.join(ROOT, "sage2_b200", "csrc", "host_phase_c.h")]

_lib = None


def lib():
    global _lib
    if _lib is None:
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
            subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-x", "c++", "-o", SO] + SRCS)
        _lib = C.CDLL(SO)
        _lib.hemu_run.restype = C.c_void_p
        _lib.hemu_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        _lib.hemu_free.argtypes = [C.c_void_p]
        _lib.hemu_prepare.restype = C.c_void_p
        _lib.hemu_prepare.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        _lib.hemu_phase_a.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _lib.hemu_phase_a_arrays.restype = C.c_uint64
        _lib.hemu_phase_a_arrays.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 4
        _lib.hemu_finish.argtypes = [C.c_void_p]
        u64p, vpp = C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)
        _lib.hemu_shard_build.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _lib.hemu_phase_a_sharded_begin.argtypes = [C.c_void_p, C.c_int, C.c_int, u64p, u64p]
        _lib.hemu_route_begin.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, vpp, u64p, u64p]
        _lib.hemu_shard_answer.argtypes = [C.c_void_p, C.c_void_p, u64p, C.c_int, C.c_int, vpp, vpp, u64p]
        _lib.hemu_route_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, u64p]
        _lib.hemu_phase_a_routed.restype = C.c_uint64
        _lib.hemu_phase_a_routed.argtypes = [C.c_void_p]
        _lib.hemu_phase_b.argtypes = [C.c_void_p]
        _lib.hemu_sizes.argtypes = [C.c_void_p, C.c_void_p]
        _lib.hemu_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 9
        _lib.hemu_get_bases.restype = C.c_uint64
        _lib.hemu_get_bases.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib.hemu_pack.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        _lib.hemu_revcomp.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _lib.hemu_key.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.hemu_overlap.restype = C.c_int
        _lib.hemu_overlap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return _lib


class EmuRun:
    """Whole pipeline at once, or (rank, world, exchange) the multi-GPU split: steps 1-2, the rank's slice of
    phase A, `exchange(views, chunk)` on in-place numpy views of the padded phase-A arrays, then the rest."""

    def __init__(self, bases, offsets, k, rank=0, world=1, exchange=None):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        L = lib()
        if exchange is None:
            h = L.hemu_run(bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1, k)
        else:
            h = L.hemu_prepare(bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1, k)
            L.hemu_phase_a(h, rank, world)
            p = [C.c_void_p() for _ in range(4)]
            n = int(L.hemu_phase_a_arrays(h, *(C.byref(x) for x in p)))
            if n:
                mk = lambda ptr, ct, dt: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).view(dt)
                views = {"right": mk(p[0], C.c_uint64, np.int64), "left": mk(p[1], C.c_uint64, np.int64),
                         "over_limit": mk(p[2], C.c_uint8, np.uint8), "contained_by": mk(p[3], C.c_uint32, np.int32)}
                exchange(views, n // world)
            L.hemu_finish(h)
        sz = np.zeros(14, dtype=np.uint64)
        L.hemu_sizes(h, sz.ctypes.data)
        (self.U, self.SW, self.N, self.total_bp, self.n_edges, self.over, self.distinct, self.compare_calls,
         self.inserted, self.removed, self.contained, self.contained_size, self.slow_reads, self.fast_reads) = (int(x) for x in sz)
        U, SW, E = self.U, self.SW, self.n_edges
        self.F = np.zeros(U * SW, np.uint64); self.RC = np.zeros(U * SW, np.uint64)
        self.len = np.zeros(U, np.uint16); self.freq = np.zeros(U, np.uint16)
        self.extR = np.zeros(U, np.uint64); self.extL = np.zeros(U, np.uint64)
        self.explored_a = np.zeros(U, np.uint8); self.explored_b = np.zeros(U, np.uint8)
        self.edges = np.zeros(2 * E, np.uint64)
        L.hemu_copy(h, *(a.ctypes.data for a in (self.F, self.RC, self.len, self.freq, self.extR, self.extL,
                                                  self.explored_a, self.explored_b, self.edges)))
        L.hemu_free(h)
        self.F = self.F.reshape(U, SW); self.RC = self.RC.reshape(U, SW)


class EmuShard:
    """One rank of the sharded-table build on the CPU: the methods sage2_b200/multi.py's sharded_graph_steps calls on
    api.Sage2Gpu, with host buffers (tests/host_emul.cpp restates csrc/shard.cu)."""

    _registry = {}

    def __init__(self, bases, offsets, k):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.L = lib()
        self.h = self.L.hemu_prepare(bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1, k)
        self.redone = 0

    def close(self):
        if self.h:
            self.L.hemu_free(self.h)
            self.h = None

    def _sizes(self):
        sz = np.zeros(14, dtype=np.uint64)
        self.L.hemu_sizes(self.h, sz.ctypes.data)
        return [int(x) for x in sz]

    def counters(self):
        sz = self._sizes()
        return {"unique_reads": sz[0], "n_edges": sz[4], "compare_calls": sz[7], "distinct_keys": sz[6], "probe_restarts": self.redone}

    def build_hash_table_shard(self, rank, world):
        self.L.hemu_shard_build(self.h, rank, world)

    def phase_a_sharded_begin(self, rank, world):
        first, count = C.c_uint64(), C.c_uint64()
        self.L.hemu_phase_a_sharded_begin(self.h, rank, world, C.byref(first), C.byref(count))
        self.world = world
        return int(first.value), int(count.value)

    def route_begin(self, what, first, count, exact, world):
        q, n = C.c_void_p(), C.c_uint64()
        counts = (C.c_uint64 * world)()
        self.L.hemu_route_begin(self.h, what, first, count, int(bool(exact)), world, C.byref(q), counts, C.byref(n))
        return {"ptr": q.value or 0, "counts": [int(x) for x in counts], "words": 2 if exact else 1, "n_reads": int(n.value)}

    def shard_answer(self, queries_ptr, counts_per_source, exact, world):
        cps = (C.c_uint64 * world)(*[int(x) for x in counts_per_source])
        resp, ent = C.c_void_p(), C.c_void_p()
        ecnt = (C.c_uint64 * world)()
        self.L.hemu_shard_answer(self.h, queries_ptr or None, cps, int(bool(exact)), world, C.byref(resp), C.byref(ent), ecnt)
        return {"resp": resp.value or 0, "entries": ent.value or 0, "entry_counts": [int(x) for x in ecnt]}

    def route_finish(self, resp_ptr, entries_ptr, entry_counts):
        ec = (C.c_uint64 * len(entry_counts))(*[int(x) for x in entry_counts])
        self.L.hemu_route_finish(self.h, resp_ptr or None, entries_ptr or None, ec)

    # the peer-memory transport of multi.sharded_graph_steps(p2p=True), in ONE process: a "mailbox" is the peer object
    def mailbox_create(self, rank, world, max_reads_per_batch):
        self.rank, self.world, self.peers, self.mailbox_batch_reads = rank, world, {}, max_reads_per_batch
        EmuShard._registry[id(self)] = self
        return {"handle": None, "ptr": id(self)}

    def mailbox_open(self, peer_rank, handle=None, ptr=0):
        self.peers[peer_rank] = EmuShard._registry[ptr]

    def route_post(self, what, first, count, exact):
        rb = self.route_begin(what, first, count, exact, self.world)
        q = np.ctypeslib.as_array(C.cast(rb["ptr"], C.POINTER(C.c_uint64)), shape=(max(1, sum(rb["counts"]) * rb["words"]),)).copy() \
            if sum(rb["counts"]) else np.zeros(0, np.uint64)
        self.posted, o = {}, 0
        for g, c in enumerate(rb["counts"]):
            self.posted[g] = q[o:o + c * rb["words"]]
            o += c * rb["words"]
        self.inbox = {}
        return rb["n_reads"], int(sum(c for g, c in enumerate(rb["counts"]) if g != self.rank)) * 8 * rb["words"]

    def answer_post(self, exact):
        peers = dict(self.peers)
        peers[self.rank] = self
        words = 2 if exact else 1
        segs = [peers[s].posted[self.rank] for s in range(self.world)]
        counts = [len(x) // words for x in segs]
        q = np.ascontiguousarray(np.concatenate(segs)) if sum(counts) else np.zeros(1, np.uint64)
        ans = self.shard_answer(q.ctypes.data, counts, exact, self.world)
        resp = np.ctypeslib.as_array(C.cast(ans["resp"], C.POINTER(C.c_uint64)), shape=(max(1, sum(counts)),)).copy() if sum(counts) else np.zeros(0, np.uint64)
        ne = sum(ans["entry_counts"])
        ent = np.ctypeslib.as_array(C.cast(ans["entries"], C.POINTER(C.c_uint32)), shape=(max(1, ne),)).copy() if ne else np.zeros(0, np.uint32)
        o = eo = sent = 0
        for s in range(self.world):
            peers[s].inbox[self.rank] = (resp[o:o + counts[s]], ent[eo:eo + ans["entry_counts"][s]])
            if s != self.rank:
                sent += counts[s] * 8 + ans["entry_counts"][s] * 4
            o += counts[s]
            eo += ans["entry_counts"][s]
        return sent

    def route_collect(self):
        resp = np.ascontiguousarray(np.concatenate([self.inbox[g][0] for g in range(self.world)]).astype(np.uint64))
        ent = np.ascontiguousarray(np.concatenate([self.inbox[g][1] for g in range(self.world)]).astype(np.uint32))
        ecs = [len(self.inbox[g][1]) for g in range(self.world)]
        self.route_finish(resp.ctypes.data if len(resp) else 0, ent.ctypes.data if len(ent) else 0, ecs)

    def phase_a_routed(self):
        n = int(self.L.hemu_phase_a_routed(self.h))
        self.redone += n
        return n

    def phase_a_sharded_end(self):
        pass

    def phase_a_buffers(self):
        p = [C.c_void_p() for _ in range(4)]
        n = int(self.L.hemu_phase_a_arrays(self.h, *(C.byref(x) for x in p)))
        return {"right": p[0].value or 0, "left": p[1].value or 0, "over_limit": p[2].value or 0, "contained_by": p[3].value or 0,
                "chunk": n // self.world, "unique_reads": self._sizes()[0]}

    def phase_b(self):
        self.L.hemu_phase_b(self.h)

    def finish_graph(self):
        self.L.hemu_finish(self.h)

    def edges(self):
        E = self._sizes()[4]
        w = np.zeros(2 * E, np.uint64)
        self.L.hemu_copy(self.h, None, None, None, None, None, None, None, None, w.ctypes.data)
        return w

    def extensions(self):
        U = self._sizes()[0]
        r, l = np.zeros(U, np.uint64), np.zeros(U, np.uint64)
        self.L.hemu_copy(self.h, None, None, None, None, r.ctypes.data, l.ctypes.data, None, None, None)
        return r, l


def host_phase_a_views(bufs, world):
    """torch views (sharing memory) of an EmuShard's padded phase-A arrays, as multi.exchange_phase_a wants them."""
    from sage2_b200 import multi
    n = bufs["chunk"] * world
    return {"right": multi.host_view(bufs["right"], n, "<i8"), "left": multi.host_view(bufs["left"], n, "<i8"),
            "over_limit": multi.host_view(bufs["over_limit"], n, "|u1"), "contained_by": multi.host_view(bufs["contained_by"], n, "<i4")}


def records_to_bytes(rec: np.ndarray, lens: np.ndarray) -> np.ndarray:
    """(U,SW) word-big-endian records -> concatenated reference-layout bytes (utils.cpp:96-119)."""
    if len(rec) == 0:
        return np.zeros(0, np.uint8)
    be = rec.astype(">u8").view(np.uint8).reshape(len(rec), -1)
    nb = (lens.astype(np.int64) + 3) // 4
    mask = np.arange(be.shape[1])[None, :] < nb[:, None]
    return be[mask]


def unpack_ext(e: np.ndarray):
    return (e & np.uint64(0xFFFFFFFF)).astype(np.uint64), ((e >> np.uint64(32)) & np.uint64(1)).astype(np.uint32), \
        ((e >> np.uint64(33)) & np.uint64(0x3FFFFF)).astype(np.uint32)


def unpack_edges(w: np.ndarray):
    w0, w1 = w[0::2], w[1::2]
    return (w0 >> np.uint64(32)).astype(np.uint64), (w0 & np.uint64(0xFFFFFFFF)).astype(np.uint64), \
        ((w1 >> np.uint64(20)) & np.uint64(3)).astype(np.uint32), (w1 & np.uint64(0xFFFFF)).astype(np.uint32)


class EmuPart:
    """One rank of the fully partitioned build on the CPU: the methods multi.partitioned_graph_steps calls on
    api.Sage2Gpu, with host buffers (tests/host_emul.cpp restates reads.cu's key-range organise; the table is rebuilt
    from the gathered reads, so its two all-gathers carry nothing here)."""

    def __init__(self):
        self.L = lib()
        self.L.hemu_prepare_partition.restype = C.c_void_p
        self.L.hemu_prepare_partition.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        self.L.hemu_reads_gather_layout.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                                    C.POINTER(C.c_void_p)]
        self.L.hemu_reads_gather_finish.argtypes = [C.c_void_p]
        self.h = None

    def load_reads_partition(self, bases_ptr, offsets_ptr, n_reads, k, on_device, rank, world):
        u = C.c_uint64()
        self.rank, self.world = rank, world
        self.h = self.L.hemu_prepare_partition(bases_ptr, offsets_ptr, n_reads, k, rank, world, C.byref(u))
        return int(u.value)

    def _sizes(self):
        sz = np.zeros(14, dtype=np.uint64)
        self.L.hemu_sizes(self.h, sz.ctypes.data)
        return [int(x) for x in sz]

    def reads_gather_layout(self, counts):
        cs = (C.c_uint64 * len(counts))(*counts)
        p = [C.c_void_p() for _ in range(3)]
        self.L.hemu_reads_gather_layout(self.h, cs, self.rank, self.world, *(C.byref(x) for x in p))
        return {"records": p[0].value or 0, "lengths": p[1].value or 0, "frequencies": p[2].value or 0, "first": sum(counts[:self.rank]),
                "total": sum(counts), "stride": self._sizes()[1], "counts": list(counts)}

    def reads_gather_finish(self):
        self.L.hemu_reads_gather_finish(self.h)

    def build_hash_table_part(self, rank, world):
        pass

    def table_shard_info(self):
        sz = self._sizes()
        return {"slots": 0, "entries": 0, "distinct": sz[6], "over": sz[5]}

    def table_gather_layout(self, entry_counts):
        return {"slots": 0, "entries": 0, "slots_per_shard": 0, "entries_first": 0}

    def table_gather_finish(self, entry_counts, distinct, over):
        pass

    def phase_a_partition(self, rank, world):
        self.L.hemu_phase_a(self.h, rank, world)

    def phase_a_buffers(self):
        p = [C.c_void_p() for _ in range(4)]
        n = int(self.L.hemu_phase_a_arrays(self.h, *(C.byref(x) for x in p)))
        return {"right": p[0].value or 0, "left": p[1].value or 0, "over_limit": p[2].value or 0, "contained_by": p[3].value or 0,
                "chunk": n // self.world, "unique_reads": self._sizes()[0]}

    def finish_graph(self):
        self.L.hemu_finish(self.h)

    def result(self):
        sz = self._sizes()
        U, SW, E = sz[0], sz[1], sz[4]
        F = np.zeros(U * SW, np.uint64); RC = np.zeros(U * SW, np.uint64)
        ln = np.zeros(U, np.uint16); fr = np.zeros(U, np.uint16)
        eR = np.zeros(U, np.uint64); eL = np.zeros(U, np.uint64)
        xa = np.zeros(U, np.uint8); xb = np.zeros(U, np.uint8)
        edges = np.zeros(2 * E, np.uint64)
        self.L.hemu_copy(self.h, *(a.ctypes.data for a in (F, RC, ln, fr, eR, eL, xa, xb, edges)))
        return {"U": U, "len": ln, "freq": fr, "F": F.reshape(U, SW), "RC": RC.reshape(U, SW), "edges": edges, "distinct": sz[6], "over": sz[5],
                "compare_calls": sz[7]}

    def close(self):
        if self.h:
            self.L.hemu_free(self.h)
            self.h = None
