"""Multi-GPU overlap-graph build: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch)
for the plumbing.

Round-1 partitioning (the replicated-table variant): every rank organises the reads
and builds the table (both deterministic, so all ranks hold identical copies), rank r searches read ids
[r*chunk, (r+1)*chunk) (sage2gpu_phase_a_partition), then ONE exchange step makes the phase-A state
complete everywhere:

    all-gather        rightExtension / leftExtension records, connections>300 flags
    all-reduce(MAX)   largest id whose scan found the read contained (economyGraph.cpp:735)

after which phases B and C and the canonical edge sort run (sage2gpu_finish_graph).  The exchange
works on any torch tensors, so the world_size-2 gloo test drives it on the CPU with emulated slices.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class _DevArray:
    """Zero-copy view of library-owned device memory (CUDA array interface v2)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_views(bufs: dict, world: int, device) -> dict:
    """torch views of the arrays sage2gpu_phase_a_buffers exposes (int64 / uint8 / int32 bit patterns)."""
    n = bufs["chunk"] * world
    if n == 0:
        z = lambda dt: torch.zeros(0, dtype=dt, device=device)
        return {"right": z(torch.int64), "left": z(torch.int64), "over_limit": z(torch.uint8), "contained_by": z(torch.int32)}
    mk = lambda key, ts: torch.as_tensor(_DevArray(bufs[key], n, ts), device=device)
    return {"right": mk("right", "<i8"), "left": mk("left", "<i8"), "over_limit": mk("over_limit", "|u1"),
            "contained_by": mk("contained_by", "<i4")}


def exchange_phase_a(views: dict, chunk: int, rank: int, world: int) -> int:
    """In place: every rank ends with the complete phase-A arrays.  Returns the bytes this rank sent."""
    if world == 1 or chunk == 0:
        return 0
    sent = 0
    for key in ("right", "left", "over_limit"):
        full = views[key]
        mine = full[rank * chunk:(rank + 1) * chunk]
        if full.is_cuda:
            dist.all_gather_into_tensor(full, mine)
        else:      # gloo
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine.clone())
            full.copy_(torch.cat(parts))
        sent += mine.numel() * mine.element_size()
    dist.all_reduce(views["contained_by"], op=dist.ReduceOp.MAX)
    sent += views["contained_by"].numel() * 4
    return sent


def build_overlap_graph(gpu, rank: int, world: int, device=None) -> int:
    """sage2gpu_build_overlap_graph over `world` GPUs (reads loaded and table built on every rank)."""
    gpu.phase_a_partition(rank, world)
    if world == 1:
        gpu.finish_graph()
        return 0
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    bufs = gpu.phase_a_buffers()
    views = device_views(bufs, world, device)
    sent = exchange_phase_a(views, bufs["chunk"], rank, world)      # phase_a_partition returned synchronised
    torch.cuda.current_stream(device).synchronize()
    gpu.finish_graph()
    return sent


def build_partitioned(gpu, rank: int, world: int, device, bases_ptr: int, offsets_ptr: int, n_reads: int, k: int, on_device: bool,
                      stats: dict | None = None) -> int:
    """One process per GPU (torchrun): steps 1-3 with every stage partitioned; returns the bytes this rank contributed."""
    if world == 1:
        gpu.load_reads_ptr(bases_ptr, offsets_ptr, n_reads, k, on_device)
        gpu.build_hash_table()
        gpu.build_overlap_graph()
        return 0
    sent = [0]
    steps = partitioned_graph_steps(gpu, rank, world, device_view_fn(device), bases_ptr, offsets_ptr, n_reads, k, on_device, sent)
    run_dist(steps, rank, world, device, stats)
    return sent[0]


def build_partitioned_slices(gpu, rank: int, world: int, device, bases_ptr: int, offsets_ptr: int, n_slice: int, k: int, on_device: bool,
                             max_read_length: int, stats: dict | None = None) -> int:
    """One process per GPU: steps 1-3 from THIS RANK'S SLICE of the input (host or device buffers): the slice is packed
    here, the packed records are all-gathered, every later stage is partitioned.  Returns the bytes this rank contributed."""
    sent = [0]
    steps = partitioned_slice_steps(gpu, rank, world, device_view_fn(device), bases_ptr, offsets_ptr, n_slice, k, on_device, max_read_length, sent)
    if world == 1:
        run_local([steps])
    else:
        run_dist(steps, rank, world, device, stats)
    return sent[0]


def upload_partitioned(h_bases: torch.Tensor, h_offsets: torch.Tensor, rank: int, world: int, device) -> tuple:
    """Host -> device of the raw reads with every rank moving only its 1/world share over PCIe; the shares are
    then all-gathered over NVLink.  h_* are (pinned) host tensors holding the WHOLE input on every rank.
    Returns (d_bases uint8, d_offsets int64, bytes copied host->device by this rank)."""
    if world == 1:
        return h_bases.to(device, non_blocking=True), h_offsets.to(device, non_blocking=True), h_bases.numel() + 8 * h_offsets.numel()
    out, moved = [], 0
    for h in (h_bases, h_offsets):
        n = h.numel()
        share = -(-n // world)
        full = torch.empty(share * world, dtype=h.dtype, device=device)
        lo, hi = min(n, rank * share), min(n, (rank + 1) * share)
        mine = full[rank * share:(rank + 1) * share]
        if hi > lo:
            mine[:hi - lo].copy_(h[lo:hi], non_blocking=True)
        moved += (hi - lo) * h.element_size()
        dist.all_gather_into_tensor(full, mine)
        out.append(full[:n])
    return out[0], out[1], moved


# ---------------------------------------------------------------------------------------------------------------
# The table sharded by key hash (north_star): shard r holds the keys it owns, the window probes of
# a batch of reads travel to their owners and the answers travel back -- NCCL all-to-all over NVLink / NVSwitch.
# The library (include/sage2gpu.h, csrc/shard.cu) fills and consumes the device buffers; this module moves them.
#
# The steps of one rank are written once, as a generator that yields its exchange requests:
#     ("counts", send_counts)                     -> counts received from every rank       (all-to-all of one int each)
#     ("a2a", tensor, send_counts, recv_counts)   -> the elements received, by source rank (all-to-all, variable splits)
#     ("max", value)                              -> the maximum over the ranks
#     ("phase_a", bufs)                           -> exchange_phase_a on the arrays of sage2gpu_phase_a_buffers
# run_dist() serves them with torch.distributed (one process per GPU); run_local() drives several contexts of ONE
# process in lockstep (tests on one GPU, and the CPU emulation).
# ---------------------------------------------------------------------------------------------------------------

def host_view(ptr: int, n: int, typestr: str) -> torch.Tensor:
    """Zero-copy torch view of host memory (the CPU emulation's buffers)."""
    import ctypes
    import numpy as np
    dt = np.dtype(typestr)
    if n == 0 or not ptr:
        return torch.zeros(0, dtype=torch.from_numpy(np.zeros(0, dt)).dtype)
    buf = (ctypes.c_char * (n * dt.itemsize)).from_address(ptr)
    return torch.from_numpy(np.frombuffer(buf, dtype=dt, count=n))


def device_view_fn(device):
    def view(ptr: int, n: int, typestr: str) -> torch.Tensor:
        if n == 0 or not ptr:
            return torch.zeros(0, dtype={"<i8": torch.int64, "<i4": torch.int32, "<i2": torch.int16, "|u1": torch.uint8}[typestr], device=device)
        return torch.as_tensor(_DevArray(ptr, n, typestr), device=device)
    return view


def _ptr(t: torch.Tensor) -> int:
    return t.data_ptr() if t.numel() else 0


def _route(gpu, what: int, first: int, count: int, exact: bool, world: int, view, begun=None):
    """One routed batch: queries to their owners, answers and bucket entries back (csrc/shard.cu)."""
    rb = begun if begun is not None else gpu.route_begin(what, first, count, exact, world)
    qw, send_counts = rb["words"], rb["counts"]
    recv_counts = yield ("counts", send_counts)
    queries = view(rb["ptr"], sum(send_counts) * qw, "<i8")
    recv_q = yield ("a2a", queries, [c * qw for c in send_counts], [c * qw for c in recv_counts])
    ans = gpu.shard_answer(_ptr(recv_q), recv_counts, exact, world)
    recv_resp = yield ("a2a", view(ans["resp"], sum(recv_counts), "<i8"), recv_counts, send_counts)
    ecs = ans["entry_counts"]
    recv_ecs = yield ("counts", ecs)
    recv_ent = yield ("a2a", view(ans["entries"], sum(ecs), "<i4"), ecs, recv_ecs)
    gpu.route_finish(_ptr(recv_resp), _ptr(recv_ent), recv_ecs)


def _route_p2p(gpu, what: int, first: int, count: int, exact: bool, sent: list, posted=None):
    """One routed batch over peer memory: the kernels store / copy into the other ranks' mailboxes, the host only
    separates the three steps with barriers (csrc/shard.cu, second half)."""
    if posted is None:
        posted = gpu.route_post(what, first, count, exact)
    sent[0] += posted[1]
    yield ("device_barrier", gpu)
    sent[0] += gpu.answer_post(exact)
    yield ("device_barrier", gpu)
    gpu.route_collect()


def mailbox_steps(gpu, rank: int, world: int, batch_reads: int):
    """Generator: create this rank's mailbox and map everybody else's (once per context; the reads must be loaded)."""
    mine = gpu.mailbox_create(rank, world, batch_reads)
    boxes = yield ("mailboxes", mine)
    for r, b in enumerate(boxes):
        if r != rank:
            gpu.mailbox_open(r, handle=b.get("handle"), ptr=b.get("ptr", 0))
    yield ("barrier",)


def sharded_graph_steps(gpu, rank: int, world: int, view, batch_reads: int = 1 << 20, p2p: bool = False, sent: list | None = None):
    """Generator: sage2gpu_build_overlap_graph with the table sharded over `world` ranks.  Reads loaded on every rank
    and sage2gpu_build_hash_table_shard(rank, world) done.  p2p: exchange through the ranks' mailboxes (mailbox_steps
    done, batch_reads <= the mailbox's) instead of all-to-all requests; sent[0] accumulates the bytes stored remotely."""
    sent = sent if sent is not None else [0]

    def route(what, first, count, exact, begun=None):      # one batch through the chosen transport
        if p2p:
            return _route_p2p(gpu, what, first, count, exact, sent, posted=begun)
        return _route(gpu, what, first, count, exact, world, view, begun=begun)

    first, count = gpu.phase_a_sharded_begin(rank, world)
    U = gpu.counters()["unique_reads"]
    chunk = -(-U // world) if U else 0
    n_batches = -(-chunk // batch_reads) if chunk else 0          # the same on every rank: collectives stay in lockstep
    redo = 0
    for b in range(n_batches):
        lo = min(first + b * batch_reads, first + count)
        n = min(batch_reads, first + count - lo)
        yield from route(0, lo, n, False)
        redo += gpu.phase_a_routed()
    def fit_mailbox(n_list):
        # a list batch (redo reads, reads left for phase C) travels as ONE batch: the mailboxes of all ranks grow to hold
        # the largest list any rank has (error-rich data sends most reads through phase C)
        need = yield ("max", n_list)
        if p2p and need > getattr(gpu, "mailbox_batch_reads", 0):
            yield from mailbox_steps(gpu, rank, world, need)
        return need

    if (yield ("max", redo)):
        # 24-bit tag collisions (about U*W / 2^24 reads): those reads once more, with probes the owners verify
        yield from fit_mailbox(redo)
        yield from route(2, 0, 0, True)
        left = gpu.phase_a_routed()
        if left:
            raise RuntimeError(f"{left} reads still unresolved after the verified pass")
    gpu.phase_a_sharded_end()
    yield ("phase_a", gpu.phase_a_buffers())
    gpu.phase_b()
    cnt = gpu.counters()
    yield from fit_mailbox(cnt.get("left_to_explore", cnt.get("unique_reads", 0)))
    if p2p:
        posted = gpu.route_post(1, 0, 0, True)          # reads left for phase C: identical list on every rank
        if posted[0]:
            yield from route(1, 0, 0, True, begun=posted)
    else:
        rb = gpu.route_begin(1, 0, 0, True, world)
        if rb["n_reads"]:
            yield from route(1, 0, 0, True, begun=rb)
    gpu.finish_graph()


# ---------------------------------------------------------------------------------------------------------------
# Every stage partitioned, results replicated by all-gathers (include/sage2gpu.h "Several GPUs, every stage partitioned"):
# rank r sorts + dedupes the reads of its key range, builds the table shard of the keys it owns and searches its slice of
# the read ids; the unique reads, the table shards and the phase-A arrays are all-gathered over NVLink.  Two more requests:
#     ("gather_counts", values)            -> the lists of all ranks, by rank
#     ("gather_var", full, counts)         -> in place: rank q's block of `full` (counts[q] elements, blocks back to back)
#                                             is filled from rank q; this rank's block is already in place
# ---------------------------------------------------------------------------------------------------------------

def _partitioned_from_organized(gpu, rank: int, world: int, view, u_local: int, sent: list):
    """The steps after this rank has organised the reads of its key range (u_local unique reads)."""
    counts = [c[0] for c in (yield ("gather_counts", [u_local]))]
    lay = gpu.reads_gather_layout(counts)
    tot, stride = lay["total"], lay["stride"]
    yield ("gather_var", view(lay["records"], tot * stride, "<i8"), [c * stride for c in counts])
    yield ("gather_var", view(lay["lengths"], tot, "<i2"), counts)
    yield ("gather_var", view(lay["frequencies"], tot, "<i2"), counts)
    sent[0] += counts[rank] * (8 * stride + 4)
    gpu.reads_gather_finish()
    gpu.build_hash_table_part(rank, world)
    info = gpu.table_shard_info()
    infos = yield ("gather_counts", [info["entries"], info["distinct"], info["over"]])
    ec = [x[0] for x in infos]
    tl = gpu.table_gather_layout(ec)
    yield ("gather_var", view(tl["slots"], tl["slots_per_shard"] * world, "<i8"), [tl["slots_per_shard"]] * world)
    yield ("gather_var", view(tl["entries"], sum(ec), "<i4"), ec)
    sent[0] += 8 * tl["slots_per_shard"] + 4 * ec[rank]
    gpu.table_gather_finish(ec, [x[1] for x in infos], [x[2] for x in infos])
    gpu.phase_a_partition(rank, world)
    yield ("phase_a", gpu.phase_a_buffers())
    gpu.finish_graph()


def partitioned_graph_steps(gpu, rank: int, world: int, view, bases_ptr: int, offsets_ptr: int, n_reads: int, k: int, on_device: bool,
                            sent: list | None = None):
    """Generator: steps 1-3 over `world` ranks; every rank holds the whole input (replicated, or all-gathered before)."""
    sent = sent if sent is not None else [0]
    u_local = gpu.load_reads_partition(bases_ptr, offsets_ptr, n_reads, k, on_device, rank, world)
    yield from _partitioned_from_organized(gpu, rank, world, view, u_local, sent)


def partitioned_slice_steps(gpu, rank: int, world: int, view, bases_ptr: int, offsets_ptr: int, n_slice: int, k: int, on_device: bool,
                            max_read_length: int, sent: list | None = None):
    """Generator: the same with the ingest partitioned too -- every rank holds (and packs) only its slice of the input, the
    packed records are all-gathered (4 x fewer bytes than the characters, and no rank ever holds all characters)."""
    sent = sent if sent is not None else [0]
    info = gpu.pack_slice(bases_ptr, offsets_ptr, n_slice, k, on_device, max_read_length)
    infos = yield ("gather_counts", [n_slice, info["good_reads"], info["total_bp"]])
    counts = [x[0] for x in infos]
    lay = gpu.raw_gather_layout(rank, world, counts)
    sw = info["record_words"]
    yield ("gather_var", view(lay["records"], lay["total"] * sw, "<i8"), [c * sw for c in counts])
    sent[0] += 8 * sw * n_slice
    gpu.raw_gather_finish(sum(counts), sum(x[1] for x in infos), sum(x[2] for x in infos))
    u_local = gpu.organize_partition(rank, world)
    if world == 1:
        gpu.build_hash_table()
        gpu.build_overlap_graph()
        return
    yield from _partitioned_from_organized(gpu, rank, world, view, u_local, sent)


def partitioned_slice_sharded_steps(gpu, rank: int, world: int, view, bases_ptr: int, offsets_ptr: int, n_slice: int, k: int, on_device: bool,
                                    max_read_length: int, batch_reads: int = 1 << 18, p2p: bool = True, sent: list | None = None):
    """Generator: north_star's layout at full scale -- ingest and read organisation partitioned (as partitioned_slice_steps),
    the packed reads replicated by all-gather, the TABLE sharded by key hash with the window probes routed to their owners
    (sharded_graph_steps): for read sets whose table should not be held on every GPU (config #5: 29 GB)."""
    sent = sent if sent is not None else [0]
    info = gpu.pack_slice(bases_ptr, offsets_ptr, n_slice, k, on_device, max_read_length)
    infos = yield ("gather_counts", [n_slice, info["good_reads"], info["total_bp"]])
    counts = [x[0] for x in infos]
    lay = gpu.raw_gather_layout(rank, world, counts)
    sw = info["record_words"]
    yield ("gather_var", view(lay["records"], lay["total"] * sw, "<i8"), [c * sw for c in counts])
    sent[0] += 8 * sw * n_slice
    gpu.raw_gather_finish(sum(counts), sum(x[1] for x in infos), sum(x[2] for x in infos))
    u_local = gpu.organize_partition(rank, world)
    if world > 1:
        ucounts = [c[0] for c in (yield ("gather_counts", [u_local]))]
        rl = gpu.reads_gather_layout(ucounts)
        tot, stride = rl["total"], rl["stride"]
        yield ("gather_var", view(rl["records"], tot * stride, "<i8"), [c * stride for c in ucounts])
        yield ("gather_var", view(rl["lengths"], tot, "<i2"), ucounts)
        yield ("gather_var", view(rl["frequencies"], tot, "<i2"), ucounts)
        sent[0] += ucounts[rank] * (8 * stride + 4)
        gpu.reads_gather_finish()
    gpu.build_hash_table_shard(rank, world)
    if p2p and getattr(gpu, "mailbox_batch_reads", 0) < batch_reads:
        yield from mailbox_steps(gpu, rank, world, batch_reads)
    yield from sharded_graph_steps(gpu, rank, world, view, batch_reads, p2p=p2p, sent=sent)


def _gather_var_dist(full: torch.Tensor, counts, rank: int, world: int):
    """All-gather of blocks of different sizes: equal blocks in place, otherwise through a padded copy."""
    if full.dtype != torch.uint8:      # bytes travel (neither NCCL nor gloo has a 16-bit integer type)
        es = full.element_size()
        return _gather_var_dist(full.view(torch.uint8), [c * es for c in counts], rank, world)
    if len(set(counts)) == 1 and full.is_cuda:
        c = counts[0]
        if c:
            dist.all_gather_into_tensor(full, full[rank * c:(rank + 1) * c])
        return
    if max(counts) == 0:
        return
    # blocks of different sizes: grouped sends / receives straight between the blocks (no padded copy: at config #5 the
    # records of the unique reads alone are 35 GB)
    offs = [sum(counts[:q]) for q in range(world)]
    mine = full[offs[rank]:offs[rank] + counts[rank]]
    ops = []
    for q in range(world):
        if q == rank:
            continue
        if counts[rank]:
            ops.append(dist.P2POp(dist.isend, mine, q))
        if counts[q]:
            ops.append(dist.P2POp(dist.irecv, full[offs[q]:offs[q] + counts[q]], q))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def serve_one(req, rank: int, world: int, device, sent: list | None = None):
    """Serve ONE exchange request of a step generator with torch.distributed (NCCL for CUDA tensors, gloo on the CPU)."""
    sent = sent if sent is not None else [0]
    cuda = torch.device(device).type == "cuda"
    kind = req[0]
    if kind == "counts":
        t = torch.tensor(req[1], dtype=torch.int64, device=device)
        out = torch.empty_like(t)
        dist.all_to_all_single(out, t)
        sent[0] += 8 * (world - 1)
        return out.tolist()
    if kind == "a2a":
        send, sc, rc = req[1], req[2], req[3]
        out = torch.empty(sum(rc), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(out, send, rc, sc)
        if cuda:
            torch.cuda.current_stream(device).synchronize()       # the library reads it on its own stream
        sent[0] += (sum(sc) - sc[rank]) * send.element_size()
        return out
    if kind == "max":
        t = torch.tensor([req[1]], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t.item())
    if kind == "gather_counts":
        t = torch.tensor(req[1], dtype=torch.int64, device=device)
        if cuda:
            out = torch.empty(world * t.numel(), dtype=torch.int64, device=device)
            dist.all_gather_into_tensor(out, t)
        else:
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            out = torch.cat(parts)
        return out.view(world, -1).tolist()
    if kind == "gather_var":
        _gather_var_dist(req[1], req[2], rank, world)
        if cuda:
            torch.cuda.current_stream(device).synchronize()       # the library reads it on its own stream
        sent[0] += req[2][rank] * req[1].element_size()
        return None
    if kind == "barrier":           # every rank's library call before it has returned (= its stream is drained)
        t = torch.zeros(1, dtype=torch.int32, device=device)
        dist.all_reduce(t)
        return int(t.item())
    if kind == "device_barrier":    # flags in the peers' mailboxes, written and awaited by a kernel (csrc/shard.cu)
        req[1].mailbox_barrier()
        return 0
    if kind == "mailboxes":         # CUDA IPC handles of all ranks (other processes cannot use the pointer)
        boxes = [None] * world
        dist.all_gather_object(boxes, {"handle": req[1]["handle"]})
        return boxes
    if kind == "phase_a":
        bufs = req[1]
        views = req[2] if len(req) > 2 else device_views(bufs, world, device)
        sent[0] += exchange_phase_a(views, bufs["chunk"], rank, world)
        if cuda:
            torch.cuda.current_stream(device).synchronize()
        return None
    raise ValueError(kind)


def run_dist(gen, rank: int, world: int, device, stats: dict | None = None) -> int:
    """Serve one rank's requests with torch.distributed (NCCL for CUDA tensors, gloo on the CPU).  Returns bytes sent.
    stats (optional): wall milliseconds spent in the exchanges, by request kind, are added to it."""
    import time
    sent = [0]
    try:
        req = next(gen)
        while True:
            t0 = time.perf_counter()
            val = serve_one(req, rank, world, device, sent)
            if stats is not None:
                kind = req[0]
                stats[kind] = stats.get(kind, 0.0) + (time.perf_counter() - t0) * 1e3
                stats["n_" + kind] = stats.get("n_" + kind, 0) + 1
            req = gen.send(val)
    except StopIteration:
        pass
    return sent[0]


def run_local(gens: list, views_of=None) -> None:
    """Drive the generators of all ranks of ONE process in lockstep, doing the exchanges in memory.
    views_of(rank, bufs) -> the dict of torch views exchange_phase_a works on (default: device views on cuda:current)."""
    world = len(gens)
    reqs = [next(g) for g in gens]
    while True:
        kinds = {r[0] for r in reqs}
        assert len(kinds) == 1, f"ranks out of step: {kinds}"
        kind = kinds.pop()
        if kind == "counts":
            vals = [[reqs[s][1][r] for s in range(world)] for r in range(world)]
        elif kind == "a2a":
            vals = []
            for r in range(world):
                parts = []
                for s in range(world):
                    send, sc = reqs[s][1], reqs[s][2]
                    assert reqs[r][3][s] == sc[r]
                    o = sum(sc[:r])
                    parts.append(send[o:o + sc[r]])
                vals.append(torch.cat(parts).clone())
            if vals[0].is_cuda:
                torch.cuda.synchronize()
        elif kind == "max":
            vals = [max(r[1] for r in reqs)] * world
        elif kind == "gather_counts":
            vals = [[list(r[1]) for r in reqs]] * world
        elif kind == "gather_var":
            counts = reqs[0][2]
            offs = [sum(counts[:q]) for q in range(world)]
            for q in range(world):
                blk = reqs[q][1][offs[q]:offs[q] + counts[q]].clone()
                for r in range(world):
                    if r != q and counts[q]:
                        reqs[r][1][offs[q]:offs[q] + counts[q]] = blk
            if reqs[0][1].is_cuda:
                torch.cuda.synchronize()
            vals = [None] * world
        elif kind in ("barrier", "device_barrier"):      # one process drives the ranks in turn: nothing to wait for
            vals = [0] * world
        elif kind == "mailboxes":             # one process: plain pointers
            vals = [[{"ptr": r[1]["ptr"]} for r in reqs]] * world
        elif kind == "phase_a":
            views = []
            for r in range(world):
                bufs = reqs[r][1]
                views.append(views_of(r, bufs) if views_of else device_views(bufs, world, torch.device("cuda", torch.cuda.current_device())))
            chunk = reqs[0][1]["chunk"]
            if chunk:
                for key in ("right", "left", "over_limit"):
                    for s in range(world):
                        mine = views[s][key][s * chunk:(s + 1) * chunk].clone()
                        for r in range(world):
                            views[r][key][s * chunk:(s + 1) * chunk] = mine
                m = views[0]["contained_by"].clone()
                for r in range(1, world):
                    m = torch.maximum(m, views[r]["contained_by"])
                for r in range(world):
                    views[r]["contained_by"].copy_(m)
                if m.is_cuda:
                    torch.cuda.synchronize()
            vals = [None] * world
        else:
            raise ValueError(kind)
        nxt = []
        done = 0
        for g, v in zip(gens, vals):
            try:
                nxt.append(g.send(v))
            except StopIteration:
                done += 1
        if done:
            assert done == world, "ranks finished at different steps"
            return
        reqs = nxt


def build_overlap_graph_sharded(gpu, rank: int, world: int, device=None, batch_reads: int = 1 << 20, stats: dict | None = None,
                                p2p: bool = False) -> int:
    """One process per GPU (torchrun): the sharded-table build on this rank; returns the bytes this rank sent.
    p2p: the routed probes travel through peer-memory mailboxes (set up on first use) instead of NCCL all-to-all."""
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    drive = (lambda g: run_local([g])) if world == 1 else (lambda g: run_dist(g, rank, world, device, stats))
    sent = [0]
    if p2p and getattr(gpu, "mailbox_batch_reads", 0) < batch_reads:
        drive(mailbox_steps(gpu, rank, world, batch_reads))
    steps = sharded_graph_steps(gpu, rank, world, device_view_fn(device), batch_reads, p2p=p2p, sent=sent)
    other = drive(steps)
    return sent[0] + (other or 0)
