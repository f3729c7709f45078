"""N > 1 on the CPU (world_size 2 and 3, gloo): the partition arithmetic and the one exchange step of the
multi-GPU build (sage2_b200/multi.py: all-gather of the extension records / flags, all-reduce MAX of the
containment ids), driven with the CPU emulation's phase-A slices.  Every rank must end with the oracle's
edge list."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import datasets
import emul
from oracle import oracle
from sage2_b200 import multi, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        reads, k = datasets.get(name)
        b, off = synth.concat(reads)
        sent = []

        def exchange(views, chunk):
            tv = {key: torch.from_numpy(v) for key, v in views.items()}      # share memory with the emulator's arrays
            sent.append(multi.exchange_phase_a(tv, chunk, rank, world))

        e = emul.EmuRun(b, off, k, rank=rank, world=world, exchange=exchange)
        np.save(os.path.join(out, f"edges{rank}.npy"), e.edges)
        np.save(os.path.join(out, f"ext{rank}.npy"), np.stack([e.extR, e.extL]))
        np.save(os.path.join(out, f"sent{rank}.npy"), np.array(sent + [e.U]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("rep", 2), ("varlen_err", 2), ("deep_varlen", 3), ("single", 2)])
def test_partitioned_phase_a_with_exchange_equals_oracle(name, world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    reads, k = datasets.get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    for r in range(world):
        a, bb, t, d = emul.unpack_edges(np.load(tmp_path / f"edges{r}.npy"))
        assert len(a) == o.n_edges
        np.testing.assert_array_equal(a, o.edges["from"])
        np.testing.assert_array_equal(bb, o.edges["to"])
        np.testing.assert_array_equal(t, o.edges["type"])
        np.testing.assert_array_equal(d, o.edges["delta"])
        ext = np.load(tmp_path / f"ext{r}.npy")
        np.testing.assert_array_equal(emul.unpack_ext(ext[0])[0], o.right_ext["id"][1:])
        np.testing.assert_array_equal(emul.unpack_ext(ext[1])[0], o.left_ext["id"][1:])
        sent = np.load(tmp_path / f"sent{r}.npy")
        U = int(sent[-1])
        chunk = -(-U // world)
        assert int(sent[0]) == chunk * (8 + 8 + 1) + chunk * world * 4


def _upload_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        reads, _ = datasets.get("varlen")
        b, off = synth.concat(reads)
        tb, to, moved = multi.upload_partitioned(torch.from_numpy(b), torch.from_numpy(off), rank, world, torch.device("cpu"))
        np.save(os.path.join(out, f"up{rank}.npy"), np.array([int((tb.numpy() == b).all()), int((to.numpy() == off).all()), moved]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_upload_reassembles_the_input(world, tmp_path):
    """Every rank copies only its 1/world share of the input; the all-gather must give every rank the whole of it."""
    mp.spawn(_upload_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    reads, _ = datasets.get("varlen")
    b, off = synth.concat(reads)
    total = 0
    for r in range(world):
        ok_b, ok_o, moved = np.load(tmp_path / f"up{r}.npy")
        assert ok_b == 1 and ok_o == 1
        total += int(moved)
    assert total == b.nbytes + off.nbytes


def _requests_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        got = {}

        def steps():      # the request kinds of multi.run_dist that the sharded steps use besides all-to-all
            got["boxes"] = yield ("mailboxes", {"handle": bytes([rank]) * 64, "ptr": 1234 + rank})
            got["barrier"] = yield ("barrier",)
            got["max"] = yield ("max", 10 * rank)
            got["counts"] = yield ("counts", [rank * 100 + d for d in range(world)])
            got["a2a"] = yield ("a2a", torch.arange(world * 2, dtype=torch.int64) + 1000 * rank, [2] * world, [2] * world)

        stats = {}
        multi.run_dist(steps(), rank, world, "cpu", stats)
        assert [b["handle"][0] for b in got["boxes"]] == list(range(world)) and all("ptr" not in b for b in got["boxes"])
        assert got["max"] == 10 * (world - 1)
        assert got["counts"] == [s * 100 + rank for s in range(world)]
        assert got["a2a"].tolist() == [1000 * s + 2 * rank + i for s in range(world) for i in range(2)]
        assert stats["n_a2a"] == 1 and stats["n_barrier"] == 1
        open(os.path.join(out, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_exchange_requests_over_gloo(tmp_path):
    world = 3
    mp.spawn(_requests_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


# ---- every stage partitioned (multi.partitioned_graph_steps) over gloo ---------------------------------------------------

def _part_worker(rank, world, port, name, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        reads, k = datasets.get(name)
        b, off = synth.concat(reads)
        g = emul.EmuPart()
        views = {}

        def view(ptr, n, ts):
            return multi.host_view(ptr, n, ts)

        steps = multi.partitioned_graph_steps(g, rank, world, view, b.ctypes.data, off.ctypes.data, len(off) - 1, k, False)

        def serve():
            # run_dist with the phase-A exchange on host views of the emulator's arrays
            req = next(steps)
            while True:
                if req[0] == "phase_a":
                    bufs = req[1]
                    n = bufs["chunk"] * world
                    tv = {"right": multi.host_view(bufs["right"], n, "<i8"), "left": multi.host_view(bufs["left"], n, "<i8"),
                          "over_limit": multi.host_view(bufs["over_limit"], n, "|u1"), "contained_by": multi.host_view(bufs["contained_by"], n, "<i4")}
                    multi.exchange_phase_a(tv, bufs["chunk"], rank, world)
                    val = None
                else:
                    val = multi.serve_one(req, rank, world, "cpu")
                try:
                    req = steps.send(val)
                except StopIteration:
                    return
        serve()
        r = g.result()
        np.save(os.path.join(out, f"edges{rank}.npy"), r["edges"])
        np.save(os.path.join(out, f"freq{rank}.npy"), r["freq"])
        np.save(os.path.join(out, f"F{rank}.npy"), r["F"])
        np.save(os.path.join(out, f"calls{rank}.npy"), np.array([r["compare_calls"], r["U"]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("rep", 2), ("varlen_err", 3), ("hicopy", 2), ("single", 2), ("allbad", 2)])
def test_every_stage_partitioned_over_gloo(name, world, tmp_path):
    mp.spawn(_part_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    reads, k = datasets.get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    calls = 0
    for r in range(world):
        a, bb, t, d = emul.unpack_edges(np.load(tmp_path / f"edges{r}.npy"))
        assert len(a) == o.n_edges
        np.testing.assert_array_equal(a, o.edges["from"])
        np.testing.assert_array_equal(bb, o.edges["to"])
        np.testing.assert_array_equal(t, o.edges["type"])
        np.testing.assert_array_equal(d, o.edges["delta"])
        np.testing.assert_array_equal(np.load(tmp_path / f"freq{r}.npy"), o.frequency[1:])
        c = np.load(tmp_path / f"calls{r}.npy")
        assert c[1] == o.U
        calls += int(c[0])
        F = np.load(tmp_path / f"F{r}.npy")
        np.testing.assert_array_equal(emul.records_to_bytes(F, o.length[1:]), o.fwd)
    assert calls == o.compare_calls
