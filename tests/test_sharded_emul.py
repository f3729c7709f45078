"""The sharded-table build without a GPU: the steps of sage2_b200/multi.py (sharded_graph_steps: routed window
probes, redo pass, phase-A exchange, routed phase C) driven with the CPU emulation of csrc/shard.cu
(tests/host_emul.cpp, same buffers and wire format).  In one process with the in-memory exchanges (run_local), and
as world_size 2 / 3 gloo process groups with the real all-to-all plumbing (run_dist).  Every rank must end with the
oracle's edge list."""
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


def _check(edges_words, o):
    a, b, t, d = emul.unpack_edges(edges_words)
    assert len(a) == o.n_edges
    np.testing.assert_array_equal(a, o.edges["from"])
    np.testing.assert_array_equal(b, o.edges["to"])
    np.testing.assert_array_equal(t, o.edges["type"])
    np.testing.assert_array_equal(d, o.edges["delta"])


def _local(name, world, batch_reads=1 << 19):
    reads, k = datasets.get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    shards = []
    for r in range(world):
        s = emul.EmuShard(b, off, k)
        s.build_hash_table_shard(r, world)
        shards.append(s)
    multi.run_local([multi.sharded_graph_steps(s, r, world, multi.host_view, batch_reads) for r, s in enumerate(shards)],
                    views_of=lambda r, bufs: emul.host_phase_a_views(bufs, world))
    return o, shards


@pytest.mark.parametrize("name,world", [("clean", 1), ("rep", 2), ("hicopy", 3), ("varlen_err", 2), ("deep_varlen", 4),
                                        ("mixed", 3), ("empty", 2), ("single", 2)])
def test_sharded_steps_in_one_process(name, world):
    o, shards = _local(name, world)
    assert sum(s.counters()["compare_calls"] for s in shards) == o.compare_calls
    assert sum(s.counters()["distinct_keys"] for s in shards) == o.distinct_keys
    for s in shards:
        _check(s.edges(), o)
        r, l = s.extensions()
        np.testing.assert_array_equal(emul.unpack_ext(r)[0], o.right_ext["id"][1:])
        np.testing.assert_array_equal(emul.unpack_ext(l)[0], o.left_ext["id"][1:])
        s.close()


@pytest.mark.parametrize("name,world", [("rep", 2), ("varlen_err", 3), ("hicopy", 2)])
def test_sharded_steps_over_mailboxes(name, world, monkeypatch):
    """The peer-memory transport (route_post / answer_post / route_collect between barriers) of the same steps."""
    if name == "varlen_err":
        monkeypatch.setenv("SAGE2_EMUL_FAKE_TAG_COLLISIONS", "0x3f00000000")
    reads, k = datasets.get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    shards = []
    for r in range(world):
        s = emul.EmuShard(b, off, k)
        s.build_hash_table_shard(r, world)
        shards.append(s)
    multi.run_local([multi.mailbox_steps(s, r, world, 3000) for r, s in enumerate(shards)])
    sent = [[0] for _ in shards]
    multi.run_local([multi.sharded_graph_steps(s, r, world, multi.host_view, 3000, p2p=True, sent=sent[r]) for r, s in enumerate(shards)],
                    views_of=lambda r, bufs: emul.host_phase_a_views(bufs, world))
    assert sum(s.counters()["compare_calls"] for s in shards) == o.compare_calls
    assert all(x[0] > 0 for x in sent)
    for s in shards:
        _check(s.edges(), o)


def test_small_batches(monkeypatch):
    o, shards = _local("rep", 3, batch_reads=700)
    for s in shards:
        _check(s.edges(), o)


def test_fake_tag_collisions_take_the_verified_pass(monkeypatch):
    monkeypatch.setenv("SAGE2_EMUL_FAKE_TAG_COLLISIONS", "0x3f00000000")
    o, shards = _local("varlen_err", 2)
    assert sum(s.counters()["probe_restarts"] for s in shards) > 0
    assert sum(s.counters()["compare_calls"] for s in shards) == o.compare_calls
    for s in shards:
        _check(s.edges(), o)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out, fake):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if fake:
        os.environ["SAGE2_EMUL_FAKE_TAG_COLLISIONS"] = "0x1f00000000"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        reads, k = datasets.get(name)
        b, off = synth.concat(reads)
        s = emul.EmuShard(b, off, k)
        s.build_hash_table_shard(rank, world)

        sent = multi.run_dist(_with_views(multi.sharded_graph_steps(s, rank, world, multi.host_view, 5000), world), rank, world, "cpu")
        np.save(os.path.join(out, f"edges{rank}.npy"), s.edges())
        np.save(os.path.join(out, f"info{rank}.npy"), np.array([sent, s.counters()["probe_restarts"], s.counters()["compare_calls"]]))
    finally:
        dist.destroy_process_group()


def _with_views(gen, world):
    """Adds the host views to the generator's ("phase_a", bufs) request (run_dist would build device views)."""
    val = None
    try:
        req = next(gen)
        while True:
            if req[0] == "phase_a":
                req = ("phase_a", req[1], emul.host_phase_a_views(req[1], world))
            val = yield req
            req = gen.send(val)
    except StopIteration:
        return


@pytest.mark.parametrize("name,world,fake", [("rep", 2, False), ("varlen_err", 2, True), ("deep_varlen", 3, False), ("hicopy", 2, False)])
def test_sharded_steps_over_gloo(name, world, fake, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path), fake), nprocs=world, join=True)
    reads, k = datasets.get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    calls = restarts = 0
    for r in range(world):
        _check(np.load(tmp_path / f"edges{r}.npy"), o)
        info = np.load(tmp_path / f"info{r}.npy")
        assert info[0] > 0
        restarts += int(info[1]); calls += int(info[2])
    assert calls == o.compare_calls
    assert (restarts > 0) == fake
