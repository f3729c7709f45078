"""Two real GPUs (skipped on a one-GPU box): torchrun-style ranks over NCCL build one graph with the
partitioned phase A + exchange of sage2_b200/multi.py; every rank must hold the oracle's edge list."""
import os
import socket

import numpy as np
import pytest
import torch

import datasets
from oracle import oracle
from sage2_b200 import synth

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out, sharded=False, p2p=False, partitioned=False):
    import torch.distributed as dist
    from sage2_b200 import api, multi
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        reads, k = datasets.get(name)
        b, off = synth.concat(reads)
        hb, ho = torch.from_numpy(b).pin_memory(), torch.from_numpy(off).pin_memory()
        dev = torch.device("cuda", rank)
        tb, to, _ = multi.upload_partitioned(hb, ho, rank, world, dev)
        torch.cuda.synchronize()
        g = api.Sage2Gpu(rank)
        if partitioned:  # every stage partitioned, reads / table shards / phase-A arrays all-gathered over NCCL
            multi.build_partitioned(g, rank, world, dev, tb.data_ptr(), to.data_ptr(), len(off) - 1, k, True)
            np.save(os.path.join(out, f"edges{rank}.npy"), g.edges())
            np.save(os.path.join(out, f"freq{rank}.npy"), g.reads()["frequency"])
            return
        g.load_reads_ptr(tb.data_ptr(), to.data_ptr(), len(off) - 1, k, device=True)
        if sharded:      # key-hash shard per GPU, probes routed by NCCL all-to-all
            g.build_hash_table_shard(rank, world)
            multi.build_overlap_graph_sharded(g, rank, world, dev, batch_reads=20000, p2p=p2p)
        else:
            g.build_hash_table()
            multi.build_overlap_graph(g, rank, world, dev)
        np.save(os.path.join(out, f"edges{rank}.npy"), g.edges())
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("name", ["rep", "varlen_err"])
def test_two_gpus_one_graph(name, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), name, str(tmp_path)), nprocs=2, join=True)
    reads, k = datasets.get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    for r in range(2):
        e = np.load(tmp_path / f"edges{r}.npy")
        assert len(e) == o.n_edges
        for f in ("from", "to", "type", "delta", "delta_twin"):
            np.testing.assert_array_equal(e[f], o.edges[f])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("name", ["rep", "varlen_err", "hicopy"])
def test_two_gpus_sharded_table(name, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), name, str(tmp_path), True), nprocs=2, join=True)
    reads, k = datasets.get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    for r in range(2):
        e = np.load(tmp_path / f"edges{r}.npy")
        assert len(e) == o.n_edges
        for f in ("from", "to", "type", "delta", "delta_twin"):
            np.testing.assert_array_equal(e[f], o.edges[f])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("name", ["rep", "varlen_err", "hicopy"])
def test_two_gpus_sharded_table_over_peer_memory(name, tmp_path):
    """Two processes, CUDA IPC mailboxes: the routing kernel stores into the other GPU's memory over NVLink."""
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), name, str(tmp_path), True, True), nprocs=2, join=True)
    reads, k = datasets.get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    for r in range(2):
        e = np.load(tmp_path / f"edges{r}.npy")
        assert len(e) == o.n_edges
        for f in ("from", "to", "type", "delta", "delta_twin"):
            np.testing.assert_array_equal(e[f], o.edges[f])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("name", ["rep", "varlen_err", "hicopy", "mixed"])
def test_two_gpus_every_stage_partitioned(name, tmp_path):
    """Two processes over NCCL: reads organised by key range, table built by key-hash shard, phase A by id slice; the
    all-gathers of multi.partitioned_graph_steps make every rank hold the oracle's reads and edge list."""
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), name, str(tmp_path), False, False, True), nprocs=2, join=True)
    reads, k = datasets.get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    for r in range(2):
        e = np.load(tmp_path / f"edges{r}.npy")
        assert len(e) == o.n_edges
        for f in ("from", "to", "type", "delta", "delta_twin"):
            np.testing.assert_array_equal(e[f], o.edges[f])
        np.testing.assert_array_equal(np.load(tmp_path / f"freq{r}.npy"), o.frequency[1:])
