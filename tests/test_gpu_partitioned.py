"""GPU parity of the fully partitioned multi-GPU build (include/sage2gpu.h "Several GPUs, every stage partitioned"):
`world` contexts on ONE GPU each organise the reads of one key range, build one key-hash shard of the table and search
one slice of the read ids; the all-gathers of sage2_b200/multi.py are done in memory (run_local).  Every context must end
with the oracle's reads (ids, frequencies, both strands), table counters, extension records and edge list."""
import json
import os

import numpy as np
import pytest
import torch

import datasets
from oracle import oracle
from sage2_b200 import api, multi, synth
from test_gpu_parity import _compare, BIG

pytestmark = pytest.mark.gpu


def _run(name, world):
    reads, k = datasets.get(name) if name in datasets.DATASETS else synth.config_cached(name)
    b, off = synth.concat(reads)
    dev = torch.device("cuda", 0)
    tb, to = torch.from_numpy(b).to(dev), torch.from_numpy(off).to(dev)
    torch.cuda.synchronize()
    view = multi.device_view_fn(dev)
    gpus = [api.Sage2Gpu(0) for _ in range(world)]
    multi.run_local([multi.partitioned_graph_steps(g, r, world, view, tb.data_ptr(), to.data_ptr(), len(off) - 1, k, True)
                     for r, g in enumerate(gpus)])
    return b, off, k, gpus


@pytest.mark.parametrize("name,world", [("clean", 2), ("rep", 3), ("err", 4), ("hicopy", 2), ("varlen_err", 5), ("deep_varlen", 2),
                                        ("tandem", 8), ("mixed", 3), ("k31", 2), ("k70", 4), ("single", 2), ("allbad", 2), ("empty", 3)])
def test_partitioned_build_equals_oracle(name, world):
    b, off, k, gpus = _run(name, world)
    o = oracle.OracleRun(b, off, k)
    total_calls = 0
    for g in gpus:
        c = g.counters()
        total_calls += c["compare_calls"]
        c_fix = dict(c)
        # compare_calls is per rank (its slice of the reads); everything else must equal the oracle on every rank
        assert c["unique_reads"] == o.U and c["distinct_keys"] == o.distinct_keys and c["keys_over_threshold"] == o.keys_over_threshold
        r = g.reads()
        np.testing.assert_array_equal(r["length"], o.length[1:])
        np.testing.assert_array_equal(r["frequency"], o.frequency[1:])
        np.testing.assert_array_equal(r["fwd"], o.fwd)
        np.testing.assert_array_equal(r["rc"], o.rc)
        e = g.edges()
        assert len(e) == o.n_edges
        for f in ("from", "to", "type", "delta", "delta_twin"):
            np.testing.assert_array_equal(e[f], o.edges[f])
        np.testing.assert_array_equal(g.extensions()["explored"], o.explored_b[1:])
    assert total_calls == o.compare_calls


@pytest.mark.parametrize("name,world", [("rep", 3), ("hicopy", 2), ("hicopy", 8), ("varlen_err", 5), ("single", 2), ("empty", 3)])
def test_partitioned_build_with_bucketed_table_shards(name, world, monkeypatch):
    """The same with every table shard built from bucketed (hash, entry) records (csrc/table.cu: owned records compacted,
    one radix pass, inserts in slot order); `hicopy` is the uneven key split that overflows the first compaction."""
    monkeypatch.setenv("SAGE2GPU_TABLE_BUILD", "bucketed")
    test_partitioned_build_equals_oracle(name, world)
    monkeypatch.setenv("SAGE2GPU_TABLE_ROOM", "8")      # the compaction runs out of room and is repeated with room for all
    test_partitioned_build_equals_oracle(name, world)


def test_partitioned_cfg2_full_size_equals_reference():
    """cfg2 at full size over 4 ranks: digests of the unmodified reference's `.reads` / `.graph3` on every rank."""
    _, _, _, gpus = _run("cfg2", 4)
    for g in gpus:
        d, c = g.digest(), g.counters()
        assert d["edges"] == BIG["cfg2"]["edges_digest"] and d["reads"] == BIG["cfg2"]["reads_digest"]
        assert c["n_edges"] == BIG["cfg2"]["n_edges"] and c["unique_reads"] == BIG["cfg2"]["unique_reads"]


@pytest.mark.parametrize("world", [1, 3])
def test_device_generated_slices_equal_oracle(world):
    """Config-#5 path at a size the oracle handles: every rank generates its slice of a synthetic read set ON THE DEVICE
    (sage2gpu_synth_reads), packs it (sage2gpu_pack_slice), the packed records are all-gathered and the rest runs partitioned;
    the characters, copied back, go through the oracle."""
    G, L, k, n_pairs = 150_000, 100, 45, 26_000
    dev = torch.device("cuda", 0)
    view = multi.device_view_fn(dev)
    gpus = [api.Sage2Gpu(0) for _ in range(world)]
    chunk = -(-n_pairs // world)
    slices, steps = [], []
    for r, g in enumerate(gpus):
        p0, p1 = min(n_pairs, r * chunk), min(n_pairs, (r + 1) * chunk)
        n = 2 * (p1 - p0)
        tb = torch.empty(max(1, n * L), dtype=torch.uint8, device=dev)
        to = torch.empty(n + 1, dtype=torch.int64, device=dev)
        g.synth_reads(tb.data_ptr(), to.data_ptr(), p0, p1 - p0, G, L, 300.0, 20.0, 7)
        slices.append((tb, to, n))
        steps.append(multi.partitioned_slice_steps(g, r, world, view, tb.data_ptr(), to.data_ptr(), n, k, True, L))
    multi.run_local(steps)
    bases = np.concatenate([tb[:n * L].cpu().numpy() for tb, _, n in slices])
    reads = bases.reshape(-1, L)
    assert set(np.unique(reads)) <= set(b"ACGT")
    # mates of a pair come from one fragment: mate 2 is the reverse complement of a stretch downstream of mate 1 (or the other way round)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    assert o.U > 0.9 * len(reads) * 0.5 and o.n_edges > 0
    for g in gpus:
        c = g.counters()
        assert c["good_reads"] == len(reads) and c["unique_reads"] == o.U
        e = g.edges()
        assert len(e) == o.n_edges
        for f in ("from", "to", "type", "delta", "delta_twin"):
            np.testing.assert_array_equal(e[f], o.edges[f])
        r = g.reads()
        np.testing.assert_array_equal(r["frequency"], o.frequency[1:])
        np.testing.assert_array_equal(r["fwd"], o.fwd)
