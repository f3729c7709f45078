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
    reads, k = datasets.get(name) if name in datasets.DATASETS else synth.config(name)
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


def test_partitioned_cfg2_full_size_equals_reference():
    """cfg2 at full size over 4 ranks: digests of the unmodified reference's `.reads` / `.graph3` on every rank."""
    _, _, _, gpus = _run("cfg2", 4)
    for g in gpus:
        d, c = g.digest(), g.counters()
        assert d["edges"] == BIG["cfg2"]["edges_digest"] and d["reads"] == BIG["cfg2"]["reads_digest"]
        assert c["n_edges"] == BIG["cfg2"]["n_edges"] and c["unique_reads"] == BIG["cfg2"]["unique_reads"]
