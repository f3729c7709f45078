"""Seeded random sweep (CPU): read sets with random genome size, read length, coverage, k, error rate, repeats and variable
lengths through the CPU emulation of the product's primitives (tests/host_emul.cpp over core.cuh + host_phase_c.cpp) -- the
single-table pipeline and the sharded-table steps of sage2_b200/multi.py -- against the oracle.  The fixed data sets of
tests/datasets.py pin the known corner cases; this sweep looks for the ones nobody thought of."""
import numpy as np
import pytest

import emul
from oracle import oracle
from sage2_b200 import multi, synth
from test_host_emulation import compare_stage_outputs


def _random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    G = int(rng.integers(4_000, 30_000))
    L = int(rng.integers(50, 140))
    k = int(rng.integers(20, min(L - 5, 95)))
    cov = float(rng.integers(8, 70))
    g = synth.random_genome(G, 5000 + seed)
    kind = int(rng.integers(0, 4))
    if kind == 1:
        g = synth.add_repeats(g, int(rng.integers(3, 150)), int(rng.integers(k, 3 * L)), 6000 + seed)
    elif kind == 2:
        g = synth.add_tandem(g, int(rng.integers(3, 40)), int(rng.integers(10, 80)), int(rng.integers(0, G // 2)), 7000 + seed)
    err = [0.0, 0.0, 0.004, 0.012][int(rng.integers(0, 4))]
    reads = synth.paired_reads(g, L, cov, seed=8000 + seed, mu=3 * L, sigma=L // 5, err_rate=err)
    if rng.random() < 0.4:
        reads = synth.variable_length(reads, max(k - 5, L // 2), seed=9000 + seed)
    return reads, k


@pytest.mark.parametrize("seed", range(10))
def test_random_read_sets_equal_oracle(seed):
    reads, k = _random_case(seed)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    e = emul.EmuRun(b, off, k)
    assert (e.N, e.distinct, e.over, e.compare_calls) == (o.N, o.distinct_keys, o.keys_over_threshold, o.compare_calls)
    assert (e.inserted, e.removed) == (o.edges_inserted_c, o.transitive_removed)
    compare_stage_outputs(o, e.U, e.len, e.freq, e.F, e.RC, e.extR, e.extL, e.explored_b, e.edges, e.explored_a)


@pytest.mark.parametrize("seed,world,p2p", [(20, 2, False), (21, 3, True), (22, 5, False), (23, 2, True)])
def test_random_read_sets_sharded(seed, world, p2p, monkeypatch):
    if seed % 2:
        monkeypatch.setenv("SAGE2_EMUL_FAKE_TAG_COLLISIONS", "0x7f00000000")
    reads, k = _random_case(seed)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    shards = []
    for r in range(world):
        s = emul.EmuShard(b, off, k)
        s.build_hash_table_shard(r, world)
        shards.append(s)
    batch = 1500
    views = lambda r, bufs: emul.host_phase_a_views(bufs, world)
    if p2p:
        multi.run_local([multi.mailbox_steps(s, r, world, batch) for r, s in enumerate(shards)])
    multi.run_local([multi.sharded_graph_steps(s, r, world, multi.host_view, batch, p2p=p2p) for r, s in enumerate(shards)], views_of=views)
    assert sum(s.counters()["compare_calls"] for s in shards) == o.compare_calls
    for s in shards:
        a, bb, t, d = emul.unpack_edges(s.edges())
        assert len(a) == o.n_edges
        np.testing.assert_array_equal(a, o.edges["from"])
        np.testing.assert_array_equal(bb, o.edges["to"])
        np.testing.assert_array_equal(t, o.edges["type"])
        np.testing.assert_array_equal(d, o.edges["delta"])
        s.close()
