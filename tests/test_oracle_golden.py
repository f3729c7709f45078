"""Pin the CPU oracle (oracle/sage2_oracle.c) against the UNMODIFIED reference.

tests/golden/golden.json holds md5s of the reference's own `.reads` / `.graph3` (from
`SAGE2 -s -M 3`, OMP_NUM_THREADS=1) and its log counters for seeded inputs; see
tests/golden/make_golden.py.  The oracle must reproduce every byte and every counter.
"""
import gzip
import hashlib
import json
import os

import pytest

import datasets
from oracle import oracle
from sage2_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def _md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def _run(name):
    reads, k = datasets.get(name) if name in datasets.DATASETS else synth.config(name)
    b, o = synth.concat(reads)
    return oracle.OracleRun(b, o, k)


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_matches_reference_bytes(name, tmp_path):
    g = GOLD[name]
    r = _run(name)
    r.write_reads(str(tmp_path / "o.reads"))
    r.write_graph3(str(tmp_path / "o.graph3"))
    assert _md5(tmp_path / "o.reads") == g["reads_md5"]
    assert _md5(tmp_path / "o.graph3") == g["graph3_md5"]
    assert r.U == g["unique_reads"]
    assert r.N == g["good_reads"]
    assert r.keys_over_threshold == g["over_threshold"]
    assert r.contained_ext == g["contained_ext"]
    assert r.contained_size == g["contained_size"]
    assert r.left_to_explore == g["left_to_explore"]
    assert r.edges_inserted_c == g["edges_inserted"]
    assert r.transitive_removed == g["transitive_removed"]


def test_oracle_matches_committed_fixture(tmp_path):
    r = _run("mixed")
    r.write_reads(str(tmp_path / "o.reads"))
    r.write_graph3(str(tmp_path / "o.graph3"))
    for ext in ("reads", "graph3"):
        want = gzip.open(os.path.join(HERE, "golden", "mixed." + ext + ".gz"), "rb").read()
        assert open(tmp_path / ("o." + ext), "rb").read() == want


@pytest.mark.parametrize("name", ["empty", "allbad", "single"])
def test_oracle_degenerate_inputs(name):
    r = _run(name)
    assert r.n_edges == 0
    assert r.U == (1 if name == "single" else 0)


MAPIDS = json.load(open(os.path.join(HERE, "golden", "mapids.json")))


@pytest.mark.parametrize("name", sorted(MAPIDS))
def test_map_reads_equals_reference_getIdOfRead(name):
    """oracle sgo_map_reads (restating readLoader.cpp:319-353 + the isGoodRead gate of matePair.cpp:176-179) against
    the ids the UNMODIFIED reference's ReadLoader::getIdOfRead printed for the same queries (make_mapids_golden.py)."""
    import hashlib
    g = MAPIDS[name]
    reads, k = datasets.get(name)
    queries, _ = datasets.map_queries(name)
    assert k == g["k"] and len(queries) == g["n_queries"]
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    qb, qoff = synth.concat(queries)
    ids, good = o.map_reads(qb, qoff, k)
    text = ["bad" if not gd else str(int(i)) for i, gd in zip(ids, good)]
    assert text[:64] == g["first"]
    assert hashlib.md5("\n".join(text).encode()).hexdigest() == g["md5"]
    assert (int((ids > 0).sum()), int((ids < 0).sum()), int(((ids == 0) & (good == 1)).sum()), int((good == 0).sum())) == \
        (g["positive"], g["negative"], g["absent"], g["bad"])
