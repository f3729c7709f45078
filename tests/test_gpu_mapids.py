"""GPU parity of the step-6 mapping (sage2gpu_map_reads = ReadLoader::getIdOfRead, readLoader.cpp:319-353):
against the ids the UNMODIFIED reference printed for the same queries (tests/golden/mapids.json), against the oracle on
more data sets, and at cfg2's full size through properties (every input read maps to a read of its own sequence; the
number of reads mapped to an id is that read's frequency)."""
import hashlib
import json
import os

import numpy as np
import pytest

import datasets
from oracle import oracle
from sage2_b200 import api, synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
MAPIDS = json.load(open(os.path.join(HERE, "golden", "mapids.json")))


def _loader(reads, k):
    loader = api.ReadLoader(k)
    loader.readDatasetInBytes(reads)
    loader.organizeReads()
    return loader


@pytest.mark.parametrize("name", sorted(MAPIDS))
def test_map_reads_equals_reference(name):
    g = MAPIDS[name]
    reads, k = datasets.get(name)
    queries, _ = datasets.map_queries(name)
    loader = _loader(reads, k)
    qb, qoff = synth.concat(queries)
    ids, good, _ = loader.gpu.map_reads(qb, qoff)
    text = ["bad" if not gd else str(int(i)) for i, gd in zip(ids, good)]
    assert text[:64] == g["first"]
    assert hashlib.md5("\n".join(text).encode()).hexdigest() == g["md5"]
    np.testing.assert_array_equal(loader.getIdOfRead(queries), ids)


@pytest.mark.parametrize("name", ["k31", "k64", "hicopy", "deep_varlen", "adapter", "single", "empty", "allbad", "err"])
def test_map_reads_equals_oracle(name):
    reads, k = datasets.get(name)
    queries, _ = datasets.map_queries(name, n=3000, seed=31)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    qb, qoff = synth.concat(queries)
    want_ids, want_good = o.map_reads(qb, qoff, k)
    g = api.Sage2Gpu(0)
    g.load_reads(b, off, k)
    ids, good, _ = g.map_reads(qb, qoff)
    np.testing.assert_array_equal(good, want_good)
    np.testing.assert_array_equal(ids, want_ids)


def test_map_reads_cfg2_full_size_properties():
    reads, k = synth.config("cfg2")
    b, off = synth.concat(reads)
    g = api.Sage2Gpu(0)
    g.load_reads(b, off, k)
    ids, good, ms = g.map_reads(b, off)
    assert good.all() and (ids != 0).all()
    r = g.reads()
    U = len(r["length"])
    np.testing.assert_array_equal(np.bincount(np.abs(ids), minlength=U + 1)[1:], r["frequency"])      # a checksum of the whole map
    # sampled reads: the stored forward strand of |id| is the read (id > 0) or its reverse complement (id < 0)
    rng = np.random.default_rng(5)
    comp = np.zeros(256, np.uint8)
    comp[[65, 67, 71, 84]] = [84, 71, 67, 65]
    for q in rng.integers(0, len(ids), 2000):
        i = abs(int(ids[q]))
        L = int(r["length"][i - 1])
        packed = r["fwd"][int(r["byte_off"][i - 1]):int(r["byte_off"][i])]
        codes = np.stack([(packed >> s) & 3 for s in (6, 4, 2, 0)], axis=1).reshape(-1)[:L]
        stored = np.frombuffer(b"ACGT", np.uint8)[codes]
        read = np.asarray(reads[q])
        want = read if ids[q] > 0 else comp[read[::-1]]
        np.testing.assert_array_equal(stored, want)
    print(f"map_reads cfg2: {len(ids)} reads, kernel {ms:.3f} ms = {len(ids) / ms / 1e3:.1f} M reads/s")
