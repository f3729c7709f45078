"""tests/digest.py (the parity gate of bench.py) on the CPU: the digests computed from the reference's own text files
equal the digests computed from the arrays the oracle returns for the same input, and the committed goldens of the
benchmark-size workloads (tests/golden/golden_big.json, made by the unmodified reference) are complete."""
import gzip
import json
import os
import shutil

import numpy as np

import datasets
import digest
from oracle import oracle
from sage2_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


def _gunzip(name, tmp_path):
    out = tmp_path / name
    with gzip.open(os.path.join(HERE, "golden", name + ".gz"), "rb") as fi, open(out, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    return str(out)


def test_file_digests_equal_array_digests(tmp_path):
    """`mixed.reads` / `mixed.graph3` are complete files the unmodified reference wrote (tests/golden/make_golden.py)."""
    reads, k = datasets.get("mixed")
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    d_edges, E = digest.graph3_file_digest(_gunzip("mixed.graph3", tmp_path), chunk=1000)
    assert E == o.n_edges and d_edges == digest.edges_digest_total(o.edges)
    d_reads = digest.reads_file_digest(_gunzip("mixed.reads", tmp_path), chunk=777)
    rows, U = [], o.U
    for i in range(1, U + 1):
        n = int(o.length[i])
        bits = np.unpackbits(o.fwd[int(o.byte_off[i]):int(o.byte_off[i + 1])])
        codes = bits[0:2 * n:2] * 2 + bits[1:2 * n:2]
        rows.append(bytes(np.frombuffer(b"ACGT", np.uint8)[codes]))
    want = digest.finish(digest.reads_digest_ragged(np.arange(1, U + 1), o.frequency[1:], o.length[1:], rows), U)
    assert d_reads == want


def test_digest_is_order_and_content_sensitive():
    reads, k = datasets.get("rep")
    b, off = synth.concat(reads)
    e = oracle.OracleRun(b, off, k).edges
    d0 = digest.edges_digest_total(e)
    sw = e.copy()
    sw[[0, 1]] = sw[[1, 0]]
    assert digest.edges_digest_total(sw) != d0
    ch = e.copy()
    ch["delta"][len(ch) // 2] += 1
    assert digest.edges_digest_total(ch) != d0
    assert digest.edges_digest_total(e[:-1]) != d0


def test_big_goldens_are_complete():
    big = json.load(open(os.path.join(HERE, "golden", "golden_big.json")))
    for name in ("cfg2", "cfg3-40", "cfg3-60", "cfg3-90"):
        g = big[name]
        for key in ("reads_md5", "graph3_md5", "edges_digest", "reads_digest", "n_edges", "unique_reads", "good_reads"):
            assert g.get(key) is not None, (name, key)
    # cfg3 shares cfg2's reads: identical `.reads`, and (error-free data) the same chain of edges whatever k
    assert big["cfg3-40"]["reads_md5"] == big["cfg2"]["reads_md5"] == big["cfg3-90"]["reads_md5"]
