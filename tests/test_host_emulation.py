"""CPU checks of the product's bit arithmetic and host logic (no GPU needed).

tests/host_emul.cpp drives sage2_b200/csrc/core.cuh (the host/device primitives the kernels are built
from) and host_phase_c.cpp sequentially; its every stage must equal the oracle's.
"""
import os

import numpy as np
import pytest

import datasets
import emul
from oracle import oracle
from sage2_b200 import synth

NAMES = ["clean", "k31", "k70", "k64", "err", "rep", "hicopy", "deep", "varlen", "varlen_err", "deep_varlen",
         "tandem", "mixed", "adapter", "empty", "allbad", "single"]


def compare_stage_outputs(o, U, lens, freq, F, RC, extR, extL, explored_b, edges, explored_a=None):
    """Shared by the emulation tests and the GPU parity tests: every stage equals the oracle."""
    assert U == o.U
    np.testing.assert_array_equal(lens, o.length[1:])
    np.testing.assert_array_equal(freq, o.frequency[1:])
    np.testing.assert_array_equal(emul.records_to_bytes(F, lens), o.fwd)
    np.testing.assert_array_equal(emul.records_to_bytes(RC, lens), o.rc)
    for mine, ref in ((extR, o.right_ext), (extL, o.left_ext)):
        i, t, l = emul.unpack_ext(mine)
        np.testing.assert_array_equal(i, ref["id"][1:])
        has = i != 0          # the reference leaves type/length of absent extensions uninitialised
        np.testing.assert_array_equal(t[has], ref["type"][1:][has])
        np.testing.assert_array_equal(l[has], ref["length"][1:][has])
    if explored_a is not None:
        np.testing.assert_array_equal(explored_a, o.explored_a[1:])
    np.testing.assert_array_equal(explored_b, o.explored_b[1:])
    a, b, t, d = emul.unpack_edges(edges)
    assert len(a) == o.n_edges
    np.testing.assert_array_equal(a, o.edges["from"])
    np.testing.assert_array_equal(b, o.edges["to"])
    np.testing.assert_array_equal(t, o.edges["type"])
    np.testing.assert_array_equal(d, o.edges["delta"])


@pytest.mark.parametrize("name", NAMES)
def test_emulated_pipeline_equals_oracle(name):
    reads, k = datasets.get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    e = emul.EmuRun(b, off, k)
    assert e.N == o.N
    assert e.over == o.keys_over_threshold
    assert e.distinct == o.distinct_keys
    assert e.compare_calls == o.compare_calls
    assert (e.contained, e.contained_size) == (o.contained_ext, o.contained_size)
    assert (e.inserted, e.removed) == (o.edges_inserted_c, o.transitive_removed)
    compare_stage_outputs(o, e.U, e.len, e.freq, e.F, e.RC, e.extR, e.extL, e.explored_b, e.edges, e.explored_a)


def test_window_extraction_all_offsets():
    """get_bases == get64BitInt (utils.cpp:189-207) at every start phase and length 1..32."""
    rng = np.random.default_rng(7)
    L = emul.lib()
    s = bytes(rng.choice(list(b"ACGT"), size=200).astype(np.uint8))
    for SW in (7, 8):
        rec = np.zeros(SW, np.uint64)
        L.hemu_pack(s, 200, SW, rec.ctypes.data, 0)
        packed = np.zeros(52, np.uint8)
        oracle.lib().sgo_chars_to_bytes(s, 200, packed.ctypes.data)
        for start in list(range(0, 40)) + list(range(150, 168)):
            for n in range(1, 33):
                assert L.hemu_get_bases(rec.ctypes.data, SW, start, n) == oracle.lib().sgo_get64(packed.ctypes.data, start, n)


def test_revcomp_record_all_lengths():
    rng = np.random.default_rng(8)
    L = emul.lib()
    for length in list(range(1, 70)) + [95, 96, 97, 127, 128, 129, 150, 250]:
        s = bytes(rng.choice(list(b"ACGT"), size=length).astype(np.uint8))
        for SW in {max(2, (2 * length + 16 + 63) // 64), 8}:
            f = np.zeros(SW, np.uint64); r = np.zeros(SW, np.uint64); want = np.zeros(SW, np.uint64)
            L.hemu_pack(s, length, SW, f.ctypes.data, 0)
            L.hemu_pack(s, length, SW, want.ctypes.data, 1)
            L.hemu_revcomp(f.ctypes.data, r.ctypes.data, SW, length)
            np.testing.assert_array_equal(r, want)


def test_record_order_is_read_order():
    """word-wise record compare == Read::operator< (readLoader.cpp:11-18) incl. length tie-break."""
    rng = np.random.default_rng(9)
    L = emul.lib()
    SW = 4
    strs = [b"A" * 40, b"A" * 41, b"A" * 44, b"A" * 40 + b"C", b"ACGT" * 10, b"ACGT" * 10 + b"A", b"ACGT" * 10 + b"AAAA",
            b"ACGT" * 10 + b"AAAAA", b"T" * 50, b"T" * 49]
    strs += [bytes(rng.choice(list(b"ACGT"), size=int(n)).astype(np.uint8)) for n in rng.integers(30, 100, size=60)]
    recs, packed = [], []
    for s in strs:
        r = np.zeros(SW, np.uint64)
        L.hemu_pack(s, len(s), SW, r.ctypes.data, 0)
        p = np.zeros(26, np.uint8)
        oracle.lib().sgo_chars_to_bytes(s, len(s), p.ctypes.data)
        recs.append(tuple(int(x) for x in r)); packed.append(p)
    for i in range(len(strs)):
        for j in range(len(strs)):
            c = oracle.lib().sgo_string_compare(packed[i].ctypes.data, len(strs[i]), packed[j].ctypes.data, len(strs[j]))
            mine = (recs[i] > recs[j]) - (recs[i] < recs[j])
            assert mine == c, (strs[i], strs[j])


def test_phase_c_traversal_from_device_lists_equals_the_walk_on_random_graphs(tmp_path):
    """tests/phase_c_order_fuzz.cpp: the traversal graph.cu runs for asymmetric candidate sets (run_host_phase_c_order_lists, over
    the lists phase_c_sorted_lists prepares) explores in the same order as the walk itself (economyGraph.cpp:513-638) on random
    candidate graphs: asymmetric sets, several candidates per pair, self candidates, negative overhangs, phase-B records."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    exe = str(tmp_path / "phase_c_order_fuzz")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(here, "phase_c_order_fuzz.cpp"),
                           os.path.join(root, "sage2_b200", "csrc", "host_phase_c.cpp")])
    for seed in (1, 2, 3):
        out = subprocess.run([exe, "300", str(seed)], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        assert "same order" in out.stdout
