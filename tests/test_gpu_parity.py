"""GPU parity tests (run on the B200 box with -m gpu).  Every stage of the CUDA path, called through
the C ABI, must equal the oracle on the same seeded inputs; the text outputs must be byte-identical
to the UNMODIFIED reference's (md5s in tests/golden/golden.json)."""
import hashlib
import json
import os

import numpy as np
import pytest

import datasets
import emul
from oracle import oracle
from sage2_b200 import api, synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))
SMALL = ["clean", "k31", "k70", "k64", "err", "rep", "hicopy", "deep", "varlen", "varlen_err", "deep_varlen",
         "tandem", "mixed", "adapter", "empty", "allbad", "single"]


def _md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def _get(name):
    if name in datasets.DATASETS:
        return datasets.get(name)
    if name == "cfg2" or name.startswith("cfg3-"):       # cfg3 = the cfg2 reads with another k
        reads, _ = synth.config_cached("cfg2")
        return reads, (63 if name == "cfg2" else int(name.split("-")[1]))
    return synth.config(name)


def _run_gpu(reads, k):
    loader = api.ReadLoader(k)
    loader.readDatasetInBytes(reads)
    loader.organizeReads()
    table = api.HashTable(k, loader)
    table.hashPrefixesAndSuffix()
    graph = api.EconomyGraph(k, table)
    graph.buildInitialOverlapGraph()
    graph.buildOverlapGraphEconomy()
    graph.sortEconomyGraph()
    return loader, graph


def _compare(o, gpu):
    c = gpu.counters()
    assert c["good_reads"] == o.N and c["unique_reads"] == o.U and c["avg_len"] == o.avg_len
    assert c["distinct_keys"] == o.distinct_keys and c["keys_over_threshold"] == o.keys_over_threshold
    assert c["compare_calls"] == o.compare_calls
    assert (c["contained_ext"], c["contained_size"], c["left_to_explore"]) == (o.contained_ext, o.contained_size, o.left_to_explore)
    assert (c["edges_inserted_c"], c["transitive_removed"]) == (o.edges_inserted_c, o.transitive_removed)
    r = gpu.reads()
    np.testing.assert_array_equal(r["length"], o.length[1:])
    np.testing.assert_array_equal(r["frequency"], o.frequency[1:])
    np.testing.assert_array_equal(r["fwd"], o.fwd)
    np.testing.assert_array_equal(r["rc"], o.rc)
    x = gpu.extensions()
    for mine, ref in ((x["right"], o.right_ext), (x["left"], o.left_ext)):
        i, t, l = emul.unpack_ext(mine)
        np.testing.assert_array_equal(i, ref["id"][1:])
        has = i != 0
        np.testing.assert_array_equal(t[has], ref["type"][1:][has])
        np.testing.assert_array_equal(l[has], ref["length"][1:][has])
    np.testing.assert_array_equal(x["explored"], o.explored_b[1:])
    e = gpu.edges()
    assert len(e) == o.n_edges
    for f in ("from", "to", "type", "delta", "delta_twin"):
        np.testing.assert_array_equal(e[f], o.edges[f])


@pytest.mark.parametrize("name", SMALL)
def test_every_stage_equals_oracle(name, tmp_path):
    reads, k = _get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    loader, graph = _run_gpu(reads, k)
    _compare(o, loader.gpu)
    loader.saveReadsInFile(str(tmp_path / "g.reads"))
    graph.saveOverlapGraphInFile(str(tmp_path / "g.graph3"))
    if name in GOLD:       # bytes the unmodified reference wrote for the same input
        assert _md5(tmp_path / "g.reads") == GOLD[name]["reads_md5"]
        assert _md5(tmp_path / "g.graph3") == GOLD[name]["graph3_md5"]


@pytest.mark.parametrize("name", ["cfg1", "cfg4mini"])
def test_baseline_configs_byte_identical_to_reference(name, tmp_path):
    reads, k = _get(name)
    loader, graph = _run_gpu(reads, k)
    loader.saveReadsInFile(str(tmp_path / "g.reads"))
    graph.saveOverlapGraphInFile(str(tmp_path / "g.graph3"))
    g = GOLD[name]
    c = loader.gpu.counters()
    assert c["unique_reads"] == g["unique_reads"] and c["good_reads"] == g["good_reads"]
    assert c["contained_ext"] == g["contained_ext"] and c["left_to_explore"] == g["left_to_explore"]
    assert c["edges_inserted_c"] == g["edges_inserted"] and c["transitive_removed"] == g["transitive_removed"]
    assert _md5(tmp_path / "g.reads") == g["reads_md5"]
    assert _md5(tmp_path / "g.graph3") == g["graph3_md5"]


def test_rerun_same_context_is_idempotent():
    reads, k = _get("rep")
    b, off = synth.concat(reads)
    gpu = api.Sage2Gpu(0)
    gpu.run_steps123(b, off, k)
    e1 = gpu.edges().copy()
    gpu.run_steps123(b, off, k)
    e2 = gpu.edges()
    np.testing.assert_array_equal(e1, e2)


def test_input_order_does_not_matter():
    """Read ids are ranks in the sorted unique set: shuffling the input must not change any output."""
    reads, k = _get("err")
    b, off = synth.concat(reads)
    gpu = api.Sage2Gpu(0)
    gpu.run_steps123(b, off, k)
    e1 = gpu.edges().copy()
    perm = np.random.default_rng(1).permutation(len(reads))
    b2, off2 = synth.concat(reads[perm])
    gpu.run_steps123(b2, off2, k)
    np.testing.assert_array_equal(e1, gpu.edges())


BIG = json.load(open(os.path.join(HERE, "golden", "golden_big.json")))


def _md5_file(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def _check_against_big_golden(gpu, name, tmp_path=None):
    """Counters + digests (+ the bytes of the -s files) against what the UNMODIFIED reference produced for the same
    full-size input (`SAGE2 -s -M 3`, tests/golden/make_golden_big.py)."""
    g, c = BIG[name], gpu.counters()
    assert c["good_reads"] == g["good_reads"] and c["unique_reads"] == g["unique_reads"]
    assert c["keys_over_threshold"] == g["over_threshold"]
    assert (c["contained_ext"], c["contained_size"], c["left_to_explore"]) == (g["contained_ext"], g["contained_size"], g["left_to_explore"])
    assert (c["edges_inserted_c"], c["transitive_removed"]) == (g["edges_inserted"], g["transitive_removed"])
    assert c["n_edges"] == g["n_edges"]
    d = gpu.digest()
    assert d["edges"] == g["edges_digest"] and d["reads"] == g["reads_digest"]
    if tmp_path is not None:
        gpu.write_graph3(str(tmp_path / "g.graph3"))
        assert _md5_file(tmp_path / "g.graph3") == g["graph3_md5"]
        os.remove(tmp_path / "g.graph3")
        gpu.write_reads(str(tmp_path / "g.reads"))
        assert _md5_file(tmp_path / "g.reads") == g["reads_md5"]
        os.remove(tmp_path / "g.reads")


def test_cfg2_full_size_byte_identical_to_reference(tmp_path):
    """BASELINE config #2 at full size: `.reads` / `.graph3` md5, counters and digests of the unmodified reference."""
    reads, k = synth.config_cached("cfg2")
    b, off = synth.concat(reads)
    gpu = api.Sage2Gpu(0)
    gpu.run_steps123(b, off, k)
    r = gpu.reads()
    assert int(r["frequency"].astype(np.int64).sum()) == len(reads)         # dedupe conserves reads
    _check_against_big_golden(gpu, "cfg2", tmp_path)
    # the same through the min-hash schedule of phase A (results must not depend on the order reads are searched in)
    for order in (0, 1):
        gpu.set_option("read_order", order)
        gpu.run_steps123(b, off, k)
        _check_against_big_golden(gpu, "cfg2")


@pytest.mark.skipif("cfg4" not in BIG, reason="no cfg4 golden committed")
def test_cfg4_full_size_byte_identical_to_reference(tmp_path):
    """BASELINE config #4 (100 Mbp + 2 % repeats, 33.3 M reads, -k 75) at full size against the unmodified reference."""
    reads, k = synth.config_cached("cfg4")
    b, off = synth.concat(reads)
    del reads
    gpu = api.Sage2Gpu(0)
    gpu.run_steps123(b, off, k)
    _check_against_big_golden(gpu, "cfg4", tmp_path if os.environ.get("SAGE2_TEST_CFG4_FILES") else None)


def test_digest_equals_numpy_definition():
    """sage2gpu_digest against tests/digest.py on the arrays the C ABI returns (small sets, ragged lengths)."""
    import digest
    for name in ("varlen_err", "mixed", "rep", "empty"):
        reads, k = _get(name)
        b, off = synth.concat(reads)
        gpu = api.Sage2Gpu(0)
        gpu.run_steps123(b, off, k)
        d, e, r = gpu.digest(), gpu.edges(), gpu.reads()
        assert d["edges"] == digest.edges_digest_total(e)
        U = len(r["length"])
        rows = []
        for i in range(U):
            codes = _decode(r["fwd"][int(r["byte_off"][i]):int(r["byte_off"][i + 1])], int(r["length"][i]))
            rows.append(bytes(np.frombuffer(b"ACGT", np.uint8)[codes]))
        want = digest.finish(digest.reads_digest_ragged(np.arange(1, U + 1), r["frequency"], r["length"], rows), U)
        assert d["reads"] == want


@pytest.mark.parametrize("name", SMALL)
def test_general_kernel_alone_equals_oracle(name):
    """Phase A without the superstring scan (set_option("fast_scan", 0)): the hit-by-hit kernel on every read."""
    reads, k = _get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    gpu = api.Sage2Gpu(0)
    gpu.set_option("fast_scan", 0)
    gpu.run_steps123(b, off, k)
    assert gpu.counters()["fast_path_reads"] == 0
    _compare(o, gpu)


def test_superstring_scan_takes_the_clean_reads():
    """The fast path certifies every read of an error-free random genome and leaves contradicting hits to the general kernel."""
    for name, lo, hi in (("clean", 1.0, 1.0), ("rep", 0.9, 1.0), ("err", 0.05, 0.6), ("deep", 0.0, 0.2)):
        reads, k = _get(name)
        b, off = synth.concat(reads)
        gpu = api.Sage2Gpu(0)
        gpu.run_steps123(b, off, k)
        c = gpu.counters()
        assert lo <= c["fast_path_reads"] / c["unique_reads"] <= hi, (name, c["fast_path_reads"], c["unique_reads"])


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("how", ["direct", "bucketed"])
def test_both_table_builds_equal_oracle(name, how, monkeypatch):
    """The table is built entry by entry in read order (small indexes) or from (hash, entry) records bucketed by slot range
    (indexes far larger than L2, csrc/table.cu); SAGE2GPU_TABLE_BUILD forces either on every data set."""
    monkeypatch.setenv("SAGE2GPU_TABLE_BUILD", how)
    reads, k = _get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    gpu = api.Sage2Gpu(0)
    gpu.run_steps123(b, off, k)
    _compare(o, gpu)


@pytest.mark.parametrize("name", SMALL)
def test_minhash_read_order_changes_nothing(name):
    """SAGE2GPU_READ_ORDER=minhash / set_option("read_order", 1): phase A in min-hash order must give the oracle's graph."""
    reads, k = _get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    gpu = api.Sage2Gpu(0)
    gpu.set_option("read_order", 1)
    gpu.run_steps123(b, off, k)
    _compare(o, gpu)


@pytest.mark.parametrize("name,world", [("rep", 2), ("deep_varlen", 3), ("varlen_err", 4)])
def test_partitioned_phase_a_slices_compose(name, world):
    """The multi-GPU form on one GPU: `world` contexts hold the same reads and table, each searches its
    slice (sage2gpu_phase_a_partition); the exchange of sage2_b200/multi.py is done here with plain device
    copies (all-gather of the slices, element-wise max of the containment ids); every context must then
    finish with the oracle's graph."""
    import torch
    from sage2_b200 import multi
    reads, k = _get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    dev = torch.device("cuda", 0)
    ctxs, views = [], []
    for r in range(world):
        g = api.Sage2Gpu(0)
        g.load_reads(b, off, k)
        g.build_hash_table()
        g.phase_a_partition(r, world)
        ctxs.append(g)
        views.append(multi.device_views(g.phase_a_buffers(), world, dev))
    chunk = ctxs[0].phase_a_buffers()["chunk"]
    assert chunk == -(-o.U // world)
    torch.cuda.synchronize()
    full = {key: torch.cat([views[r][key][r * chunk:(r + 1) * chunk] for r in range(world)]) for key in ("right", "left", "over_limit")}
    cmax = torch.stack([v["contained_by"] for v in views]).max(dim=0).values
    for v in views:
        for key in full:
            v[key].copy_(full[key])
        v["contained_by"].copy_(cmax)
    torch.cuda.synchronize()
    total_calls = 0
    for g in ctxs:
        total_calls += g.counters()["compare_calls"]
        g.finish_graph()
        e = g.edges()
        assert len(e) == o.n_edges
        for f in ("from", "to", "type", "delta", "delta_twin"):
            np.testing.assert_array_equal(e[f], o.edges[f])
        x = g.extensions()
        np.testing.assert_array_equal(x["explored"], o.explored_b[1:])
    assert total_calls == o.compare_calls


def _decode(packed: np.ndarray, length: int) -> np.ndarray:
    """reference byte layout (utils.cpp:96-119) -> base codes"""
    bits = np.unpackbits(packed)
    return (bits[0:2 * length:2] * 2 + bits[1:2 * length:2]).astype(np.uint8)


def _check_edges_are_exact_overlaps(gpu, k, sample=3000, seed=5):
    """Independent of the oracle: every sampled edge is an exact overlap of at least k bases of the kind its
    type says (economyGraph.cpp:607-626), and its twin overhang is consistent (overlapGraph.cpp:147)."""
    r, e = gpu.reads(), gpu.edges()
    if len(e) == 0:
        return 0
    rng = np.random.default_rng(seed)
    pick = rng.choice(len(e), size=min(sample, len(e)), replace=False)
    off, ln = r["byte_off"].astype(np.int64), r["length"].astype(np.int64)
    for x in e[pick]:
        u, v, t, d = int(x["from"]) - 1, int(x["to"]) - 1, int(x["type"]), int(x["delta"])
        lu, lv = int(ln[u]), int(ln[v])
        ov = lv - d
        assert k <= ov <= min(lu, lv), (x, ov)
        uf = _decode(r["fwd"][off[u]:off[u + 1]], lu)
        vs = _decode((r["rc"] if t in (1, 2) else r["fwd"])[off[v]:off[v + 1]], lv)
        if t in (2, 3):       # v extends u to the right: prefix of v == suffix of u
            assert np.array_equal(uf[lu - ov:], vs[:ov]), x
        else:                 # v extends u to the left: suffix of v == prefix of u
            assert np.array_equal(vs[lv - ov:], uf[:ov]), x
        assert int(x["delta_twin"]) == lu - (lv - d)
    return len(pick)


@pytest.mark.parametrize("name", ["cfg3-40", "cfg3-60", "cfg3-90"])
def test_cfg3_k_sweep_full_size(name, tmp_path):
    """BASELINE config #3 (k = 40 / 60 / 90 on the cfg2 reads) at full size: the unmodified reference's counters,
    digests and `.graph3` bytes, plus oracle-independent properties."""
    reads, k = _get(name)
    b, off = synth.concat(reads)
    gpu = api.Sage2Gpu(0)
    gpu.run_steps123(b, off, k)
    c = gpu.counters()
    assert c["hash_len"] == min(k, 64) and c["good_reads"] == len(reads)
    assert c["window_probes"] == c["unique_reads"] * (150 - min(k, 64) + 1)
    _check_against_big_golden(gpu, name)
    gpu.write_graph3(str(tmp_path / "g.graph3"))
    assert _md5_file(tmp_path / "g.graph3") == BIG[name]["graph3_md5"]
    assert _check_edges_are_exact_overlaps(gpu, k) > 0


def test_small_sets_edges_are_exact_overlaps():
    for name in ("err", "varlen_err", "tandem", "hicopy"):
        reads, k = _get(name)
        b, off = synth.concat(reads)
        gpu = api.Sage2Gpu(0)
        gpu.run_steps123(b, off, k)
        _check_edges_are_exact_overlaps(gpu, k, sample=1500)


def test_streamed_upload_equals_one_shot():
    reads, k = _get("mixed")
    b, off = synth.concat(reads)
    gpu = api.Sage2Gpu(0)
    gpu.run_steps123(b, off, k)
    e1, c1 = gpu.edges().copy(), gpu.counters()
    gpu.load_reads_chunked(b, off, k, reads_per_chunk=3001)
    gpu.build_hash_table()
    gpu.build_overlap_graph()
    c2 = gpu.counters()
    np.testing.assert_array_equal(e1, gpu.edges())
    for f in ("total_reads", "good_reads", "unique_reads", "total_bp", "compare_calls"):
        assert c1[f] == c2[f], f


@pytest.mark.parametrize("name,on_device", [("err", 1), ("rep", 1), ("tandem", 1), ("varlen_err", None), ("deep_varlen", None), ("hicopy", None), ("mixed", None)])
def test_phase_c_device_and_host_walk_agree(name, on_device, monkeypatch):
    """Phase C runs on the device when the candidate set is symmetric (fixed read length, no masked-key asymmetry) and
    through the host walk otherwise; both must give the oracle's graph, and forcing the walk changes nothing."""
    reads, k = _get(name)
    b, off = synth.concat(reads)
    o = oracle.OracleRun(b, off, k)
    gpu = api.Sage2Gpu(0)
    gpu.run_steps123(b, off, k)
    c = gpu.counters()
    if on_device is not None:
        assert c["phase_c_on_device"] == on_device
    e1 = gpu.edges().copy()
    assert (c["edges_inserted_c"], c["transitive_removed"]) == (o.edges_inserted_c, o.transitive_removed)
    monkeypatch.setenv("SAGE2GPU_PHASE_C_HOST", "1")
    gpu.run_steps123(b, off, k)
    c2 = gpu.counters()
    assert c2["phase_c_on_device"] == 0
    np.testing.assert_array_equal(e1, gpu.edges())
    assert (c2["edges_inserted_c"], c2["transitive_removed"]) == (o.edges_inserted_c, o.transitive_removed)
    for f in ("from", "to", "type", "delta", "delta_twin"):
        np.testing.assert_array_equal(e1[f], o.edges[f])
