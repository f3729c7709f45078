"""GPU tests of the C++ host program: `sage2gpu -f <fastq> -k K -M 3` must write the same bytes as the
unmodified reference's `SAGE2 -s -M 3` (md5s in tests/golden/golden.json, fixtures for `mixed`), and the
reference's own steps 4-7 continued from our files must give byte-identical contig / scaffold FASTA."""
import gzip
import hashlib
import json
import os
import subprocess

import pytest

import datasets
from oracle import oracle
from sage2_b200 import api, synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))
BIN = os.path.join(os.path.dirname(api.LIB_PATH), "sage2gpu")


def _md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def _get(name):
    return datasets.get(name) if name in datasets.DATASETS else synth.config(name)


@pytest.mark.parametrize("name", ["mixed", "rep", "varlen_err", "k70", "cfg1"])
def test_cli_files_byte_identical_to_reference(name, tmp_path):
    reads, k = _get(name)
    fq = tmp_path / "in.fastq"
    synth.write_fastq(str(fq), reads)
    out = tmp_path / "out"
    subprocess.run([BIN, "-f", str(fq), "-k", str(k), "-o", str(out), "-p", "g", "-M", "3"], check=True)
    assert _md5(out / "g.reads") == GOLD[name]["reads_md5"]
    assert _md5(out / "g.graph3") == GOLD[name]["graph3_md5"]
    if name == "mixed":      # the complete files the reference wrote are committed fixtures
        for ext in (".reads", ".graph3"):
            want = gzip.open(os.path.join(HERE, "golden", "mixed" + ext + ".gz"), "rb").read()
            assert open(out / ("g" + ext), "rb").read() == want
    log = open(out / "g.log").read()
    assert f"Number of unique reads: {GOLD[name]['unique_reads']}" in log


def test_cli_list_input_two_mate_files(tmp_path):
    reads, k = _get("rep")
    lst = synth.to_list(reads)
    for mate in (0, 1):
        synth.write_fastq(str(tmp_path / f"m{mate + 1}.fastq"), lst[mate::2])
    (tmp_path / "in.list").write_text(f"f1={tmp_path}/m1.fastq\nf2={tmp_path}/m2.fastq\n")
    out = tmp_path / "out"
    subprocess.run([BIN, "-l", str(tmp_path / "in.list"), "-k", str(k), "-o", str(out), "-p", "g", "-M", "3"], check=True)
    assert _md5(out / "g.reads") == GOLD["rep"]["reads_md5"]
    assert _md5(out / "g.graph3") == GOLD["rep"]["graph3_md5"]


@pytest.mark.skipif(not os.access(oracle.REF_SAGE2, os.X_OK), reason="the compiled reference did not travel (oracle/_ref/SAGE2)")
def test_downstream_fasta_identical_through_reference_steps_4_to_7(tmp_path):
    """north_star: byte-identical contig and scaffold FASTA downstream.  Both arms restart at step 4 (the
    reference's restart path differs from its straight-through path)."""
    reads, k = _get("cfg4mini")
    fq = tmp_path / "in.fastq"
    synth.write_fastq(str(fq), reads)
    ours, ref = tmp_path / "ours", tmp_path / "ref"
    env = dict(os.environ, OMP_NUM_THREADS="1")
    subprocess.run([BIN, "-f", str(fq), "-k", str(k), "-o", str(ours), "-p", "g", "--reference", oracle.REF_SAGE2], check=True, env=env)
    subprocess.run([oracle.REF_SAGE2, "-f", str(fq), "-k", str(k), "-o", str(ref), "-p", "g", "-s", "-M", "3"], check=True, env=env,
                   stdout=subprocess.DEVNULL)
    subprocess.run([oracle.REF_SAGE2, "-f", str(fq), "-k", str(k), "-o", str(ref), "-p", "g", "-m", "4"], check=True, env=env,
                   stdout=subprocess.DEVNULL)
    for ext in (".reads", ".graph3"):
        assert _md5(ours / ("g" + ext)) == _md5(ref / ("g" + ext))
    fastas = sorted(f for f in os.listdir(ref) if f.endswith(".fasta"))
    assert fastas, "the reference wrote no FASTA"
    for f in fastas:
        assert open(ours / f, "rb").read() == open(ref / f, "rb").read(), f
