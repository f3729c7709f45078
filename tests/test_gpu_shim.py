"""The in-process binding: the reference's OWN main.cpp (unmodified) with steps 1-3 on the GPU.

oracle/_ref/SAGE2_gpu = /root/reference's main.cpp and steps 4-7 + sage2_b200/host/sage2gpuShim.cpp + libsage2gpu
(oracle/Makefile, target ref_gpu; built in the build container, travels to the GPU box).  A straight-through run of it
must write the same contig and scaffold files, byte for byte, as a straight-through run of the unmodified reference
(tests/golden/fasta_golden.json, tests/golden/make_fasta_golden.py): the drop-in test of SURVEY.md section 8(b) option 1,
including OverlapGraph::convertGraph consuming and freeing the lists the shim hands over (overlapGraph.cpp:84-115) and
~ReadLoader freeing the reads one by one (readLoader.cpp:61-71).
"""
import hashlib
import json
import os
import subprocess

import pytest

import datasets
from sage2_b200 import synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "SAGE2_gpu")
GOLD = json.load(open(os.path.join(HERE, "golden", "fasta_golden.json")))


@pytest.mark.skipif(not os.access(BIN, os.X_OK), reason="oracle/_ref/SAGE2_gpu is built where /root/reference exists")
@pytest.mark.parametrize("name", sorted(GOLD))
def test_reference_main_with_gpu_steps123_writes_the_reference_fasta(name, tmp_path):
    reads, k = datasets.get(name) if name in datasets.DATASETS else synth.config(name)
    fq = str(tmp_path / "in.fastq")
    synth.write_fastq(fq, reads)
    out = str(tmp_path / "out") + "/"
    os.makedirs(out)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    subprocess.check_call([BIN, "-f", fq, "-k", str(k), "-o", out, "-p", "g"], env=env, cwd=str(tmp_path),
                          stdout=subprocess.DEVNULL, timeout=1800)
    for f in ("g_contig.fasta", "g_scaffold.fasta", "g_contig.gdl", "g_scaffold.gdl"):
        data = open(out + f, "rb").read()
        assert len(data) == GOLD[name][f + ".bytes"], f
        assert hashlib.md5(data).hexdigest() == GOLD[name][f], f
    log = open(out + "g.log").read()
    assert "Number of unique reads" in log and "Function convertGraph()" in log


@pytest.mark.skipif(not os.access(BIN, os.X_OK), reason="oracle/_ref/SAGE2_gpu is built where /root/reference exists")
def test_shim_saves_the_reference_intermediate_files(tmp_path):
    """-s -M 3 through the reference's own saveReadsInFile / saveOverlapGraphInFile on the structures the shim filled."""
    gold = json.load(open(os.path.join(HERE, "golden", "golden.json")))
    reads, k = datasets.get("mixed")
    fq = str(tmp_path / "in.fastq")
    synth.write_fastq(fq, reads)
    out = str(tmp_path / "out") + "/"
    os.makedirs(out)
    subprocess.check_call([BIN, "-f", fq, "-k", str(k), "-o", out, "-p", "g", "-s", "-M", "3"], cwd=str(tmp_path),
                          stdout=subprocess.DEVNULL, timeout=600)
    assert hashlib.md5(open(out + "g.reads", "rb").read()).hexdigest() == gold["mixed"]["reads_md5"]
    assert hashlib.md5(open(out + "g.graph3", "rb").read()).hexdigest() == gold["mixed"]["graph3_md5"]
