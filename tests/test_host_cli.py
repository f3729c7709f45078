"""CPU checks of the C++ host program (sage2_b200/host): the FASTA/Q parser and list grammar against a
Python restatement of the record format, SAGE2's option handling (main.cpp:384-521), and the loud
failure without a CUDA device."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

from sage2_b200 import api, synth

BIN = os.path.join(os.path.dirname(api.LIB_PATH), "sage2gpu")


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(BIN) or not os.path.exists(api.LIB_PATH):
        api.build_library()


def fnv(seqs):
    h = 1469598103934665603
    for s in seqs:
        for c in s:
            h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        h = ((h ^ 0xFF) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def parse_only(*args):
    r = subprocess.run([BIN, *args, "-k", "20", "--parse-only"], capture_output=True, text=True, check=True)
    return json.loads(r.stdout)


def expect(seqs):
    return {"reads": len(seqs), "bases": sum(len(s) for s in seqs), "fnv1a": fnv(seqs)}


def test_fastq_fasta_multiline_and_gzip(tmp_path):
    rng = np.random.default_rng(3)
    seqs = [bytes(rng.choice(list(b"ACGTN"), size=int(n)).astype(np.uint8)) for n in rng.integers(1, 300, size=400)]
    fq, fa = tmp_path / "a.fastq", tmp_path / "a.fa"
    with open(fq, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b"@r%d/%d some comment\n%s\n+\n%s\n" % (i // 2, i % 2 + 1, s, b"@" * len(s)))     # '@' in the quality line
    with open(fa, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">r%d\n" % i)
            for p in range(0, len(s), 60):
                f.write(s[p:p + 60] + b"\r\n")
            f.write(b"\n")
    assert parse_only("-f", str(fq)) == expect(seqs)
    assert parse_only("-f", str(fa)) == expect(seqs)
    gz = tmp_path / "a.fastq.gz"
    with gzip.open(gz, "wb") as f:
        f.write(open(fq, "rb").read())
    assert parse_only("-f", str(gz)) == expect(seqs)


def test_multiline_fastq_and_truncated_quality(tmp_path):
    p = tmp_path / "m.fastq"
    p.write_bytes(b"@a\nACGT\nACG\n+\nIIII\nIII\n@b\nTTTT\n+b\nII\n")      # record b: quality shorter than sequence
    assert parse_only("-f", str(p)) == expect([b"ACGTACG"])


def test_list_input_pairs_and_interleaved(tmp_path):
    m1 = [b"AAAA", b"CCCC", b"GGGG"]
    m2 = [b"TT", b"GG"]                                    # shorter mate file: reading stops when it ends
    single = [b"ACGTACGT", b"TTTTAAAA"]
    for name, seqs in (("m1.fa", m1), ("m2.fa", m2), ("s.fa", single)):
        with open(tmp_path / name, "wb") as f:
            for i, s in enumerate(seqs):
                f.write(b">x%d\n%s\n" % (i, s))
    lst = tmp_path / "in.list"
    lst.write_text(f"# comment\nf1 = {tmp_path}/m1.fa\nf2 = {tmp_path}/m2.fa\n\nf={tmp_path}/s.fa\n")
    assert parse_only("-l", str(lst)) == expect([m1[0], m2[0], m1[1], m2[1], m1[2]] + single)


def test_required_options_like_the_reference(tmp_path):
    r = subprocess.run([BIN, "-f", "x.fastq"], capture_output=True, text=True)
    assert r.returncode == 0 and "Option -k|--minOverlap is required" in r.stdout
    r = subprocess.run([BIN, "-k", "40"], capture_output=True, text=True)
    assert "One of the options -f|--fileInput or -l|--listInput is required" in r.stdout
    r = subprocess.run([BIN, "-f", "a", "-l", "b", "-k", "40"], capture_output=True, text=True)
    assert "mutually exclusive" in r.stdout
    r = subprocess.run([BIN, "-f", "a", "-k", "40", "-m", "5", "-M", "3"], capture_output=True, text=True)
    assert "maxStep should not be smaller than minStep" in r.stdout


def test_fails_loudly_without_a_device(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    fq = tmp_path / "r.fastq"
    synth.write_fastq(str(fq), [b"ACGT" * 30] * 4)
    r = subprocess.run([BIN, "-f", str(fq), "-k", "40", "-o", str(tmp_path / "out"), "-M", "3"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no usable CUDA device" in r.stderr and "no CPU fallback" in r.stderr
    assert not os.path.exists(tmp_path / "out" / "untitled.graph3")
