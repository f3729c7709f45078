#!/usr/bin/env python
"""Random check of the oracle against the UNMODIFIED reference (build container only; needs oracle/_ref from
`make -C oracle ref`):   python tests/golden/fuzz_oracle.py [first_seed] [n]

For every seed: a random read set (genome 40-90 kbp, read length 70-140, k 25-95, coverage 35-70x, optional interspersed /
tandem repeats, substitution errors, variable lengths), `SAGE2 -s -M 3` with one thread, and the oracle on the same reads;
the two `.reads` and the two `.graph3` must be byte-identical.  Then 1,500 mapping queries per data set: the oracle's
sgo_map_reads against the reference's own ReadLoader::getIdOfRead (oracle/_ref/ref_mapids).
Round 1: seeds 700-729 (steps 1-3) and 500-539 (mapping): no mismatch."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import datasets  # noqa: E402
from oracle import oracle  # noqa: E402
from sage2_b200 import synth  # noqa: E402


def md5(p):
    return hashlib.md5(open(p, "rb").read()).hexdigest()


def case(seed):
    rng = np.random.default_rng(seed)
    G = int(rng.integers(40_000, 90_000))
    L = int(rng.integers(70, 140))
    k = int(rng.integers(25, min(L - 5, 95)))
    cov = float(rng.integers(35, 70))
    g = synth.random_genome(G, seed)
    kind = int(rng.integers(0, 4))
    if kind == 1:
        g = synth.add_repeats(g, int(rng.integers(3, 150)), int(rng.integers(k, 3 * L)), seed + 1)
    elif kind == 2:
        g = synth.add_tandem(g, int(rng.integers(3, 40)), int(rng.integers(10, 80)), int(rng.integers(0, G // 2)), seed + 2)
    err = [0.0, 0.0, 0.004, 0.012][int(rng.integers(0, 4))]
    reads = synth.paired_reads(g, L, cov, seed=seed + 3, mu=3 * L, sigma=L // 5, err_rate=err)
    if rng.random() < 0.4:
        reads = synth.variable_length(reads, max(k - 5, L // 2), seed=seed + 4)
    return reads, k


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 700
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    mapids = os.path.join(ROOT, "oracle", "_ref", "ref_mapids")
    bad = done = 0
    for seed in range(first, first + n):
        reads, k = case(seed)
        b, off = synth.concat(reads)
        o = oracle.OracleRun(b, off, k)
        if o.U < 12600:          # the reference's prime table needs more unique reads (tests/datasets.py NO_REFERENCE)
            continue
        tmp = tempfile.mkdtemp(prefix="fuzz_")
        fq = os.path.join(tmp, "d.fastq")
        synth.write_fastq(fq, reads)
        prefix = oracle.run_reference(fq, k, os.path.join(tmp, "ref"), "g", max_step=3, threads=1, save=True, timeout=600)
        o.write_reads(os.path.join(tmp, "o.reads"))
        o.write_graph3(os.path.join(tmp, "o.graph3"))
        ok = md5(prefix + ".reads") == md5(os.path.join(tmp, "o.reads")) and md5(prefix + ".graph3") == md5(os.path.join(tmp, "o.graph3"))
        rl = synth.to_list(reads) if not isinstance(reads, list) else reads
        datasets.DATASETS["_fuzz"] = lambda r=rl, kk=k: (r, kk)
        queries, _ = datasets.map_queries("_fuzz", n=1500, seed=seed)
        qq = os.path.join(tmp, "q.fastq")
        synth.write_fastq(qq, queries)
        ref_ids = subprocess.check_output([mapids, fq, qq, str(k)]).decode().split()
        qb, qoff = synth.concat(queries)
        ids, good = o.map_reads(qb, qoff, k)
        ok = ok and ["bad" if not gd else str(int(i)) for i, gd in zip(ids, good)] == ref_ids
        done += 1
        if not ok:
            bad += 1
            print("MISMATCH seed", seed, flush=True)
        shutil.rmtree(tmp)
    print(f"compared {done} data sets, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
