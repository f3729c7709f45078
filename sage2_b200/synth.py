"""Seeded synthetic genomes and paired-end reads for the parity tests and bench.py.

The shapes follow BASELINE.json `configs` / SURVEY.md section 8(d): a uniform-random genome over
{A,C,G,T}, error-free fixed-length paired-end reads written as interleaved FASTQ (quality all 'I'),
fragment start uniform, insert size ~ N(mu, sigma) clipped to >= 2L, mate 2 = reverse complement of
the fragment end.  Extra knobs (repeats, substitution errors, variable read length, palindromes,
N bases, lower case) produce the adversarial inputs the reference's code paths handle
(economyGraph.cpp:443,735; hashTable.cpp:116-121; utils.cpp:144-166).

Everything is numpy `default_rng(seed)`; the same arguments always give the same bytes.
"""
from __future__ import annotations

import numpy as np

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


def random_genome(size: int, seed: int) -> np.ndarray:
    """Uniform random genome as codes 0..3 (A C G T)."""
    return np.random.default_rng(seed).integers(0, 4, size=size, dtype=np.uint8)


def add_repeats(genome: np.ndarray, copies: int, rep_len: int, seed: int) -> np.ndarray:
    """Overwrite `copies` seeded sites with one random `rep_len` element (interspersed repeat)."""
    rng = np.random.default_rng(seed)
    g = genome.copy()
    element = rng.integers(0, 4, size=rep_len, dtype=np.uint8)
    sites = rng.integers(0, len(g) - rep_len, size=copies)
    for s in sites:
        g[s:s + rep_len] = element
    return g


def add_tandem(genome: np.ndarray, unit_len: int, n_units: int, pos: int, seed: int) -> np.ndarray:
    """Insert a tandem repeat (unit repeated n_units times) at pos (overwrites)."""
    rng = np.random.default_rng(seed)
    g = genome.copy()
    unit = rng.integers(0, 4, size=unit_len, dtype=np.uint8)
    rep = np.tile(unit, n_units)
    g[pos:pos + len(rep)] = rep[: max(0, len(g) - pos)]
    return g


def paired_reads(genome: np.ndarray, read_len: int, coverage: float, seed: int,
                 mu: float = 400.0, sigma: float = 30.0, err_rate: float = 0.0) -> np.ndarray:
    """Return an (N, read_len) uint8 array of ASCII reads, mates interleaved (/1,/2,/1,/2...)."""
    rng = np.random.default_rng(seed)
    G = len(genome)
    n_pairs = int(G * coverage / (2 * read_len))
    ins = np.rint(rng.normal(mu, sigma, size=n_pairs)).astype(np.int64)
    ins = np.clip(ins, 2 * read_len, G)
    start = (rng.random(n_pairs) * (G - ins + 1)).astype(np.int64)
    # whole fragments are sampled from either strand with equal probability
    flip = rng.random(n_pairs) < 0.5
    ar = np.arange(read_len, dtype=np.int64)
    codes = np.empty((2 * n_pairs, read_len), dtype=np.uint8)
    step = 1 << 20          # pairs per block: bounds the index temporaries (same bytes as one big gather)
    for p0 in range(0, n_pairs, step):
        p1 = min(n_pairs, p0 + step)
        st, en = start[p0:p1], (start[p0:p1] + ins[p0:p1] - 1)
        m1 = genome[st[:, None] + ar[None, :]]
        # mate 2: reverse complement of the last read_len bases of the fragment
        m2 = _COMP[genome[en[:, None] - ar[None, :]]]
        f = flip[p0:p1, None]
        codes[2 * p0:2 * p1:2] = np.where(f, m2, m1)
        codes[2 * p0 + 1:2 * p1:2] = np.where(f, m1, m2)
    if err_rate > 0.0:
        errs = rng.random(codes.shape) < err_rate
        shift = rng.integers(1, 4, size=codes.shape, dtype=np.uint8)
        codes = np.where(errs, (codes + shift) & 3, codes).astype(np.uint8)
    return _ASCII[codes]


def variable_length(reads: np.ndarray, min_len: int, seed: int) -> list[bytes]:
    """Trim each read to a seeded length in [min_len, L] (creates containments)."""
    rng = np.random.default_rng(seed)
    L = reads.shape[1]
    lens = rng.integers(min_len, L + 1, size=len(reads))
    return [bytes(r[:n]) for r, n in zip(reads, lens)]


def to_list(reads) -> list[bytes]:
    if isinstance(reads, np.ndarray):
        return [bytes(r) for r in reads]
    return list(reads)


def write_fastq(path: str, reads) -> None:
    """Interleaved FASTQ, names r<i>/1 r<i>/2, quality 'I'."""
    if isinstance(reads, np.ndarray) and reads.ndim == 2:
        n, L = reads.shape
        qual = b"I" * L
        with open(path, "wb") as f:
            chunk = []
            for i in range(n):
                chunk.append(b"@r%d/%d\n" % (i // 2, 1 + (i & 1)))
                chunk.append(reads[i].tobytes())
                chunk.append(b"\n+\n")
                chunk.append(qual)
                chunk.append(b"\n")
                if len(chunk) >= 50000:
                    f.write(b"".join(chunk))
                    chunk = []
            f.write(b"".join(chunk))
        return
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            r = bytes(r)
            f.write(b"@r%d/%d\n%s\n+\n%s\n" % (i // 2, 1 + (i & 1), r, b"I" * len(r)))


def concat(reads) -> tuple[np.ndarray, np.ndarray]:
    """(bases, offsets): one uint8 ASCII buffer + int64 offsets[n+1] (the C-ABI input form)."""
    if isinstance(reads, np.ndarray) and reads.ndim == 2:
        n, L = reads.shape
        return np.ascontiguousarray(reads).reshape(-1), np.arange(n + 1, dtype=np.int64) * L
    lens = np.fromiter((len(r) for r in reads), dtype=np.int64, count=len(reads))
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    buf = np.frombuffer(b"".join(bytes(r) for r in reads), dtype=np.uint8).copy() if len(reads) else np.zeros(0, np.uint8)
    return buf, off


# ---- named configurations (BASELINE.json `configs`) ------------------------------------------

def config(name: str):
    """Return (reads, k) for a named workload.

    cfg1  : 1 Mbp, 100 bp, 40x, k=50                      (reference CPU run end to end)
    cfg2  : 4.6 Mbp, 150 bp, 100x, k=63                   (single B200; the bench workload)
    cfg3-40/60/90 : cfg2 reads with k = 40 / 60 / 90
    cfg4  : 100 Mbp + 2% interspersed repeats, 150 bp, 50x, k=75
    cfg4mini : 1 Mbp + 20 x 1 kb repeat, 150 bp, 50x, k=75 (cfg4-shaped, oracle finishes in seconds)
    """
    if name == "cfg1":
        g = random_genome(1_000_000, 12345)
        return paired_reads(g, 100, 40, seed=12346, mu=400, sigma=30), 50
    if name == "cfg2" or name.startswith("cfg3-"):
        g = random_genome(4_600_000, 4600)
        reads = paired_reads(g, 150, 100, seed=4601, mu=450, sigma=30)
        k = 63 if name == "cfg2" else int(name.split("-")[1])
        return reads, k
    if name == "cfg4":
        g = add_repeats(random_genome(100_000_000, 100), 2000, 1000, 101)
        return paired_reads(g, 150, 50, seed=102, mu=450, sigma=30), 75
    if name == "cfg4mini":
        g = add_repeats(random_genome(1_000_000, 40), 20, 1000, 41)
        return paired_reads(g, 150, 50, seed=42, mu=450, sigma=30), 75
    raise KeyError(name)


def config_cached(name: str, wait=None):
    """config(name) through a per-box cache (/dev/shm or the temp directory): the big workloads take a minute to generate
    and are used by several tests and by every bench.py run.  Returns (reads memory-mapped read-only, k).
    wait: optional callable run between "the file is written" and "the file is read" (a barrier when several ranks share
    the box and only one of them generates)."""
    import json
    import os
    import tempfile
    best = None
    for d in ("/dev/shm", tempfile.gettempdir()):
        try:
            st = os.statvfs(d)
            if st.f_bavail * st.f_frsize > (12 << 30) and os.access(d, os.W_OK):
                best = d
                break
        except OSError:
            pass
    if best is None:
        if wait:
            wait()
        return config(name)
    path = os.path.join(best, f"sage2_synth_{name}_v2.npy")
    meta = path + ".json"
    if (wait is None or os.environ.get("RANK", "0") == "0") and not (os.path.exists(path) and os.path.exists(meta)):
        reads, k = config(name)
        tmp = path + f".{os.getpid()}.tmp.npy"
        np.save(tmp, reads)
        os.replace(tmp, path)
        with open(meta + ".tmp", "w") as f:
            json.dump({"k": k}, f)
        os.replace(meta + ".tmp", meta)
        del reads
    if wait:
        wait()
    with open(meta) as f:
        k = json.load(f)["k"]
    return np.load(path, mmap_mode="r"), k
