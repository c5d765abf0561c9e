"""Deterministic synthetic assemblies and long reads: iid ACGT genome cut into contigs with
log-uniform lengths, ONT-like reads with substitutions / insertions / deletions. Used by bench.py and the tests;
there is no network for real datasets."""
import numpy as np

from .api import SeqBatch

_ACGT = np.frombuffer(b"ACGT", np.uint8)
_COMP = np.zeros(256, np.uint8)
_COMP[[65, 67, 71, 84]] = [84, 71, 67, 65]


def genome(n_bases, seed):
    return _ACGT[np.random.default_rng(seed).integers(0, 4, size=n_bases, dtype=np.uint8)]


def assembly(gen, seed, min_len=1000, max_len=200000, max_gap=2000, n_frac=0.0):
    """Cut the genome into contigs: lengths log-uniform in [min_len, max_len], gaps uniform in [0, max_gap] dropped,
    50 % reverse-complemented, names ctg%07d shuffled so that name order differs from genome order."""
    rng = np.random.default_rng(seed)
    parts, pos = [], 0
    while pos < len(gen):
        L = int(np.exp(rng.uniform(np.log(min_len), np.log(max_len))))
        seg = gen[pos:pos + L]
        if len(seg) >= min_len // 2:
            if rng.random() < 0.5:
                seg = _COMP[seg[::-1]]
            if n_frac and rng.random() < n_frac and len(seg) > 400:
                seg = seg.copy()
                a = int(rng.integers(100, len(seg) - 200))
                seg[a:a + int(rng.integers(10, 100))] = ord("N")
            parts.append(seg)
        pos += L + int(rng.integers(0, max_gap + 1))
    ids = rng.permutation(len(parts))
    names = [f"ctg{int(i):07d}" for i in ids]
    offs = np.zeros(len(parts) + 1, np.uint64)
    offs[1:] = np.cumsum([len(p) for p in parts])
    return SeqBatch(np.concatenate(parts), offs, names)


def reads(gen, coverage, seed, mean_len=12000, sigma=0.7, min_len=1000, max_len=200000, sub=0.04, ins=0.03,
          dele=0.03, name_prefix="read", first_id=0, max_bases=None):
    """Simulated long reads: length lognormal(ln mean_len, sigma) clipped, uniform start, 50 % strand, iid errors."""
    rng = np.random.default_rng(seed)
    target = int(len(gen) * coverage) if max_bases is None else int(max_bases)
    lens = []
    tot = 0
    while tot < target:
        chunk = np.clip(rng.lognormal(np.log(mean_len), sigma, 4096), min_len, min(max_len, len(gen))).astype(np.int64)
        for L in chunk:
            lens.append(int(L))
            tot += int(L)
            if tot >= target:
                break
    parts = []
    for L in lens:
        a = int(rng.integers(0, len(gen) - L + 1))
        r = gen[a:a + L]
        if rng.random() < 0.5:
            r = _COMP[r[::-1]]
        if sub:
            r = r.copy()
            m = rng.random(L) < sub
            r[m] = _ACGT[(np.searchsorted(_ACGT, r[m]) + rng.integers(1, 4, size=int(m.sum()))) % 4]
        if dele:
            r = r[rng.random(len(r)) >= dele]
        if ins:
            n_ins = int(rng.binomial(len(r), ins))
            if n_ins:
                r = np.insert(r, rng.integers(0, len(r) + 1, size=n_ins), _ACGT[rng.integers(0, 4, size=n_ins)])
        parts.append(r)
    offs = np.zeros(len(parts) + 1, np.uint64)
    offs[1:] = np.cumsum([len(p) for p in parts])
    names = [f"{name_prefix}{first_id + i:09d}" for i in range(len(parts))]
    return SeqBatch(np.concatenate(parts), offs, names)
