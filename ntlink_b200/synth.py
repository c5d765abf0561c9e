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


# ------------------------------------------------------------------------------------------------------------------
# Counter-based generator (csrc/synth_logic.cuh): the plan (which genome slice every contig / read is) is made here with
# numpy, the bases come from the library -- on the device for the benchmark configurations (Context.synth_*_resident), on
# the host for tests and the CPU arm (host_contigs / host_reads). Same plan + same seed = same bytes on both.
CONTIG_DT = np.dtype([("start", "<u8"), ("len", "<u4"), ("flip", "<u4"), ("n_start", "<u4"), ("n_len", "<u4"), ("pad", "<u4", 2)])
READ_DT = np.dtype([("start", "<u8"), ("len", "<u4"), ("flip", "<u4"), ("id", "<u8")])
ONT_ERR = (2621, 1966, 1966)       # 4 % substitutions, 3 % deletions, 3 % insertions, out of 65536


def plan_assembly(genome_bp, seed, min_len=1000, max_len=200000, max_gap=2000, n_frac=0.0):
    """contigs of a genome of genome_bp bases: lengths log-uniform in [min_len, max_len], gaps uniform in [0, max_gap]
    dropped, 50 % reverse-complemented, names ctg%07d shuffled. Returns (CONTIG_DT array, names)."""
    rng = np.random.default_rng(seed)
    starts, lens = [], []
    pos = 0
    while pos < genome_bp:
        n = max(1024, int(2.5 * (genome_bp - pos) / 40000))
        L = np.exp(rng.uniform(np.log(min_len), np.log(max_len), n)).astype(np.int64)
        g = rng.integers(0, max_gap + 1, n)
        st = pos + np.concatenate(([0], np.cumsum(L + g)[:-1]))
        keep = st < genome_bp
        starts.append(st[keep]); lens.append(L[keep])
        if not keep.all():
            break
        pos = int(st[-1] + L[-1] + g[-1])
    st = np.concatenate(starts); L = np.concatenate(lens)
    L = np.minimum(L, genome_bp - st)
    ok = L >= min_len // 2
    st, L = st[ok], L[ok]
    n = len(st)
    plan = np.zeros(n, CONTIG_DT)
    plan["start"], plan["len"] = st, L
    plan["flip"] = rng.random(n) < 0.5
    if n_frac:
        pick = (rng.random(n) < n_frac) & (L > 400)
        plan["n_start"][pick] = (100 + rng.random(int(pick.sum())) * (L[pick] - 300)).astype(np.uint32)
        plan["n_len"][pick] = rng.integers(10, 100, int(pick.sum()))
    names = [f"ctg{int(i):07d}" for i in rng.permutation(n)]
    return plan, names


def plan_reads(genome_bp, total_bases, seed, first_id=0, mean_len=12000, sigma=0.7, min_len=1000, max_len=200000):
    "reads of about total_bases bases: length lognormal(ln mean_len, sigma) clipped, uniform start, 50 % strand"
    rng = np.random.default_rng(seed)
    max_len = min(max_len, genome_bp)
    lens = []
    tot = 0
    while tot < total_bases:
        n = max(1024, int(1.1 * (total_bases - tot) / (mean_len * np.exp(sigma * sigma / 2))))
        L = np.clip(rng.lognormal(np.log(mean_len), sigma, n), min_len, max_len).astype(np.int64)
        cs = np.cumsum(L)
        cut = int(np.searchsorted(cs, total_bases - tot)) + 1
        lens.append(L[:cut])
        tot += int(cs[min(cut, n) - 1])
    L = np.concatenate(lens)
    n = len(L)
    plan = np.zeros(n, READ_DT)
    plan["len"] = L
    plan["start"] = (rng.random(n) * (genome_bp - L + 1)).astype(np.uint64)
    plan["flip"] = rng.random(n) < 0.5
    plan["id"] = first_id + np.arange(n, dtype=np.uint64)
    return plan


def read_names(plan, prefix="read"):
    return [f"{prefix}{int(i):09d}" for i in plan["id"]]


def host_contigs(seed, plan, names, threads=8):
    "the contigs of a plan as a SeqBatch, generated on the host"
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    plan = np.ascontiguousarray(plan)
    off = np.zeros(len(plan) + 1, np.uint64)
    lib.ntl_synth_host_contigs(seed, plan.ctypes.data, len(plan), off.ctypes.data, None, threads)
    seq = np.full(int(off[-1]) + 64, ord("N"), np.uint8)
    rc = lib.ntl_synth_host_contigs(seed, plan.ctypes.data, len(plan), off.ctypes.data, seq.ctypes.data, threads)
    assert rc == 0
    return SeqBatch(seq[:int(off[-1])], off, names)


def host_reads(seed, plan, err=ONT_ERR, threads=8, prefix="read"):
    "the reads of a plan as a SeqBatch, generated on the host"
    from . import _lib
    lib = _lib.load()
    plan = np.ascontiguousarray(plan)
    off = np.zeros(len(plan) + 1, np.uint64)
    rc = lib.ntl_synth_host_reads(seed, plan.ctypes.data, len(plan), err[0], err[1], err[2], off.ctypes.data, None, threads)
    assert rc == 0
    seq = np.full(int(off[-1]) + 64, ord("N"), np.uint8)
    rc = lib.ntl_synth_host_reads(seed, plan.ctypes.data, len(plan), err[0], err[1], err[2], off.ctypes.data, seq.ctypes.data, threads)
    assert rc == 0
    return SeqBatch(seq[:int(off[-1])], off, read_names(plan, prefix))
