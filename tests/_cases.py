"""Shared scenario builders for the parity tests: the same inputs are pushed through the CPU
oracle (tests/_oracle.py) and the CUDA library (pedigreesim_b200) with identical calls."""
import numpy as np

import _oracle


def std_pedigree(poptype, popsize):
    """Main.generateStandardPopulation (Main.java:1466-1502) as parent index arrays."""
    if poptype == "F1":
        p0 = [-1, -1] + [0] * popsize
        p1 = [-1, -1] + [1] * popsize
    elif poptype == "F2":
        p0 = [-1, -1, 0] + [2] * popsize
        p1 = [-1, -1, 1] + [2] * popsize
    elif poptype == "BC":
        p0 = [-1, -1, 0] + [0] * popsize
        p1 = [-1, -1, 1] + [2] * popsize
    elif poptype == "S1":
        p0 = [-1] + [0] * popsize
        p1 = [-1] + [0] * popsize
    else:
        raise ValueError(poptype)
    return np.array(p0, np.int32), np.array(p1, np.int32)


def ril_pedigree(lines, generations):
    """P1 x P2 -> F1 -> F2(lines) -> ... selfing chain (BASELINE config 4 at reduced size)."""
    p0, p1 = [-1, -1, 0], [-1, -1, 1]
    prev = [2] * lines
    for _ in range(generations - 1):
        cur = []
        for k in range(lines):
            cur.append(len(p0))
            p0.append(prev[k])
            p1.append(prev[k])
        prev = cur
    return np.array(p0, np.int32), np.array(p1, np.int32)


def random_pedigree(rng, n_founders, n_total):
    p0 = [-1] * n_founders
    p1 = [-1] * n_founders
    for i in range(n_founders, n_total):
        p0.append(int(rng.integers(0, i)))
        p1.append(int(rng.integers(0, i)))
    return np.array(p0, np.int32), np.array(p1, np.int32)


def uniform_map(lengths_M, loci_per_chrom):
    off = [0]
    pos = []
    for ln, n in zip(lengths_M, loci_per_chrom):
        pos.extend((ln * (np.arange(n) + 0.5) / n).tolist() if n else [])
        off.append(off[-1] + n)
    return np.array(off, np.int64), np.array(pos, np.float64)


def hash_codes(n_loci, n_alleles, n_codes=2, salt=0):
    """Deterministic pseudo-random allele codes; every locus uses codes 0..k-1 without gaps
    (codes are ranks among the unique names, Locus.java:88-103)."""
    l = np.arange(n_loci, dtype=np.uint64)[:, None]
    a = np.arange(n_alleles, dtype=np.uint64)[None, :]
    h = (l * np.uint64(0x9E3779B97F4A7C15) + a * np.uint64(0xC2B2AE3D27D4EB4F) + np.uint64(salt)) >> np.uint64(29)
    raw = (h % np.uint64(n_codes)).astype(np.uint8)
    out = np.zeros_like(raw)
    for i in range(n_loci):  # re-rank so the codes present are 0..k-1
        u = np.unique(raw[i])
        out[i] = np.searchsorted(u, raw[i])
    return out


def random_haplostructs(rng, n_ind, lengths_M, ploidy, n_alleles, max_segments=6, on_grid=None):
    """Random HaploStructs in .hsa order; breakpoints may coincide with loci (on_grid) or repeat
    founders in neighbouring segments, like the reference's unmerged lists."""
    seg_off = [0]
    founder = []
    pos = []
    for _ in range(n_ind):
        for ln in lengths_M:
            for _h in range(ploidy):
                n = int(rng.integers(1, max_segments + 1))
                if on_grid is not None and n > 1 and rng.random() < 0.5:
                    bp = np.sort(rng.choice(on_grid[(on_grid > 0) & (on_grid < ln)], size=n - 1, replace=True))
                else:
                    bp = np.sort(rng.random(n - 1) * ln)
                pos.extend([0.0] + bp.tolist())
                founder.extend(rng.integers(0, n_alleles, size=n).tolist())
                seg_off.append(seg_off[-1] + n)
    return np.array(seg_off, np.int64), np.array(founder, np.int32), np.array(pos, np.float64)


def make_engine(kind, ploidy, **kw):
    """kind: 'oracle-java' | 'oracle-philox' | 'gpu'"""
    if kind == "gpu":
        import pedigreesim_b200
        return pedigreesim_b200.PsimContext(ploidy, **kw)
    rng = _oracle.RNG_JAVA if kind == "oracle-java" else _oracle.RNG_PHILOX
    kw = dict(kw)
    kw.pop("device", None)
    rep_first = kw.pop("replicate_first", 0)
    assert kw.pop("replicate_count", 1) == 1
    return _oracle.Oracle(ploidy, rng=rng, replicate=rep_first, **kw)


def oracle_emit_all(o, which=0, with_codes=False):
    out = {"founder": o.emit(_oracle.EMIT_FOUNDER), "dose": o.emit(_oracle.EMIT_DOSE, which=which)}
    if with_codes:
        out["allele"] = o.emit(_oracle.EMIT_ALLELE)
    return out
