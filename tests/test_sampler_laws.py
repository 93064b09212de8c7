"""Laws of the library's direct sampler (pedigreesim_b200/csrc/psim_meiosis.cuh) against the reference.

The free-running mode of the CUDA library does not replay Individual.java's draw order: it samples
the pairing configuration, the chiasmata and the segregation directly from their distributions. Here
that same sampler, compiled for the HOST (oracle/device_twin.hpp, oracle rng mode 1 "twin"), is held
against the oracle's statement-for-statement restatement of the Java driven by java.util.Random
(rng mode 0): two-sample chi-square on everything the model defines - pairing configurations
including the homolog order, segment counts and breakpoint positions per map function and
multivalent kind, double reduction, allele dose, unreduced gametes. tests/test_gpu_meiosis.py then
holds the kernels bit for bit against the twin.

Tolerance (north_star: "within a stated chi-square tolerance"): p > 1e-6 for every statistic; the
seeds are fixed, so the tests are deterministic.
"""
import numpy as np
import pytest
from scipy import stats

import _cases
import _oracle

P_FLOOR = 1e-6


def _engine(rng, ploidy, p0, p1, lengths, centro, pref=None, fq=None, seed=1, **model):
    o = _oracle.Oracle(ploidy, rng=rng, seed=seed, **model)
    o.set_chromosomes(lengths, centro, pref, fq)
    o.set_pedigree(p0, p1)
    o.simulate()
    return o


def _pair(ploidy, p0, p1, lengths, centro, pref=None, fq=None, **model):
    return (_engine(_oracle.RNG_PHILOX, ploidy, p0, p1, lengths, centro, pref, fq, seed=31, **model),
            _engine(_oracle.RNG_JAVA, ploidy, p0, p1, lengths, centro, pref, fq, seed=47, **model))


def _same_law(a, b, what, min_expected=5.0):
    """Two-sample chi-square on two histograms over the same categories (sparse cells pooled)."""
    a = np.asarray(a, float).ravel()
    b = np.asarray(b, float).ravel()
    tot = a + b
    order = np.argsort(tot)
    small = order[np.cumsum(tot[order]) < 2 * min_expected]            # pooled tail
    big = np.setdiff1d(np.arange(len(tot)), small)
    big = big[tot[big] >= 2 * min_expected] if len(big) else big
    small = np.setdiff1d(np.arange(len(tot)), big)
    rows = np.vstack([np.append(a[big], a[small].sum()), np.append(b[big], b[small].sum())])
    rows = rows[:, rows.sum(axis=0) > 0]
    if rows.shape[1] < 2:
        return
    chi2, p, _, _ = stats.chi2_contingency(rows)
    assert p > P_FLOOR, (what, p, rows)


def _config_histogram(o, ploidy, first):
    quad, seq = o.get_meiotic_configs(first=first)
    key = quad.reshape(-1).astype(np.int64)
    for h in range(ploidy):
        key = key * ploidy + seq.reshape(-1, ploidy)[:, h]
    return key


@pytest.mark.parametrize("ploidy,pref,natural,fq", [(4, 0.0, True, 0.0), (4, 0.6, True, 0.0), (6, 0.3, True, 0.0),
                                                    (4, 0.4, False, 0.5), (8, 0.5, False, 0.4)])
def test_pairing_configurations_with_homolog_order(ploidy, pref, natural, fq):
    # S1 of a founder: genomes 0..P/2-1 twice each; quadrivalent count x order of the homologs
    n = 20000 if ploidy <= 4 else 12000
    p0, p1 = _cases.std_pedigree("S1", n)
    t, j = _pair(ploidy, p0, p1, [1.0], [0.5], [pref], [fq], natural_pairing=natural)
    qt, _ = t.get_meiotic_configs(first=1)
    qj, _ = j.get_meiotic_configs(first=1)
    nq = ploidy // 4 + 1
    _same_law(np.bincount(qt.ravel(), minlength=nq), np.bincount(qj.ravel(), minlength=nq), "quadrivalents")
    if ploidy == 4:   # the full configuration: 2 x 24 categories
        kt, kj = _config_histogram(t, ploidy, 1), _config_histogram(j, ploidy, 1)
        cats = np.union1d(kt, kj)
        _same_law([(kt == c).sum() for c in cats], [(kj == c).sum() for c in cats], "configuration")
    else:             # where each homolog lands (position in chromseq), per homolog
        _, st = t.get_meiotic_configs(first=1)
        _, sj = j.get_meiotic_configs(first=1)
        for h in range(ploidy):
            _same_law(np.bincount(np.argmax(st.reshape(-1, ploidy) == h, axis=1), minlength=ploidy),
                      np.bincount(np.argmax(sj.reshape(-1, ploidy) == h, axis=1), minlength=ploidy), ("position of homolog", h))


def test_pairing_of_recombined_parents_uses_both_chromosome_ends():
    # F2-like: the parents of the last generation are recombinants, so genome numbers differ between
    # head and tail and the primary side, forced cuts and preferential partners all matter
    n = 6000
    p0 = [-1, -1] + [0] * 40 + [int(2 + k % 40) for k in range(n)]
    p1 = [-1, -1] + [1] * 40 + [int(2 + (k * 7 + 3) % 40) for k in range(n)]
    for natural, fq in ((True, 0.0), (False, 0.7)):
        t = _engine(_oracle.RNG_PHILOX, 4, p0, p1, [1.5], [0.6], [0.7], [fq], seed=5, natural_pairing=natural)
        # the same parents for the Java-order engine: copy the twin's first generations
        so, fo, po = t.get_haplostructs(first=0, count=42)
        j = _oracle.Oracle(4, rng=_oracle.RNG_JAVA, seed=9, natural_pairing=natural)
        j.set_chromosomes([1.5], [0.6], [0.7], [fq])
        j.set_pedigree(p0, p1)
        j.set_haplostructs(0, so, fo, po)
        j.simulate(7)
        kt, kj = _config_histogram(t, 4, 42), _config_histogram(j, 4, 42)
        cats = np.union1d(kt, kj)
        _same_law([(kt == c).sum() for c in cats], [(kj == c).sum() for c in cats], ("configuration", natural))
        for o in (t, j):
            o.set_map([0, 3], [0.0, 0.7, 1.5])
        ft = t.emit(_oracle.EMIT_FOUNDER, ind_first=42).reshape(3, -1)
        fj = j.emit(_oracle.EMIT_FOUNDER, ind_first=42).reshape(3, -1)
        for l in range(3):
            _same_law(np.bincount(ft[l], minlength=8), np.bincount(fj[l], minlength=8), ("founder alleles", natural, l))


MODELS = [  # ploidy, fq, parallel, paired, kosambi, allow_no_recomb
    (2, 0.0, 0.0, 0.0, False, True), (2, 0.0, 0.0, 0.0, True, True), (2, 0.0, 0.0, 0.0, False, False), (2, 0.0, 0.0, 0.0, True, False),
    (4, 1.0, 0.0, 0.0, False, True), (4, 1.0, 0.0, 0.0, True, True), (4, 1.0, 0.0, 1.0, True, False),
    (4, 1.0, 1.0, 0.0, False, True), (4, 1.0, 1.0, 0.5, True, False), (4, 1.0, 0.5, 0.5, False, True),
]


@pytest.mark.parametrize("ploidy,fq,parallel,paired,kosambi,anr", MODELS)
def test_chiasmata_segregation_and_dose(ploidy, fq, parallel, paired, kosambi, anr):
    n = 20000
    p0, p1 = _cases.std_pedigree("S1", n)
    L, centro = 1.1, 0.3
    t, j = _pair(ploidy, p0, p1, [L], [centro], [0.0], [fq], natural_pairing=False, kosambi=kosambi, allow_no_recomb=anr,
                 parallel_multivalents=parallel, paired_centromeres=paired)
    hist, where, first_bp = [], [], []
    for o in (t, j):
        so, fo, po = o.get_haplostructs(first=1)
        nseg = np.diff(so)
        hist.append(np.bincount(nseg, minlength=12)[:12])
        inner = np.ones(len(po), bool)
        inner[so[:-1]] = False
        where.append(np.histogram(po[inner], bins=22, range=(0.0, L))[0])
        fb = po[so[:-1][nseg > 1] + 1]                                  # first breakpoint of every recombined homolog
        first_bp.append(np.histogram(fb, bins=[0, 0.02, 0.05, 0.1, 0.2, 0.4, 0.7, L])[0])
    _same_law(hist[0], hist[1], "segments per homolog")
    _same_law(where[0], where[1], "breakpoint positions")
    _same_law(first_bp[0], first_bp[1], "first breakpoint (stationary start of the chiasma process)")
    pos = np.array([0.0, 0.15, 0.3, 0.6, 1.1])
    tabs = []
    for o in (t, j):
        o.set_map([0, len(pos)], pos)
        tabs.append(o.emit(_oracle.EMIT_FOUNDER, ind_first=1).reshape(len(pos), n, 2, ploidy // 2))
    for l in range(len(pos)):
        _same_law(np.bincount(tabs[0][l].ravel(), minlength=ploidy), np.bincount(tabs[1][l].ravel(), minlength=ploidy), ("allele", l))
        dose = [np.bincount((x[l] == 0).sum(axis=(1, 2)), minlength=ploidy + 1) for x in tabs]
        _same_law(dose[0], dose[1], ("dose", l))
        if ploidy == 4:
            dr = [int((x[l][:, :, 0] == x[l][:, :, 1]).sum()) for x in tabs]
            _same_law([dr[0], 2 * n - dr[0]], [dr[1], 2 * n - dr[1]], ("double reduction", l))
    for a in range(len(pos)):      # recombination between loci, per gamete slot
        for b in range(a + 1, len(pos)):
            rec = [int((x[a] != x[b]).sum()) for x in tabs]
            tot = tabs[0][a].size
            _same_law([rec[0], tot - rec[0]], [rec[1], tot - rec[1]], ("recombination", a, b))


@pytest.mark.parametrize("kind", [1, 2])
def test_unreduced_gametes(kind):
    n = 15000
    p0, p1 = _cases.std_pedigree("S1", n)
    t, j = _pair(2, p0, p1, [1.0], [0.1], unreduced_gametes=kind)
    pos = np.array([0.0, 0.1, 0.5, 1.0])
    het = []
    for o in (t, j):
        o.set_map([0, len(pos)], pos)
        f = o.emit(_oracle.EMIT_FOUNDER, ind_first=1).reshape(len(pos), n, 2)
        het.append([(f[l][:, 0] != f[l][:, 1]).sum() for l in range(len(pos))])
    for l in range(len(pos)):
        _same_law([het[0][l], n - het[0][l]], [het[1][l], n - het[1][l]], ("heterozygous 2n gametes", kind, l))
    # FDR keeps the parent's heterozygosity at the centromere, SDR loses it there
    assert (het[0][1] > 0.95 * n) if kind == 1 else (het[0][1] < 0.05 * n)


def test_the_twin_uses_the_product_stream_and_it_is_philox():
    """The stream of psim_rng.cuh (compiled for the host) yields the words of the oracle's own
    restatement of Philox4x32-10, which tests/test_oracle_pins.py holds against Random123's vectors."""
    import ctypes as C
    L = _oracle.lib()
    a = np.zeros(64, np.uint32)
    b = np.zeros(64, np.uint32)
    for seed, rep, ind, chrom, par, stage in [(12345, 0, 7, 3, 1, 0), (0xFFFFFFFF, 17, 123456, 11, 0, 5), (1, 2, 3, 4, 1, 1021)]:
        L.orc_philox_stream(C.c_uint32(seed), rep, ind, chrom, par, stage, 64, a.ctypes.data_as(C.c_void_p))
        L.orc_twin_stream(C.c_uint64(seed), C.c_uint32(rep), C.c_uint32(ind), C.c_uint32(chrom), C.c_uint32(par), C.c_uint32(stage), 64,
                          b.ctypes.data_as(C.c_void_p))
        assert np.array_equal(a, b)
    # the upper half of a 64-bit seed selects other streams
    L.orc_twin_stream(C.c_uint64(12345 + (1 << 32)), C.c_uint32(0), C.c_uint32(7), C.c_uint32(3), C.c_uint32(1), C.c_uint32(0), 64,
                      a.ctypes.data_as(C.c_void_p))
    L.orc_twin_stream(C.c_uint64(12345), C.c_uint32(0), C.c_uint32(7), C.c_uint32(3), C.c_uint32(1), C.c_uint32(0), 64,
                      b.ctypes.data_as(C.c_void_p))
    assert not np.array_equal(a, b)
