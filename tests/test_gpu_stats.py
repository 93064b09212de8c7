"""Free-running (Philox) statistical parity of the CUDA meiosis path, through the C ABI.

north_star: "the free-running RNG mode must match the reference's per-interval recombination
fractions, quadrivalent and double-reduction rates, and allele-dose distributions within a stated
chi-square tolerance". Tolerance used everywhere here: the test statistic must have p > 1e-6
(two-sided z or chi-square), i.e. a correct implementation fails once in a million runs; the
seeds are fixed, so the tests are deterministic in practice.

Expected values come from (i) the closed forms coded in the reference (Tools.java:116,140),
(ii) manual Tables 1 and 2, (iii) the CPU oracle driven by java.util.Random (the reference's own
draw sequence) for quantities without a closed form (double reduction, dose distributions).
Also: size-independent properties of the emission at BASELINE.json's full configs[1] size.
"""
import numpy as np
import pytest
from scipy import stats

import _cases
import _oracle

pytestmark = pytest.mark.gpu
P_FLOOR = 1e-6


def _gpu(ploidy, p0, p1, lengths, centro, pref=None, fq=None, seed=2024, **model):
    import pedigreesim_b200
    g = pedigreesim_b200.PsimContext(ploidy, seed=seed, **model)
    g.set_chromosomes(lengths, centro, pref, fq)
    g.set_pedigree(p0, p1)
    g.simulate()
    return g


def _java_oracle(ploidy, p0, p1, lengths, centro, pref=None, fq=None, seed=77, **model):
    o = _oracle.Oracle(ploidy, rng=_oracle.RNG_JAVA, seed=seed, **model)
    o.set_chromosomes(lengths, centro, pref, fq)
    o.set_pedigree(p0, p1)
    o.simulate()
    return o


@pytest.mark.parametrize("kosambi", [False, True])
def test_interval_recombination_fractions_follow_the_map_function(kosambi):
    n = 100000
    p0, p1 = _cases.std_pedigree("S1", n)
    g = _gpu(2, p0, p1, [1.2], [0.4], kosambi=kosambi)
    pos = np.array([0.0, 0.02, 0.05, 0.1, 0.2, 0.35, 0.6, 0.9, 1.2])
    g.set_map([0, len(pos)], pos)
    F = g.emit(founder=True, ind_first=1)["founder"].reshape(len(pos), n, 2)
    for a in range(len(pos)):
        for b in range(a + 1, len(pos)):
            d = pos[b] - pos[a]
            expect = 0.5 * np.tanh(2 * d) if kosambi else 0.5 * (1 - np.exp(-2 * d))   # Tools.java:116,140
            r = (F[a] != F[b]).mean()
            se = np.sqrt(expect * (1 - expect) / (2 * n))
            # the reference's Kosambi model is a gamma(2.63) renewal process that only approximates
            # the closed form (Tools.java:317-331): allow its documented bias of a few 1e-3
            tol = stats.norm.isf(P_FLOOR / 2) * se + (0.006 if kosambi else 0.0)
            assert abs(r - expect) < tol, (kosambi, d, r, expect)


def test_mean_breakpoints_and_transmission_probabilities():
    n = 50000
    p0, p1 = _cases.std_pedigree("S1", n)
    L = 1.3
    g = _gpu(4, p0, p1, [L], [0.5], [0.3], [0.0])
    so, _, _ = g.get_haplostructs(first=1)
    # bivalents and cross-type quadrivalents both give about one breakpoint per Morgan and chromatid
    mean_bp = (np.diff(so) - 1).mean()
    assert 0.85 * L < mean_bp < 1.15 * L
    g.set_map([0, 3], [0.1, 0.65, 1.2])
    F = g.emit(founder=True, ind_first=1)["founder"].reshape(3, n, 4)
    for l in range(3):
        for f in range(4):   # each parental homolog reaches a gamete slot with probability 1/2
            cnt = (F[l] == f).sum()
            z = (cnt - 2 * n * 0.5) / np.sqrt(2 * n * 0.25 * 1.5)
            assert abs(z) < 6, (l, f, cnt)


@pytest.mark.parametrize("ploidy,pref", [(4, 0.0), (4, 0.5), (6, 0.25), (8, 0.0)])
def test_natural_pairing_quadrivalent_counts_match_manual_table1(ploidy, pref):
    import json
    import os
    T = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "manual_table1.json")))
    j = T["pref_pairing"].index(pref)
    manual = np.array([r[j] for r in T["counts_by_ploidy"][str(ploidy)]], float)
    n = 20000
    p0, p1 = _cases.std_pedigree("S1", n)
    g = _gpu(ploidy, p0, p1, [1.0], [0.5], [pref], [0.0])
    quad, _ = g.get_meiotic_configs(first=1)
    ours = np.bincount(quad.reshape(-1), minlength=len(manual)).astype(float)
    keep = (ours + manual) > 0
    chi2, p, _, _ = stats.chi2_contingency(np.vstack([ours[keep], manual[keep]]))
    assert p > P_FLOOR, (ploidy, pref, ours, manual)


def test_fixed_quadrivalent_fraction_is_binomial():
    for ploidy, fq in [(4, 0.5), (8, 0.5), (12, 0.3)]:
        n = 20000
        p0, p1 = _cases.std_pedigree("S1", n)
        g = _gpu(ploidy, p0, p1, [1.0], [0.5], [0.1], [fq], natural_pairing=False)
        quad, _ = g.get_meiotic_configs(first=1)
        nq = ploidy // 4
        ours = np.bincount(quad.reshape(-1), minlength=nq + 1).astype(float)
        expect = stats.binom.pmf(np.arange(nq + 1), nq, fq) * ours.sum()   # manual Table 2
        chi2, p = stats.chisquare(ours, expect)
        assert p > P_FLOOR, (ploidy, ours, expect)


@pytest.mark.parametrize("kosambi,paired", [(False, 0.0), (True, 0.0), (False, 1.0)])
def test_double_reduction_and_dose_distribution_match_the_reference_draw_sequence(kosambi, paired):
    # tetraploid, all quadrivalents: gametes can carry the same founder allele twice (double
    # reduction). No closed form exists for the cross-type model (Tools.java:208-211), so the
    # reference's own RNG (oracle, java.util.Random) is the expectation: two-sample chi-square.
    n = 30000
    p0, p1 = _cases.std_pedigree("S1", n)
    args = dict(lengths=[1.0], centro=[0.2], pref=[0.0], fq=[1.0], natural_pairing=False, kosambi=kosambi,
                paired_centromeres=paired)
    g = _gpu(4, p0, p1, **args)
    o = _java_oracle(4, p0, p1, **args)
    pos = np.array([0.0, 0.2, 0.5, 1.0])
    tables = []
    for e, is_gpu in ((g, True), (o, False)):
        e.set_map([0, len(pos)], pos)
        F = e.emit(founder=True, ind_first=1)["founder"] if is_gpu else e.emit(_oracle.EMIT_FOUNDER, ind_first=1)
        tables.append(F.reshape(len(pos), n, 2, 2))     # [locus][child][gamete][slot]
    for l in range(len(pos)):
        dr = [int((t[l][:, :, 0] == t[l][:, :, 1]).sum()) for t in tables]     # double-reduction gametes
        cont = np.array([[dr[0], 2 * n - dr[0]], [dr[1], 2 * n - dr[1]]], float)
        if cont[:, 0].sum() > 0:
            chi2, p, _, _ = stats.chi2_contingency(cont)
            assert p > P_FLOOR, ("double reduction", l, dr)
        # dose of founder allele 0 in the offspring (0..4)
        dose = [np.bincount((t[l] == 0).sum(axis=(1, 2)), minlength=5).astype(float) for t in tables]
        keep = (dose[0] + dose[1]) > 0
        chi2, p, _, _ = stats.chi2_contingency(np.vstack([dose[0][keep], dose[1][keep]]))
        assert p > P_FLOOR, ("dose", l, dose)
    # double reduction grows with the distance from the centromere
    far = (tables[0][3][:, :, 0] == tables[0][3][:, :, 1]).mean()
    near = (tables[0][1][:, :, 0] == tables[0][1][:, :, 1]).mean()
    assert far > near


def test_full_size_configs1_emission_properties():
    """BASELINE.json configs[1] at full size (10 002 individuals x 24 000 loci x 4): properties
    that do not need the oracle (it would take minutes on this size)."""
    import bench
    import pedigreesim_b200
    w = bench.workload()
    g = pedigreesim_b200.PsimContext(w["P"], seed=1, **w["model"])
    g.set_chromosomes(w["lengths"], w["centro"], w["pref"], w["fq"])
    g.set_pedigree(w["p0"], w["p1"])
    g.simulate()
    g.set_map(w["off"], w["pos"])
    g.set_allele_codes(w["code"])
    out = g.emit(founder=True, allele=True, dose=True, dose_which=pedigreesim_b200.DOSE_MAX_ALLELE)
    F, A, D = out["founder"], out["allele"], out["dose"]
    N, L, P = w["N"], w["L"], w["P"]
    assert F.shape == (L, N * P) and D.shape == (L, N)
    # founders: constant columns k*P + p (Individual.calcFounderHaplostruct)
    assert np.all(F[:, :8] == np.arange(8))
    # offspring: first two homologs from P1's alleles, last two from P2's (Gamete.fertilization)
    kids = F[:, 8:].reshape(L, N - 2, 4)
    assert kids[:, :, :2].max() <= 3 and kids[:, :, 2:].min() >= 4
    # allele table = code[l][founder], dose = number of homologs carrying the locus' largest code
    rows = np.arange(L)[:, None]
    assert np.array_equal(A, w["code"][rows, F])
    mx = w["code"].max(axis=1)[:, None]
    assert np.array_equal(D, (A.reshape(L, N, P) == mx[:, :, None]).sum(axis=2))
    # a founder cell never changes more often than the homolog has segments
    so, _, _ = g.get_haplostructs()
    changes = (np.diff(F.astype(np.int16), axis=0) != 0).sum(axis=0)          # per column, over all chromosomes
    nseg = np.diff(so).reshape(N, w["C"], P)
    assert np.all(changes.reshape(N, P) <= (nseg - 1).sum(axis=1) + (w["C"] - 1))
    # idempotence and sub-range consistency
    again = g.emit(founder=True, dose=True, dose_which=pedigreesim_b200.DOSE_MAX_ALLELE)
    assert np.array_equal(again["founder"], F) and np.array_equal(again["dose"], D)
    sub = g.emit(founder=True, dose=True, dose_which=pedigreesim_b200.DOSE_MAX_ALLELE, ind_first=1234, ind_count=777,
                 loc_first=1999, loc_count=4003)
    assert np.array_equal(sub["founder"], F[1999:1999 + 4003, 1234 * P:(1234 + 777) * P])
    assert np.array_equal(sub["dose"], D[1999:1999 + 4003, 1234:1234 + 777])
    # every 1 Morgan-ish chromosome shows about one breakpoint per offspring homolog
    mean_bp = (nseg[2:] - 1).mean()
    assert 0.7 < mean_bp < 1.3


# ------------------------------------------------------------------ device reductions
def _numpy_gamete_stats(F, P, A, unreduced=False):
    """Main.testMeiosis counters (Main.java:1660-1721) from the founder table F [L][n_ind][P]."""
    L, n, _ = F.shape
    p2 = P if unreduced else P // 2                  # homologs per gamete (Main.java:1546-1548)
    G = F.reshape(L, n * (P // p2), p2)              # gametes: the two halves of every individual, or all of it
    allele = np.stack([(G == f).sum(axis=(1, 2)) for f in range(A)])                       # [A][L]
    dose = np.stack([np.stack([((G == f).sum(axis=2) == d).sum(axis=1) for d in range(p2 + 1)], axis=1) for f in range(A)])
    srt = np.sort(G, axis=2)
    homoz = (srt[:, :, 1:] == srt[:, :, :-1]).any(axis=2).sum(axis=1) if p2 > 1 else np.zeros(L, int)
    H = F.reshape(L, n * P)
    rec = np.zeros((L, L), np.int64)
    for a in range(L):
        for b in range(a):
            rec[a, b] = (H[a] != H[b]).sum()
    return allele, dose, homoz, rec


@pytest.mark.parametrize("ploidy,poptype,kw", [
    (2, "F2", {}), (4, "S1", dict(natural_pairing=False)), (4, "F1", dict(natural_pairing=True, kosambi=True)),
    (6, "F1", {}),
])
def test_device_gamete_statistics_match_counts_from_the_tables(ploidy, poptype, kw):
    n = 3000
    p0, p1 = _cases.std_pedigree(poptype, n)
    lengths, centro = [1.1, 0.6], [0.4, 0.3]
    pref, fq = ([0.4, 0.1], [0.6, 0.3]) if ploidy > 2 else (None, None)
    g = _gpu(ploidy, p0, p1, lengths, centro, pref, fq, **kw)
    loci = [37, 21]
    off, pos = _cases.uniform_map(lengths, loci)
    g.set_map(off, pos)
    first = len(p0) - n
    A = ploidy * (p0 < 0).sum()
    F = g.emit(founder=True, ind_first=first)["founder"].reshape(sum(loci), n, ploidy)
    so, fo, po = g.get_haplostructs(first)
    quad, _ = g.get_meiotic_configs(first)
    for c in range(2):
        st = g.gamete_stats(c, loci[c], A, ind_first=first)
        allele, dose, homoz, rec = _numpy_gamete_stats(F[off[c]:off[c + 1]], ploidy, A)
        assert st["n_gametes"] == 2 * n
        assert np.array_equal(st["allele_count"], allele)
        assert np.array_equal(st["dose_count"], dose)
        assert np.array_equal(st["homozygous_count"], homoz)
        assert np.array_equal(st["recomb_count"], rec)
        nseg = np.diff(so).reshape(n, 2, ploidy)[:, c, :].ravel()
        assert np.array_equal(st["segments_hist"], np.bincount(np.minimum(nseg - 1, 63), minlength=64))
        assert np.array_equal(st["quadrivalent_hist"], np.bincount(quad[:, c, :].ravel(), minlength=ploidy // 4 + 1))


def test_test_mode_setting_double_reduction_and_recombination_on_device():
    """The reference's TEST setting: one tetraploid parent, many meioses (an S1 population is two
    gametes of the parent per individual). Quadrivalents allow double reduction: the homozygous
    fraction must match the Java-LCG oracle's; recombination between the chromosome ends must match
    the Haldane value for bivalent pairing."""
    n = 60000
    p0, p1 = _cases.std_pedigree("S1", n)
    lengths, centro = [1.0], [0.3]
    pos = np.array([0.0, 0.25, 0.5, 0.75, 1.0])
    # all quadrivalents
    g = _gpu(4, p0, p1, lengths, centro, [0.0], [1.0], natural_pairing=False)
    g.set_map([0, 5], pos)
    st = g.gamete_stats(0, 5, 4, ind_first=1)
    n_o = 6000
    q0, q1 = _cases.std_pedigree("S1", n_o)
    o = _java_oracle(4, q0, q1, lengths, centro, [0.0], [1.0], natural_pairing=False)
    o.set_map([0, 5], pos)
    Fo = o.emit(_oracle.EMIT_FOUNDER, ind_first=1).reshape(5, n_o * 2, 2)
    ho = (Fo[:, :, 0] == Fo[:, :, 1]).sum(axis=1)
    for l in range(5):
        tab = np.array([[st["homozygous_count"][l], st["n_gametes"] - st["homozygous_count"][l]], [ho[l], 2 * n_o - ho[l]]])
        if tab[:, 0].sum() > 20:
            assert stats.chi2_contingency(tab)[1] > P_FLOOR, (l, tab)
    assert st["homozygous_count"][4] > 0 and st["quadrivalent_hist"].tolist() == [0, 2 * n]
    # all bivalents, no preferential pairing: Haldane between the ends of a 1 M chromosome
    g = _gpu(4, p0, p1, lengths, centro, [0.0], [0.0], natural_pairing=False)
    g.set_map([0, 5], pos)
    st = g.gamete_stats(0, 5, 4, ind_first=1)
    assert st["homozygous_count"].sum() == 0 and st["quadrivalent_hist"].tolist() == [2 * n, 0]
    r = st["recomb_count"][4, 0] / (st["n_gametes"] * 2)
    expect = 0.5 * (1 - np.exp(-2.0))
    # with 4 homologs a crossover partner carries a different founder allele with certainty, and
    # returning to the same founder needs an even number of exchanges with the SAME partner
    se = np.sqrt(0.25 / (st["n_gametes"] * 2))
    assert abs(r - expect) < stats.norm.isf(P_FLOOR / 2) * se


@pytest.mark.parametrize("kind", [1, 2])
def test_unreduced_gametes_on_device_fdr_and_sdr(kind):
    """2NGAMETES (Individual.java:341-383): device counters over 2n gametes (one per individual, G = P
    homologs) against the same counts taken from the emitted table, and the tetrad-analysis closed
    forms of tests/test_oracle_pins.py on a diploid bivalent."""
    n = 20000
    p0, p1 = _cases.std_pedigree("S1", n)
    ce = 0.2
    g = _gpu(2, p0, p1, [1.0], [ce], unreduced_gametes=kind)
    pos = np.array([ce, 0.3, 0.5, 1.0])
    g.set_map([0, 4], pos)
    st = g.gamete_stats(0, 4, 2, ind_first=1)
    assert st["n_gametes"] == n and st["dose_count"].shape == (2, 4, 3)
    F = g.emit(founder=True, ind_first=1)["founder"].reshape(4, n, 2)
    allele, dose, homoz, rec = _numpy_gamete_stats(F, 2, 2, unreduced=True)
    assert np.array_equal(st["allele_count"], allele) and np.array_equal(st["dose_count"], dose)
    assert np.array_equal(st["homozygous_count"], homoz) and np.array_equal(st["recomb_count"], rec)
    het = 1.0 - st["homozygous_count"] / n
    assert het[0] == (1.0 if kind == 1 else 0.0)
    for k in range(1, 4):
        y = 2.0 / 3.0 * (1 - np.exp(-3 * (pos[k] - ce)))
        expect = 1 - y / 2 if kind == 1 else y
        assert abs(het[k] - expect) < stats.norm.isf(P_FLOOR / 2) * np.sqrt(expect * (1 - expect) / n), (kind, k, het[k], expect)
    # tetraploid, all quadrivalents: same counters, table cross-check only
    g = _gpu(4, p0[:3001], p1[:3001], [1.0], [0.4], [0.2], [1.0], natural_pairing=False, unreduced_gametes=kind)
    g.set_map([0, 4], pos)
    st = g.gamete_stats(0, 4, 4, ind_first=1)
    F = g.emit(founder=True, ind_first=1)["founder"].reshape(4, 3000, 4)
    allele, dose, homoz, rec = _numpy_gamete_stats(F, 4, 4, unreduced=True)
    assert st["n_gametes"] == 3000 and np.array_equal(st["dose_count"], dose) and np.array_equal(st["homozygous_count"], homoz)
    assert np.array_equal(st["recomb_count"], rec) and st["quadrivalent_hist"].tolist() == [0, 3000]
