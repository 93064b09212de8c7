"""Pins of the CPU oracle (oracle/) against everything the reference offers as a known answer.

The reference (PedigreeSim 2.2) ships no tests, fixtures or expected outputs and
cannot run here (no JVM). What pins the restatement:
  * the JDK contract of java.util.Random (first outputs for seeds 0 and 42, publicly documented);
  * the Random123 known-answer vectors of Philox4x32-10;
  * the closed forms coded in the reference: Haldane / Kosambi (Tools.java:116,140);
  * manual Table 1 (natural pairing: #meioses with k quadrivalents) and Table 2 (binomial);
  * structural identities of the model.
Statistical checks use a chi-square / z bound of p > 1e-6 so that they never flake.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest
from scipy import stats

import _cases
import _oracle
from _oracle import RNG_JAVA, RNG_PHILOX, Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
P_FLOOR = 1e-6


def test_java_util_random_known_answers(oracle_lib):
    L = oracle_lib
    assert L.orc_java_next_int32(42, 0) == -1170105035
    assert L.orc_java_next_int32(0, 0) == -1155484576
    assert L.orc_java_next_double(42, 0) == 0.7275636800328681
    assert L.orc_java_next_double(0, 0) == 0.730967787376657
    assert [L.orc_java_next_int(42, 10, i) for i in range(10)] == [0, 3, 8, 4, 0, 5, 5, 8, 9, 3]


def test_philox4x32_10_known_answers(oracle_lib):
    def blk(c, k):
        ca, ka, o = (C.c_uint32 * 4)(*c), (C.c_uint32 * 2)(*k), (C.c_uint32 * 4)()
        oracle_lib.orc_philox_block(ca, ka, o)
        return list(o)
    assert blk([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert blk([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert blk([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_java_double_to_string(oracle_lib):
    def fmt(x):
        b = C.create_string_buffer(64)
        oracle_lib.orc_java_double_to_string(x, b, 64)
        return b.value.decode()
    assert fmt(25.0) == "25.0"
    assert fmt(0.0001) == "1.0E-4"
    assert fmt(1234567.0) == "1234567.0"
    assert fmt(12345678.0) == "1.2345678E7"
    assert fmt(0.001) == "0.001"
    assert fmt(37.5123) == "37.5123"
    assert fmt(100.0) == "100.0"
    x = 0.1 + 0.2
    assert float(fmt(x)) == x and fmt(x) == "0.30000000000000004"


def test_calc_recomb_dist_solves_its_equation(oracle_lib):
    # Tools.calcRecombDist (:338-386): m == 0.5 / (1 - exp(-L/m))
    for L in (0.7, 1.0, 1.5, 3.0):
        m = oracle_lib.orc_calc_recomb_dist(L)
        assert abs(m - 0.5 / (1 - np.exp(-L / m))) < 1e-4
    assert oracle_lib.orc_calc_recomb_dist(6.0) == 0.5
    assert np.isnan(oracle_lib.orc_calc_recomb_dist(0.4))


def test_gamma_sampler_moments(oracle_lib):
    # Tools.ranGamma (Cheng): mean k*theta, variance k*theta^2
    k, theta = 2.63, 0.5 / 2.63
    x = np.array([oracle_lib.orc_ran_gamma(1234, k, theta, i) for i in range(0, 4000, 1)])
    # successive skips re-create the stream, so x[i] is the (i+1)-th draw
    assert abs(x.mean() - k * theta) < 5 * np.sqrt(k * theta * theta / len(x))
    assert abs(x.var() - k * theta * theta) < 0.15 * k * theta * theta


def _selfed_founder(ploidy, n_children, lengths, centro, pref, fq, rng, seed, **model):
    o = Oracle(ploidy, rng=rng, seed=seed, **model)
    o.set_chromosomes(lengths, centro, pref, fq)
    p = np.array([-1] + [0] * n_children, np.int32)
    o.set_pedigree(p, p)
    o.simulate()
    return o


@pytest.mark.parametrize("rng", [RNG_JAVA, RNG_PHILOX])
def test_manual_table1_natural_pairing_quadrivalent_counts(rng):
    T = json.load(open(os.path.join(GOLDEN, "manual_table1.json")))
    for ploidy, rows in T["counts_by_ploidy"].items():
        ploidy = int(ploidy)
        for j, pref in enumerate(T["pref_pairing"]):
            manual = np.array([r[j] for r in rows], float)
            o = _selfed_founder(ploidy, 5000, [1.0], [0.5], [pref], [0.0], rng, seed=100 + j)
            ours = o.stats()[:len(manual)].astype(float)
            assert ours.sum() == 10000
            if pref == 1.0:
                assert ours[0] == 10000 and manual[0] == 10000
                continue
            # two-sample chi-square: the manual's counts are themselves one sample of 10000
            keep = (ours + manual) > 0
            chi2, p, _, _ = stats.chi2_contingency(np.vstack([ours[keep], manual[keep]]))
            assert p > P_FLOOR, (ploidy, pref, ours, manual, p)


@pytest.mark.parametrize("rng", [RNG_JAVA, RNG_PHILOX])
def test_manual_table2_fixed_quadrivalent_fraction(rng):
    # NATURALPAIRING = 0, fraction 0.5: Binomial(ploidy // 4, 0.5)
    for ploidy in (4, 6, 8, 10, 12):
        n = ploidy // 4
        o = _selfed_founder(ploidy, 5000, [1.0], [0.5], [0.0], [0.5], rng, seed=7, natural_pairing=False)
        ours = o.stats()[:n + 1].astype(float)
        expect = stats.binom.pmf(np.arange(n + 1), n, 0.5) * ours.sum()
        chi2, p = stats.chisquare(ours, expect)
        assert p > P_FLOOR, (ploidy, ours, expect)


def test_tetraploid_pref0_two_thirds_quadrivalents():
    # analytic: with no preference a tetraploid forms a quadrivalent in 2/3 of the meioses
    o = _selfed_founder(4, 20000, [1.0], [0.5], [0.0], [0.0], RNG_JAVA, seed=1)
    s = o.stats()
    z = (s[1] - 40000 * 2 / 3) / np.sqrt(40000 * 2 / 9)
    assert abs(z) < 5


@pytest.mark.parametrize("rng", [RNG_JAVA, RNG_PHILOX])
@pytest.mark.parametrize("kosambi", [False, True])
def test_bivalent_recombination_fractions_follow_map_function(rng, kosambi):
    n = 30000
    o = _selfed_founder(2, n, [1.0], [0.3], None, None, rng, seed=99, kosambi=kosambi)
    pos = np.array([0, 0.05, 0.1, 0.2, 0.4, 0.7, 1.0])
    o.set_map([0, len(pos)], pos)
    F = o.emit(_oracle.EMIT_FOUNDER, ind_first=1).reshape(len(pos), n, 2)
    for a in range(len(pos)):
        for b in range(a + 1, len(pos)):
            d = pos[b] - pos[a]
            expect = 0.5 * np.tanh(2 * d) if kosambi else 0.5 * (1 - np.exp(-2 * d))  # Tools.java:116,140
            r = (F[a] != F[b]).mean()
            se = np.sqrt(expect * (1 - expect) / (2 * n))
            # Kosambi is only approximated by the gamma(2.63) renewal process (Tools.java:317-331)
            tol = 5 * se + (0.006 if kosambi else 0.0)
            assert abs(r - expect) < tol, (kosambi, d, r, expect)


@pytest.mark.parametrize("ploidy", [2, 4, 6])
def test_each_parental_homolog_is_transmitted_with_probability_half(ploidy):
    n = 6000
    o = _selfed_founder(ploidy, n, [1.0], [0.5], [0.2], [0.0], RNG_PHILOX, seed=5)
    o.set_map([0, 3], [0.1, 0.5, 0.9])
    F = o.emit(_oracle.EMIT_FOUNDER, ind_first=1).reshape(3, n, ploidy)
    for l in range(3):
        for f in range(ploidy):
            cnt = (F[l] == f).sum()  # over 2n gametes x ploidy/2 slots
            expect = 2 * n * 0.5
            assert abs(cnt - expect) < 6 * np.sqrt(2 * n * 0.25 * 1.5), (ploidy, l, f, cnt)


def test_haldane_mean_breakpoints_per_chromatid():
    # Poisson chiasmata at rate 2/Morgan on the bundle -> each chromatid sees rate 1/Morgan
    n = 20000
    L = 1.3
    o = _selfed_founder(2, n, [L], [0.5], None, None, RNG_JAVA, seed=3)
    so, _, _ = o.get_haplostructs(first=1)
    mean_bp = (np.diff(so) - 1).mean()
    assert abs(mean_bp - L) < 5 * np.sqrt(L / (2 * n))


def test_founder_columns_are_their_own_alleles():
    p0, p1 = _cases.std_pedigree("F1", 10)
    o = Oracle(4, rng=RNG_JAVA, seed=1)
    o.set_chromosomes([1.0], [0.5], [0.0], [0.0])
    o.set_pedigree(p0, p1)
    o.simulate()
    o.set_map([0, 4], [0, 0.3, 0.6, 1.0])
    F = o.emit(_oracle.EMIT_FOUNDER)
    assert np.all(F[:, :8] == np.arange(8))
    # offspring: first two homologs from P1 (alleles 0-3), last two from P2 (4-7): Gamete.java:95-116
    kids = F[:, 8:].reshape(4, 10, 4)
    assert kids[:, :, :2].max() <= 3 and kids[:, :, 2:].min() >= 4


@pytest.mark.parametrize("rng", [RNG_JAVA, RNG_PHILOX])
@pytest.mark.parametrize("kind", [1, 2])
def test_unreduced_gametes_heterozygosity_follows_tetrad_analysis(rng, kind):
    """2NGAMETES (Individual.java:341-383) on a diploid bivalent, Haldane. An SDR gamete holds the two
    sister chromatids of one centromere, an FDR gamete one chromatid of each centromere. With y(d) the
    frequency of second-division segregation at d Morgan from the centromere (Mather: per chiasma a
    first-division pattern always becomes second-division, a second-division one stays so with
    probability 1/2; chiasmata Poisson with mean 2d), y = 2/3 (1 - exp(-3d)), the gamete is
    heterozygous with probability y (SDR) or 1 - y/2 (FDR)."""
    n = 8000
    ce = 0.2
    o = _selfed_founder(2, n, [1.0], [ce], None, None, rng, seed=77 + kind, unreduced_gametes=kind)
    pos = np.array([ce, 0.3, 0.5, 1.0])
    o.set_map([0, len(pos)], pos)
    F = o.emit(_oracle.EMIT_FOUNDER, ind_first=1).reshape(len(pos), n, 2)
    het = (F[:, :, 0] != F[:, :, 1]).mean(axis=1)
    assert het[0] == (1.0 if kind == 1 else 0.0)        # at the centromere: FDR always, SDR never
    for k in range(1, len(pos)):
        y = 2.0 / 3.0 * (1 - np.exp(-3 * (pos[k] - ce)))
        expect = 1 - y / 2 if kind == 1 else y
        assert abs(het[k] - expect) < 5 * np.sqrt(expect * (1 - expect) / n), (kind, pos[k], het[k], expect)
    if kind == 2:   # SDR: either centromere with probability 1/2
        c0 = (F[0, :, 0] == 0).mean()
        assert abs(c0 - 0.5) < 5 * np.sqrt(0.25 / n)
    quad, seq = o.get_meiotic_configs(first=1)
    assert (quad[:, :, 1] == -1).all() and (seq[:, :, 1] == -1).all()


def test_unreduced_tetraploid_gametes_carry_every_centromere_once_under_fdr():
    """FDR 'each gamete must get one copy of the centromere of each parental homologue'
    (Individual.java:357-358) - true for bivalent pairing; SDR gametes carry half of them twice."""
    n = 3000
    ce = 0.45
    for kind in (1, 2):
        o = _selfed_founder(4, n, [1.0], [ce], [0.5], [0.0], RNG_PHILOX, seed=21, natural_pairing=False, unreduced_gametes=kind)
        o.set_map([0, 1], [ce])
        F = np.sort(o.emit(_oracle.EMIT_FOUNDER, ind_first=1).reshape(n, 4), axis=1)
        if kind == 1:
            assert (F == np.arange(4)).all()
        else:
            assert (F[:, 0] == F[:, 1]).all() and (F[:, 2] == F[:, 3]).all() and (F[:, 1] != F[:, 2]).all()


# ---------------------------------------------------------------- the jar itself (recorded where a JVM exists)
JAR_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jar")
JAR_CASES = sorted(os.listdir(JAR_GOLDEN)) if os.path.isdir(JAR_GOLDEN) else []


@pytest.mark.skipif(not JAR_CASES, reason="no jar-recorded goldens: run tests/golden/make_jar_goldens.sh on a machine with a JVM")
@pytest.mark.parametrize("case", JAR_CASES or ["none"])
def test_oracle_java_mode_reproduces_the_jar(tmp_path, case):
    """Draw-for-draw equality of the restatement with PedigreeSim.jar: same SEED, same files
    (.ped, .hsa, _meioticconfigs.dat and the four .dat tables byte for byte; .hsb to 1e-12, Math.log ulps)."""
    import shutil
    g = os.path.join(JAR_GOLDEN, case)
    for n in ("case.par", "case.chrom", "case.map", "case.gen"):
        shutil.copy(os.path.join(g, n), tmp_path / n)
    _oracle.run_parameter_file("case.par", cwd=str(tmp_path))
    for n in (".ped", ".hsa", "_meioticconfigs.dat", "_founderalleles.dat", "_genotypes.dat", "_alleledose.dat", "_allAlleledose.dat"):
        if os.path.exists(os.path.join(g, "out" + n)):
            assert open(tmp_path / ("out" + n), "rb").read() == open(os.path.join(g, "out" + n), "rb").read(), n
    a, b = open(tmp_path / "out.hsb").read().split(), open(os.path.join(g, "out.hsb")).read().split()
    assert len(a) == len(b)
    for x, y in zip(a, b):
        if x != y:
            assert abs(float(x) - float(y)) <= 1e-12 * max(1.0, abs(float(y))), (x, y)
