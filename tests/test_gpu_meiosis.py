"""Parity of the meiosis kernels (K1) through the C ABI against the CPU oracle running on the same
Philox streams: founder alleles per segment, segment counts and meiotic configurations must be
bit-exact; breakpoints (fp64, computed with CUDA's log/exp/pow instead of libm's) must agree to
1e-12 relative. Reference path: Individual.conception/doMeiosis/calcChromConfig
(Individual.java:255-349,601-904), Multivalent.java:79-577, Quadrivalent.java:108-510."""
import numpy as np
import pytest

import _cases

pytestmark = pytest.mark.gpu

POS_RTOL = 1e-12


def run_pair(ploidy, p0, p1, lengths, centro, pref=None, fq=None, seed=12345, replicate=0, **model):
    out = []
    for kind in ("oracle-philox", "gpu"):
        e = _cases.make_engine(kind, ploidy, seed=seed, replicate_first=replicate, **model)
        e.set_chromosomes(lengths, centro, pref, fq)
        e.set_pedigree(p0, p1)
        e.simulate()
        out.append(e)
    return out


def assert_same_population(o, g, replicate=None):
    so, fo, po = o.get_haplostructs()
    sg, fg, pg = g.get_haplostructs(replicate=replicate) if replicate is not None else g.get_haplostructs()
    assert np.array_equal(so, sg), "segment counts differ"
    assert np.array_equal(fo, fg), "founder alleles differ"
    assert np.allclose(po, pg, rtol=POS_RTOL, atol=1e-15), "breakpoints differ beyond fp tolerance"
    qo, sqo = o.get_meiotic_configs()
    qg, sqg = g.get_meiotic_configs(replicate=replicate) if replicate is not None else g.get_meiotic_configs()
    assert np.array_equal(qo, qg), "quadrivalent counts differ"
    assert np.array_equal(sqo, sqg), "chromseq differs"


@pytest.mark.parametrize("kosambi", [False, True])
@pytest.mark.parametrize("poptype", ["F1", "F2", "BC", "S1"])
def test_diploid_standard_populations(poptype, kosambi):
    p0, p1 = _cases.std_pedigree(poptype, 200)
    o, g = run_pair(2, p0, p1, [1.0, 0.6, 2.5], [0.5, 0.1, 2.5], kosambi=kosambi)
    assert_same_population(o, g)


@pytest.mark.parametrize("ploidy", [4, 6, 8, 10, 12])
@pytest.mark.parametrize("natural", [True, False])
def test_polyploid_pairing_and_quadrivalents(ploidy, natural):
    p0, p1 = _cases.std_pedigree("F2", 150)  # F2: parents carry recombinant homologs
    C = 4
    o, g = run_pair(ploidy, p0, p1, [1.0, 1.3, 0.8, 2.0], [0.4, 0.65, 0.0, 1.0],
                    pref=[0.0, 0.3, 0.7, 1.0], fq=[0.5, 1.0, 0.25, 0.0], natural_pairing=natural, seed=99 + ploidy)
    assert_same_population(o, g)
    quad, _ = g.get_meiotic_configs()
    if ploidy >= 4:
        assert quad[3:, :C - 1].max() >= 1  # quadrivalents do occur


@pytest.mark.parametrize("kosambi", [False, True])
@pytest.mark.parametrize("paral,paired", [(0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (0.35, 0.6), (1.0, 1.0)])
def test_tetraploid_model_switches(kosambi, paral, paired):
    p0, p1 = _cases.std_pedigree("F2", 120)
    o, g = run_pair(4, p0, p1, [1.0, 1.6], [0.3, 0.8], pref=[0.0, 0.5], fq=[1.0, 0.5], kosambi=kosambi,
                    natural_pairing=False, parallel_multivalents=paral, paired_centromeres=paired, seed=4242)
    assert_same_population(o, g)


@pytest.mark.parametrize("ploidy", [2, 4])
@pytest.mark.parametrize("kosambi", [False, True])
def test_no_empty_bivalents(ploidy, kosambi):
    p0, p1 = _cases.std_pedigree("F2", 80)
    o, g = run_pair(ploidy, p0, p1, [0.8, 1.5], [0.4, 0.7], pref=[0.2, 0.2], fq=[0.5, 0.5], kosambi=kosambi,
                    allow_no_recomb=False, seed=31)
    assert_same_population(o, g)


def test_random_deep_pedigree():
    rng = np.random.default_rng(2)
    p0, p1 = _cases.random_pedigree(rng, 6, 400)
    o, g = run_pair(4, p0, p1, [1.0, 1.2], [0.5, 0.2], pref=[0.25, 0.0], fq=[0.0, 0.0], seed=8)
    assert_same_population(o, g)


def test_ril_selfing_chain():
    p0, p1 = _cases.ril_pedigree(64, 8)
    o, g = run_pair(2, p0, p1, [1.5] * 3, [0.75] * 3, seed=5)
    assert_same_population(o, g)
    so, _, _ = g.get_haplostructs()
    assert np.diff(so).max() > 4  # F8 homologs carry many unmerged segments


def test_replicates_use_their_own_streams():
    p0, p1 = _cases.std_pedigree("F1", 60)
    import pedigreesim_b200
    g = pedigreesim_b200.PsimContext(4, seed=77, replicate_first=5, replicate_count=3)
    g.set_chromosomes([1.0, 1.0], [0.5, 0.5], [0.25, 0.25], [0.0, 0.0])
    g.set_pedigree(p0, p1)
    g.simulate()
    for r in (5, 6, 7):
        o = _cases.make_engine("oracle-philox", 4, seed=77, replicate_first=r)
        o.set_chromosomes([1.0, 1.0], [0.5, 0.5], [0.25, 0.25], [0.0, 0.0])
        o.set_pedigree(p0, p1)
        o.simulate()
        assert_same_population(o, g, replicate=r)
    a = g.get_haplostructs(replicate=5)
    b = g.get_haplostructs(replicate=6)
    assert not (len(a[1]) == len(b[1]) and np.array_equal(a[1], b[1]))


def test_extend_population_mode3():
    # Main.java:343-373: individuals with preset haplostructs are kept, the rest simulated, new
    # founders continue the founder-allele numbering after maxfounderallele
    p0, p1 = _cases.std_pedigree("F1", 30)
    o1, g1 = run_pair(2, p0, p1, [1.0], [0.5], seed=3)
    so, fo, po = o1.get_haplostructs()
    # second cycle: a new founder (index 32) crossed with F1 individuals
    p0b = np.concatenate([p0, [-1], np.arange(2, 12, dtype=np.int32)])
    p1b = np.concatenate([p1, [-1], np.full(10, 32, np.int32)])
    out = []
    for kind in ("oracle-philox", "gpu"):
        e = _cases.make_engine(kind, 2, seed=3)
        e.set_chromosomes([1.0], [0.5])
        e.set_pedigree(p0b, p1b)
        e.set_haplostructs(0, so, fo, po)
        e.simulate(maxfounderallele=int(fo.max()))
        out.append(e)
    assert_same_population(*out)
    _, f2, _ = out[1].get_haplostructs(first=32, count=1)
    assert f2.tolist() == [4, 5]


def test_emission_after_simulation_matches_oracle():
    import _oracle
    p0, p1 = _cases.std_pedigree("F1", 300)
    o, g = run_pair(4, p0, p1, [1.0, 0.9], [0.4, 0.3], pref=[0.3, 0.3], fq=[0.5, 0.5], natural_pairing=False, seed=12)
    off, pos = _cases.uniform_map([1.0, 0.9], [400, 123])
    codes = _cases.hash_codes(int(off[-1]), 8, 2)
    for e in (o, g):
        e.set_map(off, pos)
        e.set_allele_codes(codes)
    got = g.emit(founder=True, allele=True, dose=True, dose_which=_oracle.MAX_ALLELE)
    # breakpoints may differ in the last ulp between libm and CUDA; no locus sits that close
    assert np.array_equal(got["founder"], o.emit(_oracle.EMIT_FOUNDER))
    assert np.array_equal(got["allele"], o.emit(_oracle.EMIT_ALLELE))
    assert np.array_equal(got["dose"], o.emit(_oracle.EMIT_DOSE, which=_oracle.MAX_ALLELE))


def test_sharded_generations_on_one_gpu_match_oracle():
    """The large-pedigree path without NCCL: two contexts on one GPU play the two ranks; each
    simulates its block of every generation (psim_simulate_subset) and receives the other's block
    through the device export / import calls the NCCL all-gather uses. Philox streams are keyed by
    individual, so both must equal the unsharded oracle."""
    import torch
    from pedigreesim_b200 import sharded
    p0, p1 = _cases.ril_pedigree(21, 4)
    lengths, centro = [1.0, 1.4], [0.5, 0.2]
    ranks = []
    for _ in range(2):
        g = _cases.make_engine("gpu", 2, seed=77)
        g.set_chromosomes(lengths, centro)
        g.set_pedigree(p0, p1)
        ranks.append(sharded.PsimShard(g))
    for sh in ranks:
        sh.simulate_subset(np.zeros(0, np.int32))      # founders everywhere
    for inds in sharded.levelize(p0, p1):
        blocks = sharded.block_partition(inds, 2)
        for r, sh in enumerate(ranks):
            if len(blocks[r]):
                sh.simulate_subset(blocks[r])
        for r, sh in enumerate(ranks):
            for first, count in sharded.contiguous_runs(blocks[r]):
                off, fo, po = sh.export_range(first, count)
                ranks[1 - r].import_range(first, count, off, fo, po)
        torch.cuda.synchronize()
    o = _cases.make_engine("oracle-philox", 2, seed=77)
    o.set_chromosomes(lengths, centro)
    o.set_pedigree(p0, p1)
    o.simulate()
    so, fo, po = o.get_haplostructs()
    for sh in ranks:
        sg, fg, pg = sh.ctx.get_haplostructs()
        assert np.array_equal(so, sg) and np.array_equal(fo, fg)
        assert np.allclose(po, pg, rtol=POS_RTOL, atol=1e-15)


@pytest.mark.parametrize("kind", [1, 2])   # 2NGAMETES = FDR, SDR (Individual.java:341-383)
@pytest.mark.parametrize("ploidy,natural", [(2, True), (4, False), (4, True), (6, True)])
def test_unreduced_gametes_fdr_and_sdr(ploidy, natural, kind):
    p0, p1 = _cases.std_pedigree("F2", 150)   # the 2n "individuals" of one generation are the parents of the next
    o, g = run_pair(ploidy, p0, p1, [1.0, 1.4, 0.7], [0.4, 0.7, 0.0], pref=[0.0, 0.4, 1.0], fq=[0.5, 1.0, 0.0],
                    natural_pairing=natural, unreduced_gametes=kind, seed=500 + kind)
    assert_same_population(o, g)
    quad, seq = g.get_meiotic_configs()
    assert (quad[3:, :, 1] == -1).all() and (seq[3:, :, 1] == -1).all()   # parent 1 took no part
    if ploidy == 4 and not natural:
        assert quad[3:, :2, 0].max() >= 1


def test_unreduced_gametes_kosambi_replicates():
    p0, p1 = _cases.std_pedigree("S1", 100)
    import pedigreesim_b200
    g = pedigreesim_b200.PsimContext(4, kosambi=True, natural_pairing=False, seed=9, replicate_first=2, replicate_count=2,
                                     unreduced_gametes=1, paired_centromeres=0.5)
    g.set_chromosomes([1.2], [0.5], [0.3], [0.6])
    g.set_pedigree(p0, p1)
    g.simulate()
    for r in (2, 3):
        o = _cases.make_engine("oracle-philox", 4, kosambi=True, natural_pairing=False, seed=9, replicate_first=r,
                               unreduced_gametes=1, paired_centromeres=0.5)
        o.set_chromosomes([1.2], [0.5], [0.3], [0.6])
        o.set_pedigree(p0, p1)
        o.simulate()
        assert_same_population(o, g, replicate=r)


def test_monte_carlo_loop_reuses_the_generation_lists():
    """psim_reset + psim_simulate again and again (the benchmark's loop): the cached founder and
    generation lists must give what a new context gives, also after calls that overwrite the device
    copies (a subset simulation, preset haplostructs) and after a pedigree of another shape."""
    import pedigreesim_b200
    p0, p1 = _cases.ril_pedigree(40, 4)
    lengths, centro = [1.0, 0.7], [0.5, 0.2]

    def oracle_run(pa, pb, seed, rep, preset=None):
        o = _cases.make_engine("oracle-philox", 2, seed=seed, replicate_first=rep)
        o.set_chromosomes(lengths, centro)
        o.set_pedigree(pa, pb)
        if preset is not None:
            o.set_haplostructs(0, *preset)
        o.simulate(maxfounderallele=3 if preset is not None else -1)
        return o

    g = pedigreesim_b200.PsimContext(2, seed=11)
    g.set_chromosomes(lengths, centro)
    g.set_pedigree(p0, p1)
    g.simulate()
    assert_same_population(oracle_run(p0, p1, 11, 0), g)
    for seed, rep in ((11, 0), (12, 3), (11, 7)):          # cached lists, other streams
        g.reset(seed, rep)
        g.simulate()
        assert_same_population(oracle_run(p0, p1, seed, rep), g, replicate=rep)
    # a sharded-style call sequence overwrites the device copies of the lists
    g.reset(5, 1)
    lev1 = [2]
    g.simulate_subset(lev1)
    g.simulate()
    assert_same_population(oracle_run(p0, p1, 5, 1), g, replicate=1)
    g.reset(6, 2)
    g.simulate()
    assert_same_population(oracle_run(p0, p1, 6, 2), g, replicate=2)
    # preset founders (replay / extend), then the empty state again
    so, fo, po = g.get_haplostructs(0, 2, replicate=2)
    g.reset(8, 0)
    g.set_haplostructs(0, so, fo, po)
    g.simulate(maxfounderallele=3)
    assert_same_population(oracle_run(p0, p1, 8, 0, preset=(so, fo, po)), g, replicate=0)
    g.reset(9, 0)
    g.simulate()
    assert_same_population(oracle_run(p0, p1, 9, 0), g, replicate=0)
    # another pedigree on the same context
    q0, q1 = _cases.std_pedigree("F2", 90)
    g.set_pedigree(q0, q1)
    g.simulate()
    assert_same_population(oracle_run(q0, q1, 9, 0), g, replicate=0)
    g.reset(10, 4)
    g.simulate()
    assert_same_population(oracle_run(q0, q1, 10, 4), g, replicate=4)


def test_subset_simulation_edge_cases():
    """psim_simulate_subset: unordered lists with duplicates, several generations in one call, a parent
    that is neither simulated nor listed (error, and the context stays usable), indices out of range."""
    import pedigreesim_b200
    p0, p1 = _cases.ril_pedigree(30, 4)          # 2 founders, F1, then 3 generations of 30 lines
    lengths, centro = [1.0, 0.5], [0.3, 0.25]
    o = _cases.make_engine("oracle-philox", 2, seed=3)
    o.set_chromosomes(lengths, centro)
    o.set_pedigree(p0, p1)
    o.simulate()
    g = pedigreesim_b200.PsimContext(2, seed=3)
    g.set_chromosomes(lengths, centro)
    g.set_pedigree(p0, p1)
    n = len(p0)
    with pytest.raises(pedigreesim_b200.PsimError, match="has no haplostruct yet"):
        g.simulate_subset([n - 1])               # its parent (previous generation) is not there
    with pytest.raises(pedigreesim_b200.PsimError, match="out of bounds"):
        g.simulate_subset([2, n])
    # F1 and two generations of the even lines in one call, unordered, with duplicates
    gen = lambda k: np.arange(3 + 30 * k, 3 + 30 * (k + 1))
    even = np.concatenate([gen(0)[::2], gen(1)[::2]])
    lst = np.concatenate([even[::-1], [2, 2], even[:5]]).astype(np.int32)
    g.simulate_subset(lst)
    with pytest.raises(pedigreesim_b200.PsimError, match="has no haplostruct yet"):
        g.simulate_subset(gen(1)[1::2])          # odd lines of generation 2: generation 1 odd lines are missing
    g.simulate_subset(np.concatenate([gen(1)[1::2], gen(0)[1::2]]))
    g.simulate()                                 # the rest
    assert_same_population(o, g)


def test_simulate_with_nothing_left_to_simulate_keeps_the_population():
    """Replay (HAPLOSTRUCT= covers everybody, Main.java:317-330): psim_simulate finds no individual without a haplostruct.
    The recorded segments must survive the call (the slab cursor on the device is only valid behind a batch)."""
    rng = np.random.default_rng(4)
    p0, p1 = _cases.std_pedigree("F1", 40)
    lengths = [1.0, 0.7]
    rec = _cases.random_haplostructs(rng, len(p0), lengths, 4, 8)
    g = _cases.make_engine("gpu", 4, seed=5)
    g.set_chromosomes(lengths, [0.5, 0.3], [0.3, 0.3], [0.5, 0.5])
    g.set_pedigree(p0, p1)
    g.set_haplostructs(0, *rec)
    g.simulate(7)
    g.simulate_subset(np.zeros(0, np.int32), 7)
    for a, b in zip(rec, g.get_haplostructs()):
        assert np.array_equal(a, b)
    g.close()
