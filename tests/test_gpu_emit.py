"""Replay parity of the emission kernels (K0 + K2) through the C ABI: given HaploStructs, the
founder-allele / genotype / dose tables must equal the CPU oracle's bit for bit
(reference: Main.writeGenotypesFile / writeAlleledoseFile / writeAllAlleledoseFile,
Main.java:903-1088, over HaploStruct.getFounderAt, HaploStruct.java:107-142)."""
import numpy as np
import pytest

import _cases
import _oracle

pytestmark = pytest.mark.gpu


def _both(ploidy, lengths, n_ind, n_alleles, loci, seed, codes=0, max_segments=6, grid_breaks=True):
    rng = np.random.default_rng(seed)
    off, pos = _cases.uniform_map(lengths, loci)
    grid = pos[off[0]:off[1]] if grid_breaks and loci[0] > 2 else None
    seg_off, founder, segpos = _cases.random_haplostructs(rng, n_ind, lengths, ploidy, n_alleles, max_segments, grid)
    # only emission is exercised: every individual gets a preset haplostruct; enough founders to
    # own n_alleles founder alleles, the rest nominally their offspring
    nf = min(n_ind, -(-n_alleles // ploidy))
    p = np.concatenate([-np.ones(nf, np.int32), np.zeros(n_ind - nf, np.int32)])
    engines = []
    for kind in ("oracle-philox", "gpu"):
        e = _cases.make_engine(kind, ploidy)
        e.set_chromosomes(lengths, [x * 0.4 for x in lengths])
        e.set_pedigree(p, p)
        e.set_haplostructs(0, seg_off, founder, segpos)
        e.set_map(off, pos)
        if codes:
            e.set_allele_codes(_cases.hash_codes(int(off[-1]), n_alleles, codes, salt=seed))
        engines.append(e)
    return engines


@pytest.mark.parametrize("ploidy,n_ind,loci", [
    (2, 37, [50]), (2, 1200, [300, 7, 129]), (4, 700, [257, 64]), (6, 300, [130, 200]), (8, 90, [77]),
    (10, 33, [40, 41]), (12, 20, [35]), (16, 9, [19]),
])
def test_founder_and_dose_tables_match_oracle(ploidy, n_ind, loci):
    lengths = [1.0 + 0.25 * k for k in range(len(loci))]
    n_alleles = min(255, ploidy * 5)
    o, g = _both(ploidy, lengths, n_ind, n_alleles, loci, seed=ploidy * 1000 + n_ind)
    for which in (0, 3, n_alleles - 1):
        ref = _cases.oracle_emit_all(o, which=which)
        got = g.emit(founder=True, dose=True, dose_which=which)
        assert np.array_equal(got["founder"], ref["founder"])
        assert np.array_equal(got["dose"], ref["dose"])


@pytest.mark.parametrize("ploidy,codes", [(2, 2), (4, 2), (4, 5), (6, 3), (8, 2), (12, 4)])
def test_genotype_codes_and_max_allele_dose_match_oracle(ploidy, codes):
    lengths = [1.2, 0.8]
    loci = [211, 97]
    n_alleles = ploidy * 2
    o, g = _both(ploidy, lengths, 260, n_alleles, loci, seed=77 + ploidy, codes=codes)
    for which in (_oracle.MAX_ALLELE, _oracle.MIN_ALLELE, 1):
        got = g.emit(founder=True, allele=True, dose=True, dose_which=which)
        assert np.array_equal(got["founder"], o.emit(_oracle.EMIT_FOUNDER))
        assert np.array_equal(got["allele"], o.emit(_oracle.EMIT_ALLELE))
        assert np.array_equal(got["dose"], o.emit(_oracle.EMIT_DOSE, which=which))


def test_sub_ranges_and_single_tables():
    o, g = _both(4, [1.0, 1.5], 333, 8, [150, 260], seed=5)
    full_f = o.emit(_oracle.EMIT_FOUNDER)
    full_d = o.emit(_oracle.EMIT_DOSE, which=2)
    for (i0, ni, l0, nl) in [(0, 333, 0, 410), (10, 100, 140, 30), (332, 1, 409, 1), (7, 129, 0, 1), (100, 233, 149, 2)]:
        got = g.emit(founder=True, ind_first=i0, ind_count=ni, loc_first=l0, loc_count=nl)
        assert np.array_equal(got["founder"], full_f[l0:l0 + nl, i0 * 4:(i0 + ni) * 4])
        got = g.emit(dose=True, dose_which=2, ind_first=i0, ind_count=ni, loc_first=l0, loc_count=nl)
        assert np.array_equal(got["dose"], full_d[l0:l0 + nl, i0:i0 + ni])


def test_locus_on_breakpoint_belongs_to_the_segment_starting_there():
    # HaploStruct.findSegment: a locus exactly on a breakpoint takes the founder of the new segment
    lengths = [1.0]
    off = np.array([0, 5], np.int64)
    pos = np.array([0.0, 0.25, 0.25, 0.5, 1.0])
    seg_off = np.array([0, 3, 4], np.int64)
    founder = np.array([0, 1, 0, 1], np.int32)
    segpos = np.array([0.0, 0.25, 0.5, 0.0])
    p = -np.ones(1, np.int32)
    res = []
    for kind in ("oracle-philox", "gpu"):
        e = _cases.make_engine(kind, 2)
        e.set_chromosomes(lengths, [0.5])
        e.set_pedigree(p, p)
        e.set_haplostructs(0, seg_off, founder, segpos)
        e.set_map(off, pos)
        res.append(e)
    ref = res[0].emit(_oracle.EMIT_FOUNDER)
    assert ref[:, 0].tolist() == [0, 1, 1, 0, 0]
    assert np.array_equal(res[1].emit(founder=True)["founder"], ref)


def test_unsorted_and_redundant_breakpoints_follow_findsegment():
    # hand-written .hsb files may hold equal-founder neighbours, zero-length and even unsorted
    # segments; the answer is defined by the tail scan of findSegment
    rng = np.random.default_rng(3)
    lengths = [1.0]
    off, pos = _cases.uniform_map(lengths, [64])
    n_ind, P = 50, 2
    seg_off, founder, segpos = [0], [], []
    for _ in range(n_ind * P):
        n = int(rng.integers(1, 7))
        bp = rng.random(n - 1)  # NOT sorted
        if n > 2 and rng.random() < 0.5:
            bp[1] = bp[0]  # zero-length segment
        segpos += [0.0] + bp.tolist()
        founder += rng.integers(0, 4, n).tolist()
        seg_off.append(seg_off[-1] + n)
    p = -np.ones(n_ind, np.int32)
    res = []
    for kind in ("oracle-philox", "gpu"):
        e = _cases.make_engine(kind, P)
        e.set_chromosomes(lengths, [0.5])
        e.set_pedigree(p, p)
        e.set_haplostructs(0, np.array(seg_off), np.array(founder), np.array(segpos))
        e.set_map(off, pos)
        res.append(e)
    assert np.array_equal(res[1].emit(founder=True)["founder"], res[0].emit(_oracle.EMIT_FOUNDER))


def test_all_allele_dose_rows():
    for codes in (0, 3):
        o, g = _both(4, [1.0], 120, 8, [90], seed=11, codes=codes)
        ro, out = g.emit_all_dose()
        n_loci = 90
        if codes == 0:
            assert ro.tolist() == [8 * k for k in range(n_loci + 1)]
            for a in range(8):
                assert np.array_equal(out[a::8], o.emit(_oracle.EMIT_DOSE, which=a))
        else:
            al = o.emit(_oracle.EMIT_ALLELE).reshape(n_loci, 120, 4)
            for l in range(n_loci):
                for k in range(int(ro[l + 1] - ro[l])):
                    assert np.array_equal(out[ro[l] + k], (al[l] == k).sum(axis=1))


def test_u16_founder_table_for_many_founders():
    ploidy, n_ind = 2, 40
    rng = np.random.default_rng(9)
    lengths = [1.0]
    off, pos = _cases.uniform_map(lengths, [33])
    seg_off, founder, segpos = _cases.random_haplostructs(rng, n_ind, lengths, ploidy, 1000, 4)
    # 600 founders -> 1200 founder alleles: needs the u16 table
    n_f = 600
    p0 = np.concatenate([-np.ones(n_f, np.int32), np.zeros(n_ind, np.int32)])
    p1 = np.concatenate([-np.ones(n_f, np.int32), np.ones(n_ind, np.int32)])
    res = []
    for kind in ("oracle-philox", "gpu"):
        e = _cases.make_engine(kind, ploidy)
        e.set_chromosomes(lengths, [0.5])
        e.set_pedigree(p0, p1)
        e.set_haplostructs(n_f, seg_off, founder, segpos)
        e.simulate()  # founders only
        e.set_map(off, pos)
        res.append(e)
    ref = res[0].emit(_oracle.EMIT_FOUNDER_U16)
    got = res[1].emit(founder=True, founder_bytes=2)["founder"]
    assert np.array_equal(got, ref)
    import pedigreesim_b200
    with pytest.raises(pedigreesim_b200.PsimError):
        res[1].emit(founder=True, founder_bytes=1)


@pytest.mark.parametrize("ploidy,n_ind,codes", [(2, 41000, 2), (2, 33000, 0), (4, 15011, 2), (6, 12000, 0), (4, 30001, 0)])
def test_rows_wider_than_one_shared_memory_segment(ploidy, n_ind, codes):
    """More than ~56 K cells per locus row: the row-image kernel cuts the row into segments, each
    with its own bulk stores (ragged last segment, dose tails that are not multiples of 16)."""
    lengths = [1.0, 0.6]
    loci = [70, 45]
    n_alleles = ploidy * 2
    o, g = _both(ploidy, lengths, n_ind, n_alleles, loci, seed=n_ind, codes=codes, max_segments=4)
    which = _oracle.MAX_ALLELE if codes else 1
    got = g.emit(founder=True, dose=True, dose_which=which)
    assert np.array_equal(got["founder"], o.emit(_oracle.EMIT_FOUNDER))
    assert np.array_equal(got["dose"], o.emit(_oracle.EMIT_DOSE, which=which))
    # a window of individuals and loci that starts inside a segment
    i0, ni, l0, nl = 1234, n_ind - 2000, 30, 60
    sub = g.emit(founder=True, dose=True, dose_which=which, ind_first=i0, ind_count=ni, loc_first=l0, loc_count=nl)
    assert np.array_equal(sub["founder"], got["founder"][l0:l0 + nl, i0 * ploidy:(i0 + ni) * ploidy])
    assert np.array_equal(sub["dose"], got["dose"][l0:l0 + nl, i0:i0 + ni])


def test_many_run_starts_per_row_fall_back_cleanly():
    """A sparse map (few loci, long chromosomes): nearly every cell changes from one locus to the
    next; the row-image path must hand over (or redo overflowing units) without a wrong byte."""
    o, g = _both(4, [6.0], 3000, 8, [24], seed=8, max_segments=40, grid_breaks=False)
    got = g.emit(founder=True, dose=True, dose_which=2)
    assert np.array_equal(got["founder"], o.emit(_oracle.EMIT_FOUNDER))
    assert np.array_equal(got["dose"], o.emit(_oracle.EMIT_DOSE, which=2))


def test_overflowing_units_are_redone_per_cell():
    """Force the row-image path (PSIM_PATH_ROWS) onto a map where the run-start lists overflow: those
    units go to the overflow list and are redone by the per-cell kernel; the tables must still be exact."""
    import pedigreesim_b200 as ps
    for ploidy, codes in ((4, 0), (4, 2), (2, 2)):
        o, g = _both(ploidy, [6.0, 1.0], 3000, ploidy * 2, [24, 40], seed=21 + ploidy, codes=codes, max_segments=40, grid_breaks=False)
        which = _oracle.MAX_ALLELE if codes else 3
        got = g.emit(founder=True, dose=True, dose_which=which, path=ps.PATH_ROWS)
        assert np.array_equal(got["founder"], o.emit(_oracle.EMIT_FOUNDER))
        assert np.array_equal(got["dose"], o.emit(_oracle.EMIT_DOSE, which=which))


def test_more_than_64_chromosomes_and_replicates():
    """Launch arguments hold 64 chromosome pieces: more chromosomes go out in several launches.
    Three replicate populations side by side widen the rows (columns are replicate-major)."""
    import pedigreesim_b200 as ps
    C, P, n = 70, 4, 300
    lengths = [0.5 + 0.01 * c for c in range(C)]
    loci = [3 + c % 5 for c in range(C)]
    off, pos = _cases.uniform_map(lengths, loci)
    p0, p1 = _cases.std_pedigree("F1", n)
    g = ps.PsimContext(P, seed=5, replicate_first=7, replicate_count=3)
    g.set_chromosomes(lengths, [0.4 * x for x in lengths], [0.2] * C, [0.3] * C)
    g.set_pedigree(p0, p1)
    g.simulate()
    g.set_map(off, pos)
    got = g.emit(founder=True, dose=True, dose_which=5)
    N = len(p0)
    assert got["founder"].shape == (sum(loci), 3 * N * P)
    for r in range(3):
        o = _oracle.Oracle(P, rng=_oracle.RNG_PHILOX, seed=5, replicate=7 + r)
        o.set_chromosomes(lengths, [0.4 * x for x in lengths], [0.2] * C, [0.3] * C)
        o.set_pedigree(p0, p1)
        o.simulate()
        o.set_map(off, pos)
        assert np.array_equal(got["founder"][:, r * N * P:(r + 1) * N * P], o.emit(_oracle.EMIT_FOUNDER))
        assert np.array_equal(got["dose"][:, r * N:(r + 1) * N], o.emit(_oracle.EMIT_DOSE, which=5))
    g.set_locus_names(["L%d" % l for l in range(sum(loci))])
    text = g.emit_text(founder=True)["founder"].split(b"\n")
    assert text[10].split(b"\t")[1:] == [str(int(v)).encode() for v in got["founder"][10]]


def _pack4(table):
    """Two 4-bit cells per byte, cell 2k in the low nibble; an odd last cell is padded with zero (PSIM_EMIT_PACKED4)."""
    t = np.asarray(table, dtype=np.uint8)
    if t.shape[1] % 2:
        t = np.concatenate([t, np.zeros((t.shape[0], 1), np.uint8)], axis=1)
    return (t[:, 0::2] & 15) | ((t[:, 1::2] & 15) << 4)


@pytest.mark.parametrize("ploidy,n_ind,loci,codes,which", [
    (4, 700, [257, 64], 2, _oracle.MAX_ALLELE), (4, 701, [257, 64], 2, 1), (2, 1233, [300, 7, 129], 2, _oracle.MIN_ALLELE),
    (4, 333, [150, 33], 0, 3), (2, 77, [90], 0, 0), (6, 301, [130, 20], 0, 5), (6, 90, [77], 3, _oracle.MAX_ALLELE), (8, 41, [40], 0, 9),
])
def test_packed_tables_are_the_oracle_tables_packed(ploidy, n_ind, loci, codes, which):
    import pedigreesim_b200 as ps
    lengths = [1.0 + 0.25 * k for k in range(len(loci))]
    n_alleles = min(16, ploidy * 2)
    o, g = _both(ploidy, lengths, n_ind, n_alleles, loci, seed=ploidy * 100 + n_ind, codes=codes)
    ref_f, ref_d = _pack4(o.emit(_oracle.EMIT_FOUNDER)), _pack4(o.emit(_oracle.EMIT_DOSE, which=which))
    for path in (ps.PATH_AUTO, ps.PATH_ROWS, ps.PATH_GENERIC):
        got = g.emit(founder=True, dose=True, dose_which=which, packed=True, path=path)
        assert np.array_equal(got["founder"], ref_f), path
        assert np.array_equal(got["dose"], ref_d), path
    got = g.emit(founder=True, packed=True)
    assert np.array_equal(got["founder"], ref_f)
    got = g.emit(dose=True, dose_which=which, packed=True)
    assert np.array_equal(got["dose"], ref_d)
    # a sub-range starts its own packing at its first individual
    i0, ni, l0, nl = 3, n_ind // 2 | 1, 5, loci[0] - 9
    got = g.emit(founder=True, dose=True, dose_which=which, packed=True, ind_first=i0, ind_count=ni, loc_first=l0, loc_count=nl)
    assert np.array_equal(got["founder"], _pack4(o.emit(_oracle.EMIT_FOUNDER)[l0:l0 + nl, i0 * ploidy:(i0 + ni) * ploidy]))
    assert np.array_equal(got["dose"], _pack4(o.emit(_oracle.EMIT_DOSE, which=which)[l0:l0 + nl, i0:i0 + ni]))


def test_packed_rows_wider_than_one_segment_and_overflowing_lists():
    import pedigreesim_b200 as ps
    # 60 000 individuals x 4: several row segments per table row
    o, g = _both(4, [1.0], 60000, 8, [21], seed=3, codes=2, max_segments=3)
    got = g.emit(founder=True, dose=True, dose_which=_oracle.MAX_ALLELE, packed=True)
    assert np.array_equal(got["founder"], _pack4(o.emit(_oracle.EMIT_FOUNDER)))
    assert np.array_equal(got["dose"], _pack4(o.emit(_oracle.EMIT_DOSE, which=_oracle.MAX_ALLELE)))
    # a sparse map forced onto the row-image kernels: the lists overflow, the units are redone per pair
    for ploidy, codes in ((4, 0), (4, 2), (2, 2)):
        o, g = _both(ploidy, [6.0, 1.0], 3001, ploidy * 2, [24, 40], seed=41 + ploidy, codes=codes, max_segments=40, grid_breaks=False)
        which = _oracle.MAX_ALLELE if codes else 3
        got = g.emit(founder=True, dose=True, dose_which=which, packed=True, path=ps.PATH_ROWS)
        assert np.array_equal(got["founder"], _pack4(o.emit(_oracle.EMIT_FOUNDER)))
        assert np.array_equal(got["dose"], _pack4(o.emit(_oracle.EMIT_DOSE, which=which)))


def test_packed_requests_that_cannot_be_served_are_refused():
    import pedigreesim_b200 as ps
    o, g = _both(4, [1.0], 10, 20, [9], seed=1)          # 20 founder alleles do not fit 4 bits
    with pytest.raises(ps.PsimError) as e:
        g.emit(founder=True, packed=True)
    assert e.value.code == ps.PSIM_EINVAL and "16 founder alleles" in str(e.value)


def test_founder_genotypes_with_more_than_256_founder_alleles():
    """75 tetraploid founders own 300 founder alleles: the founder table needs 16 bits, the allele codes
    are still bytes (at most 256 unique names per locus) indexed by the 16-bit founder allele."""
    o, g = _both(4, [1.0, 0.7], 150, 300, [60, 45], seed=12, codes=5)
    for which in (_oracle.MAX_ALLELE, _oracle.MIN_ALLELE, 280):
        got = g.emit(founder=True, allele=True, dose=True, dose_which=which, founder_bytes=2)
        assert np.array_equal(got["founder"], o.emit(_oracle.EMIT_FOUNDER_U16))
        assert np.array_equal(got["allele"], o.emit(_oracle.EMIT_ALLELE))
        assert np.array_equal(got["dose"], o.emit(_oracle.EMIT_DOSE, which=which))
    ro, ad = g.emit_all_dose()
    assert ad.sum(axis=0).min() >= 0 and int(ro[-1]) == ad.shape[0]


def test_preset_founder_alleles_must_exist_in_the_pedigree():
    import pedigreesim_b200 as ps
    g = ps.PsimContext(2)
    g.set_chromosomes([1.0], [0.5])
    g.set_pedigree([-1, -1, 0], [-1, -1, 1])            # 2 founders x 2 = founder alleles 0..3 (Main.java:1235)
    seg_off = np.arange(0, 5, dtype=np.int64)
    with pytest.raises(ps.PsimError) as e:
        g.set_haplostructs(0, seg_off, np.array([0, 1, 2, 4], np.int32), np.zeros(4))
    assert e.value.code == ps.PSIM_EINVAL and "invalid founder allele '4'" in str(e.value)
    g.set_haplostructs(0, seg_off, np.array([0, 1, 2, 3], np.int32), np.zeros(4))
    g.simulate(3)
    # a pedigree whose founders would be numbered past its own founder allele count
    h = ps.PsimContext(2)
    h.set_chromosomes([1.0], [0.5])
    h.set_pedigree([-1, -1, 0], [-1, -1, 1])
    with pytest.raises(ps.PsimError) as e:
        h.simulate(300)
    assert e.value.code == ps.PSIM_EINVAL and "invalid founder allele" in str(e.value)
