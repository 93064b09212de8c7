"""psim_emit_text: the table files' lines formatted on the device must be the text the reference's
writers produce from the same cells (Main.writeGenotypesFile / writeAlleledoseFile,
Main.java:903-1022): `name { TAB token } NEWLINE` per locus."""
import numpy as np
import pytest

import _cases

pytestmark = pytest.mark.gpu


def expected_lines(names, table, token):
    return b"".join(nm.encode() + b"".join(b"\t" + token(l, v) for v in row) + b"\n" for l, (nm, row) in enumerate(zip(names, table)))


def make_ctx(ploidy, n_ind, loci, n_alleles, codes, seed, lengths=(1.0, 0.7)):
    import pedigreesim_b200 as ps
    rng = np.random.default_rng(seed)
    lengths = list(lengths)[:len(loci)]
    off, pos = _cases.uniform_map(lengths, loci)
    seg_off, founder, segpos = _cases.random_haplostructs(rng, n_ind, lengths, ploidy, n_alleles, 5)
    nf = min(n_ind, -(-n_alleles // ploidy))
    p = np.concatenate([-np.ones(nf, np.int32), np.zeros(n_ind - nf, np.int32)])
    g = ps.PsimContext(ploidy)
    g.set_chromosomes(lengths, [x * 0.4 for x in lengths])
    g.set_pedigree(p, p)
    g.set_haplostructs(0, seg_off, founder, segpos)
    g.set_map(off, pos)
    L = int(off[-1])
    names = ["mk%d_%s" % (l, "x" * (l % 7)) for l in range(L)]
    g.set_locus_names(names)
    code = None
    allele_names = None
    if codes:
        code = _cases.hash_codes(L, n_alleles, codes, salt=seed)
        g.set_allele_codes(code)
        pool = ["A", "C", "G", "T", "del12", "Q", "b", "a_long_allele_name", "0", "1"]
        allele_names = [sorted(rng.choice(pool, size=int(code[l].max()) + 1, replace=False).tolist()) for l in range(L)]
        g.set_allele_names(allele_names)
    return g, names, allele_names


@pytest.mark.parametrize("ploidy,n_ind,loci,n_alleles,codes", [
    (2, 130, [60, 33], 4, 2), (4, 90, [45, 20], 8, 2), (4, 300, [130], 8, 4), (6, 40, [35, 18], 12, 3),
    (2, 1500, [70], 24, 0), (4, 77, [19, 51], 8, 0), (8, 25, [40], 200, 0),
])
def test_text_tables_match_binary_tables(ploidy, n_ind, loci, n_alleles, codes):
    import pedigreesim_b200 as ps
    g, names, allele_names = make_ctx(ploidy, n_ind, loci, n_alleles, codes, seed=ploidy * 31 + n_ind)
    which = ps.DOSE_MAX_ALLELE if codes else 1
    binary = g.emit(founder=True, allele=bool(codes), dose=True, dose_which=which)
    text = g.emit_text(founder=True, allele=bool(codes), dose=True, dose_which=which)
    num = lambda l, v: str(int(v)).encode()
    assert text["founder"] == expected_lines(names, binary["founder"], num)
    assert text["dose"] == expected_lines(names, binary["dose"], num)
    if codes:
        assert text["allele"] == expected_lines(names, binary["allele"], lambda l, v: allele_names[l][int(v)].encode())
    # sub-ranges and single tables
    i0, ni, l0, nl = 3, n_ind - 7, 5, sum(loci) - 9
    sub = g.emit_text(founder=True, ind_first=i0, ind_count=ni, loc_first=l0, loc_count=nl)
    assert sub["founder"] == expected_lines(names[l0:l0 + nl], binary["founder"][l0:l0 + nl, i0 * ploidy:(i0 + ni) * ploidy], num)


def test_wide_founder_numbers_and_capacity_error():
    import pedigreesim_b200 as ps
    # 300 diploid founders -> founder alleles up to 599 (u16 cells, 3-digit tokens)
    n_f, n_kid = 300, 20
    rng = np.random.default_rng(4)
    lengths = [1.0]
    off, pos = _cases.uniform_map(lengths, [25])
    seg_off, founder, segpos = _cases.random_haplostructs(rng, n_kid, lengths, 2, 600, 4)
    p0 = np.concatenate([-np.ones(n_f, np.int32), np.zeros(n_kid, np.int32)])
    g = ps.PsimContext(2)
    g.set_chromosomes(lengths, [0.5])
    g.set_pedigree(p0, p0)
    g.set_haplostructs(n_f, seg_off, founder, segpos)
    g.simulate()
    g.set_map(off, pos)
    names = ["m%d" % l for l in range(25)]
    g.set_locus_names(names)
    binary = g.emit(founder=True, founder_bytes=2)["founder"]
    text = g.emit_text(founder=True)["founder"]
    assert text == expected_lines(names, binary, lambda l, v: str(int(v)).encode())
    with pytest.raises(ps.PsimError) as e:
        g.emit_text(founder=True, capacity=1000)
    assert e.value.code == ps.PSIM_ECAPACITY


def test_names_are_required():
    import pedigreesim_b200 as ps
    g, names, _ = make_ctx(2, 20, [10], 4, 0, seed=1)
    g2 = ps.PsimContext(2)
    with pytest.raises(ps.PsimError):
        g2.emit_text(founder=True)


@pytest.mark.parametrize("ploidy,n_ind,loci,n_alleles,codes", [(4, 80, [30, 12], 8, 3), (2, 150, [40], 14, 0), (6, 33, [21], 12, 0)])
def test_all_dose_text_matches_binary_rows(ploidy, n_ind, loci, n_alleles, codes):
    g, names, allele_names = make_ctx(ploidy, n_ind, loci, n_alleles, codes, seed=n_ind)
    ro, rows = g.emit_all_dose()
    expect = b""
    for l in range(sum(loci)):
        for k in range(int(ro[l + 1] - ro[l])):
            sub = allele_names[l][k] if codes else str(k)
            expect += names[l].encode() + b"\t" + sub.encode() + b"".join(b"\t" + str(int(v)).encode() for v in rows[ro[l] + k]) + b"\n"
    assert g.emit_all_dose_text() == expect
    l0, nl, i0, ni = 4, sum(loci) - 6, 2, n_ind - 5
    sub_ro, sub_rows = g.emit_all_dose(ind_first=i0, ind_count=ni, loc_first=l0, loc_count=nl)
    got = g.emit_all_dose_text(ind_first=i0, ind_count=ni, loc_first=l0, loc_count=nl)
    assert got.count(b"\n") == sub_rows.shape[0]
    first = got.split(b"\n")[0].split(b"\t")
    assert first[0] == names[l0].encode() and [int(x) for x in first[2:]] == sub_rows[0].tolist()
