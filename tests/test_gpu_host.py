"""The host program `pedigreesim x.par` over libpsim on the GPU against the oracle's twin of
`java -jar PedigreeSim.jar x.par`, file by file:

* replay (mode 2): .hsa/.hsb recorded by the Java-LCG restatement of the reference -> the genotype,
  founder-allele and dose files must be byte-identical (the bit-exact half of the north star);
* free-running (mode 1) and extend (mode 3) with Philox streams: every file identical, .hsb
  breakpoints to 1e-12 (CUDA log/exp vs libm), stdout identical.
"""
import os
import shutil
import subprocess

import numpy as np
import pytest

import _oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "manual_example")
DAT = ("_founderalleles.dat", "_genotypes.dat", "_alleledose.dat", "_allAlleledose.dat")


@pytest.fixture(scope="module")
def host_bin():
    from pedigreesim_b200 import build
    build.build()
    return build.build_host()


def write(d, name, text):
    with open(os.path.join(d, name), "w") as f:
        f.write(text)


def run_host(host_bin, cwd, par, *flags):
    return subprocess.run([host_bin, par, *flags], cwd=str(cwd), check=True, capture_output=True, text=True).stdout


def read(path):
    with open(path, "rb") as f:
        return f.read()


def assert_same_files(a, b, base, names, must_exist=()):
    for n in names:
        pa, pb = os.path.join(a, base + n), os.path.join(b, base + n)
        assert os.path.exists(pa) == os.path.exists(pb), n
        if n in must_exist:
            assert os.path.exists(pa), n
        if os.path.exists(pa):
            assert read(pa) == read(pb), "%s differs" % n


def assert_hsb_close(pa, pb):
    la, lb = open(pa).read().split("\n"), open(pb).read().split("\n")
    assert len(la) == len(lb)
    for x, y in zip(la, lb):
        if x == y:
            continue
        wa, wb = x.split("\t"), y.split("\t")
        assert len(wa) == len(wb)
        for u, v in zip(wa, wb):
            if u != v:
                assert abs(float(u) - float(v)) <= 1e-12 * max(1.0, abs(float(u))), (u, v)


def scenario(tmp_path, files, par_text):
    dirs = []
    for who in ("oracle", "host"):
        d = tmp_path / who
        d.mkdir()
        for n, t in files.items():
            write(d, n, t)
        write(d, "x.par", par_text)
        dirs.append(d)
    return dirs


def par_text(**kw):
    return "".join("%s = %s\n" % (k, v) for k, v in kw.items() if v is not None)


def chrom_text(ploidy, rows):
    head = "chromosome\tlength\tcentromere" + ("\tprefPairing\tquadrivalents\n" if ploidy > 2 else "\n")
    return head + "".join("\t".join(str(x) for x in r) + "\n" for r in rows)


def map_text(chroms, loci):
    out = "marker\tchromosome\tposition\n"
    for (name, length), n in zip(chroms, loci):
        for i in range(n):
            out += "%s_%03d\t%s\t%.4f\n" % (name, i, name, length * (i + 0.5) / n)
    return out


def gen_text(founders, ploidy, marker_names, n_names, seed):
    rng = np.random.default_rng(seed)
    out = "marker" + "".join("\t%s_%d" % (f, h + 1) for f in founders for h in range(ploidy)) + "\n"
    alphabet = ["A", "C", "G", "T", "del", "b10", "B9"][:n_names]
    for m in marker_names:
        out += m + "".join("\t" + alphabet[int(rng.integers(0, n_names))] for _ in range(len(founders) * ploidy)) + "\n"
    return out


def marker_names(map_txt):
    return [l.split("\t")[0] for l in map_txt.strip().split("\n")[1:]]


# ------------------------------------------------------------------------------------ replay
@pytest.mark.parametrize("ploidy,poptype,popsize,n_names", [(2, "F2", 200, 2), (2, "F2", 60, 4), (4, "F1", 150, 2),
                                                            (4, "S1", 40, 3), (6, "BC", 50, 2)])
def test_replay_of_recorded_haplostructs_is_byte_identical(host_bin, tmp_path, ploidy, poptype, popsize, n_names):
    """BASELINE configs[0] (diploid F2 of 200, 1 chromosome 100 cM, 50 loci) and polyploid variants:
    the Java-LCG restatement records .hsa/.hsb; both sides then re-genotype from the record."""
    chroms = [("chr1", 100.0)] if ploidy == 2 else [("A", 100.0), ("B", 85.0)]
    rows = [(n, l, l / 2) + ((0.4, 0.5) if ploidy > 2 else ()) for n, l in chroms]
    mp = map_text(chroms, [50] * len(chroms))
    founders = ["P1"] if poptype == "S1" else ["P1", "P2"]
    files = {"x.chrom": chrom_text(ploidy, rows), "x.map": mp,
             "x.gen": gen_text(founders, ploidy, marker_names(mp), n_names, seed=popsize)}
    rec = tmp_path / "record"
    rec.mkdir()
    for n, t in files.items():
        write(rec, n, t)
    write(rec, "x.par", par_text(PLOIDY=ploidy, MAPFUNCTION="KOSAMBI" if ploidy == 4 else "HALDANE", CHROMFILE="x.chrom",
                                 POPTYPE=poptype, POPSIZE=popsize, OUTPUT="rec", SEED=12345, NATURALPAIRING=int(ploidy != 4)))
    _oracle.run_parameter_file("x.par", cwd=str(rec))  # java.util.Random mode = the reference's own draws
    replay = par_text(PLOIDY=ploidy, CHROMFILE="x.chrom", PEDFILE="rec.ped", HAPLOSTRUCT="rec", MAPFILE="x.map",
                      FOUNDERFILE="x.gen", OUTPUT="out")
    do, dh = scenario(tmp_path, files, replay)
    for d in (do, dh):
        for ext in (".ped", ".hsa", ".hsb"):
            shutil.copy(rec / ("rec" + ext), d / ("rec" + ext))
    log_o = _oracle.run_parameter_file("x.par", cwd=str(do))
    log_h = run_host(host_bin, dh, "x.par")
    assert log_o == log_h
    assert_same_files(do, dh, "out", DAT, must_exist=DAT[:3])
    assert os.path.exists(dh / "out_allAlleledose.dat") == (n_names > 2)
    assert not os.path.exists(dh / "out.hsa")  # nothing was simulated: no new haplostruct files
    # and without founder genotypes: founder alleles, dose of founder allele 0, all founder-allele doses
    write(do, "y.par", replay.replace("FOUNDERFILE = x.gen\n", "").replace("OUTPUT = out", "OUTPUT = nf"))
    write(dh, "y.par", open(do / "y.par").read())
    assert _oracle.run_parameter_file("y.par", cwd=str(do)) == run_host(host_bin, dh, "y.par")
    assert_same_files(do, dh, "nf", DAT, must_exist=("_founderalleles.dat", "_alleledose.dat", "_allAlleledose.dat"))


def test_replay_of_hand_written_record_with_ties_and_redundant_segments(host_bin, tmp_path):
    # breakpoints on loci, equal-founder neighbours, a zero-length segment, E-notation, short lines
    files = {
        "x.chrom": chrom_text(2, [("c", 100.0, 30.0)]),
        "x.map": "marker chromosome position\nm1 c 0\nm2 c 25\nm3 c 25\nm4 c 50.00000000000001\nm5 c 100\n",
        "x.ped": "Name Parent1 Parent2\nP1 NA NA\nP2 NA NA\nK P1 P2\n",
        "r.hsa": "P1\tc\t1\t0\tNA\tNA\tNA\nP1\tc\t2\t1\tNA\tNA\tNA\nP2\tc\t1\t2\tNA\tNA\tNA\nP2\tc\t2\t3\tNA\tNA\tNA\n"
                 "K\tc\t1\t0\t1\t1\t0\nK\tc\t2\t3\t2\t2\tNA\n",
        "r.hsb": "\tNA\tNA\tNA\n\tNA\tNA\tNA\n\tNA\tNA\tNA\n\tNA\tNA\tNA\n\t25.0\t50.00000000000001\t1.0E2\n\t2.5E1\t25.0\tNA\n",
    }
    do, dh = scenario(tmp_path, files, par_text(PLOIDY=2, CHROMFILE="x.chrom", PEDFILE="x.ped", HAPLOSTRUCT="r",
                                                MAPFILE="x.map", OUTPUT="out"))
    assert _oracle.run_parameter_file("x.par", cwd=str(do)) == run_host(host_bin, dh, "x.par")
    assert_same_files(do, dh, "out", DAT, must_exist=("_founderalleles.dat",))
    rows = open(dh / "out_founderalleles.dat").read().split("\n")
    assert rows[0] == "marker\tP1_1\tP1_2\tP2_1\tP2_2\tK_1\tK_2"
    assert [r.split("\t")[5:] for r in rows[1:6]] == [["0", "3"], ["1", "2"], ["1", "2"], ["1", "2"], ["0", "2"]]


# ------------------------------------------------------------------------------------ free running
def test_manual_example_every_file_matches_the_philox_oracle(host_bin, tmp_path):
    dirs = []
    for who in ("oracle", "host"):
        d = tmp_path / who
        shutil.copytree(GOLDEN, d)
        dirs.append(d)
    do, dh = dirs
    log_o = _oracle.run_parameter_file("example.par", rng=_oracle.RNG_PHILOX, cwd=str(do))
    log_h = run_host(host_bin, dh, "example.par")
    assert log_o == log_h
    base = "example_out"
    assert_same_files(do, dh, base, DAT + (".hsa", "_meioticconfigs.dat", ".ped"), must_exist=DAT[:3] + (".hsa", "_meioticconfigs.dat"))
    assert_hsb_close(do / (base + ".hsb"), dh / (base + ".hsb"))


@pytest.mark.parametrize("ploidy,kosambi,natural", [(2, False, True), (4, True, False), (6, False, True), (8, False, True)])
def test_standard_populations_match_the_philox_oracle(host_bin, tmp_path, ploidy, kosambi, natural):
    chroms = [("L1", 120.0), ("L2", 70.0), ("L3", 95.5)]
    rows = [(n, l, 0.3 * l) + ((0.35, 0.6) if ploidy > 2 else ()) for n, l in chroms]
    mp = map_text(chroms, [40, 17, 33])
    files = {"x.chrom": chrom_text(ploidy, rows), "x.map": mp}
    do, dh = scenario(tmp_path, files, par_text(PLOIDY=ploidy, MAPFUNCTION="KOSAMBI" if kosambi else "HALDANE", CHROMFILE="x.chrom",
                                                MAPFILE="x.map", POPTYPE="F2", POPSIZE=90, OUTPUT="o", SEED=99,
                                                NATURALPAIRING=int(natural), PAIREDCENTROMERES=0.3, PARALLELMULTIVALENTS=0.25))
    assert _oracle.run_parameter_file("x.par", rng=_oracle.RNG_PHILOX, cwd=str(do)) == run_host(host_bin, dh, "x.par")
    assert_same_files(do, dh, "o", DAT + (".hsa", "_meioticconfigs.dat", ".ped"),
                      must_exist=("_founderalleles.dat", "_alleledose.dat", "_allAlleledose.dat", ".hsa"))
    assert_hsb_close(do / "o.hsb", dh / "o.hsb")


def test_extend_mode_keeps_the_record_and_simulates_the_rest(host_bin, tmp_path):
    """Mode 3 (Main.java:343-373): individuals in the record keep their haplostructs and meiotic
    configurations; new pedigree members (including a late founder) are simulated."""
    chroms = [("A", 100.0), ("B", 60.0)]
    files = {"x.chrom": chrom_text(4, [(n, l, 0.5 * l, 0.5, 0.5) for n, l in chroms]), "x.map": map_text(chroms, [25, 15])}
    do, dh = scenario(tmp_path, files, par_text(PLOIDY=4, MAPFUNCTION="HALDANE", CHROMFILE="x.chrom", POPTYPE="F1", POPSIZE=20,
                                                OUTPUT="g1", SEED=5, NATURALPAIRING=0))
    assert _oracle.run_parameter_file("x.par", rng=_oracle.RNG_PHILOX, cwd=str(do)) == run_host(host_bin, dh, "x.par")
    ped = open(do / "g1.ped").read() + "P3\tNA\tNA\n" + "".join("X%02d\tF1_%02d\tP3\n" % (k, k + 1) for k in range(12)) + \
        "".join("Y%02d\tX%02d\tX%02d\n" % (k, k, (k + 1) % 12) for k in range(12))
    ext = par_text(PLOIDY=4, MAPFUNCTION="HALDANE", CHROMFILE="x.chrom", PEDFILE="g2.ped", HAPLOSTRUCT="g1", MAPFILE="x.map",
                   OUTPUT="g2out", SEED=6, NATURALPAIRING=0)
    for d in (do, dh):
        write(d, "g2.ped", ped)
        write(d, "x.par", ext)
        # the host's own record feeds the host, the oracle's the oracle: they are equal up to .hsb ulps
    assert _oracle.run_parameter_file("x.par", rng=_oracle.RNG_PHILOX, cwd=str(do)) == run_host(host_bin, dh, "x.par")
    assert_same_files(do, dh, "g2out", (".hsa", "_meioticconfigs.dat", "_founderalleles.dat", "_alleledose.dat", "_allAlleledose.dat"),
                      must_exist=(".hsa", "_meioticconfigs.dat", "_founderalleles.dat"))
    assert_hsb_close(do / "g2out.hsb", dh / "g2out.hsb")
    cfg = open(dh / "g2out_meioticconfigs.dat").read().split("\n")
    assert cfg[0].split("\t")[2] == "NA" and cfg[2 * 2].split("\t")[2] != "NA"  # founders missing, F1 kept from g1
    # P3 is the third founder: founder alleles 8..11
    fa = open(dh / "g2out_founderalleles.dat").read().split("\n")
    head = fa[0].split("\t")
    col = head.index("P3_1")
    assert fa[1].split("\t")[col:col + 4] == ["8", "9", "10", "11"]


def test_many_founders_use_the_wide_founder_table(host_bin, tmp_path):
    # 140 diploid founders -> 280 founder alleles (> 256): u16 founder cells in libpsim, 3-digit text
    n_f = 140
    ped = "Name\tParent1\tParent2\n" + "".join("F%03d\tNA\tNA\n" % k for k in range(n_f)) + \
        "".join("K%03d\tF%03d\tF%03d\n" % (k, k, (k * 7 + 1) % n_f) for k in range(n_f))
    files = {"x.chrom": chrom_text(2, [("c1", 100.0, 50.0)]), "x.map": map_text([("c1", 100.0)], [30]), "x.ped": ped}
    do, dh = scenario(tmp_path, files, par_text(PLOIDY=2, MAPFUNCTION="HALDANE", CHROMFILE="x.chrom", PEDFILE="x.ped",
                                                MAPFILE="x.map", OUTPUT="w", SEED=3))
    assert _oracle.run_parameter_file("x.par", rng=_oracle.RNG_PHILOX, cwd=str(do)) == run_host(host_bin, dh, "x.par")
    assert_same_files(do, dh, "w", (".hsa", "_founderalleles.dat", "_alleledose.dat", "_allAlleledose.dat"),
                      must_exist=(".hsa", "_founderalleles.dat", "_allAlleledose.dat"))
    assert "\t279" in open(dh / "w_founderalleles.dat").read()


def test_small_blocks_give_the_same_files(host_bin, tmp_path):
    # force many emission blocks (1 MiB) through the chunked writers
    chroms = [("A", 100.0)]
    files = {"x.chrom": chrom_text(4, [("A", 100.0, 40.0, 0.2, 0.3)]), "x.map": map_text(chroms, [700])}
    par = par_text(PLOIDY=4, MAPFUNCTION="HALDANE", CHROMFILE="x.chrom", MAPFILE="x.map", POPTYPE="F1", POPSIZE=900,
                   OUTPUT="o", SEED=11)
    da, db = scenario(tmp_path, files, par)
    assert run_host(host_bin, da, "x.par") == run_host(host_bin, db, "x.par", "--block-mib", "1")
    assert_same_files(da, db, "o", DAT + (".hsa", ".hsb", "_meioticconfigs.dat"), must_exist=("_founderalleles.dat", "_allAlleledose.dat"))


# ------------------------------------------------------------------------------------ jar-recorded goldens
JAR_GOLDEN = os.path.join(ROOT, "tests", "golden", "jar")
JAR_CASES = sorted(os.listdir(JAR_GOLDEN)) if os.path.isdir(JAR_GOLDEN) else []


@pytest.mark.skipif(not JAR_CASES, reason="no jar-recorded goldens: run tests/golden/make_jar_goldens.sh on a machine with a JVM")
@pytest.mark.parametrize("case", JAR_CASES or ["none"])
def test_replay_of_jar_recorded_haplostructs(host_bin, tmp_path, case):
    """north_star's first correctness leg: the replay mode consumes the reference's OWN recorded
    out.hsa / out.hsb (written by PedigreeSim.jar, Main.writeHaploStructFiles Main.java:1100-1136) and
    must reproduce the jar's genotype, dose and founder-allele files byte for byte."""
    g = os.path.join(JAR_GOLDEN, case)
    par = dict(l.strip().split(" = ") for l in open(os.path.join(g, "case.par")) if " = " in l)
    for n in ("case.chrom", "case.map", "case.gen", "out.ped", "out.hsa", "out.hsb"):
        shutil.copy(os.path.join(g, n), tmp_path / n)
    write(tmp_path, "replay.par", par_text(PLOIDY=par["PLOIDY"], CHROMFILE="case.chrom", PEDFILE="out.ped", HAPLOSTRUCT="out",
                                           MAPFILE="case.map", FOUNDERFILE="case.gen", OUTPUT="again", MISSING=par.get("MISSING", "NA")))
    run_host(host_bin, tmp_path, "replay.par")
    for n in DAT:
        if os.path.exists(os.path.join(g, "out" + n)):
            assert read(tmp_path / ("again" + n)) == read(os.path.join(g, "out" + n)), n
