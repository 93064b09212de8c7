"""Host program `pedigreesim` (pedigreesim_b200/host/), CPU side: file parsing, pedigree ordering,
messages and number formatting must agree with the oracle's restatement of Main.java /
PopulationData.java. `--parse-only` stops before the first libpsim call, so nothing here needs a
GPU; the GPU end of the host is covered by tests/test_gpu_host.py."""
import os
import struct
import subprocess

import numpy as np
import pytest

import _oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "manual_example")


@pytest.fixture(scope="module")
def host_bin():
    from pedigreesim_b200 import build
    build.build()
    return build.build_host()


def run_host(host_bin, cwd, par="x.par", *flags):
    return subprocess.run([host_bin, par, *flags], cwd=cwd, check=True, capture_output=True, text=True).stdout


def write(d, name, text):
    with open(os.path.join(d, name), "w") as f:
        f.write(text)


CHROM2 = "chromosome length centromere\nc1 100 50\nc2 80.5 10\n"
MAP2 = "marker chromosome position\n" + "".join("m%d c1 %g\n" % (i, 10.0 * i) for i in range(11)) + \
       "".join("n%d c2 %g\n" % (i, 8.05 * i) for i in range(11))


def base_par(**kw):
    d = dict(PLOIDY=2, MAPFUNCTION="HALDANE", CHROMFILE="x.chrom", MAPFILE="x.map", OUTPUT="out", SEED=7)
    d.update(kw)
    return "".join("%s = %s\n" % (k, v) for k, v in d.items() if v is not None)


def both_logs(host_bin, tmp_path, par_text, files):
    """Run the oracle CLI and the host on the same inputs in separate directories."""
    logs, dirs = [], []
    for who in ("oracle", "host"):
        d = tmp_path / who
        d.mkdir(exist_ok=True)
        write(d, "x.par", par_text)
        for n, t in files.items():
            write(d, n, t)
        if who == "oracle":
            logs.append(_oracle.run_parameter_file("x.par", cwd=str(d)))
        else:
            logs.append(run_host(host_bin, str(d), "x.par", "--parse-only"))
        dirs.append(d)
    return logs, dirs


def test_java_double_to_string_twin(host_bin):
    rng = np.random.default_rng(5)
    vals = [0.0, -0.0, 1.0, 25.0, 1e-3, 9.999e-4, 1e7, 9999999.0, 1.0e-4, 123456.789, 0.1 + 0.2, 100.0 / 3.0, 5e-324,
            1.7976931348623157e308, 2.5e-5 * 100.0, 73.90000000000001, 1e21, 1e22, 1e23, 4.35, 0.3 * 100.0]
    vals += (rng.random(3000) * 150.0).tolist()
    vals += (10.0 ** rng.uniform(-12, 12, 2000) * rng.choice([-1, 1], 2000)).tolist()
    vals += np.frombuffer(rng.bytes(8 * 2000), dtype=np.float64)[:].tolist()
    vals = [v for v in vals if v == v and abs(v) != float("inf")]
    inp = "\n".join("%016x" % struct.unpack("<Q", struct.pack("<d", v))[0] for v in vals) + "\n"
    out = subprocess.run([host_bin, "--format-doubles"], input=inp, check=True, capture_output=True, text=True).stdout.split()
    assert len(out) == len(vals)
    L = _oracle.lib()
    import ctypes as C
    buf = C.create_string_buffer(64)
    for v, got in zip(vals, out):
        L.orc_java_double_to_string(v, buf, 64)
        assert got == buf.value.decode(), (v, got, buf.value)
        assert float(got) == v  # round trip, what .hsb replay relies on


def test_parameter_file_forms_and_rejections(host_bin, tmp_path):
    files = {"x.chrom": CHROM2, "x.map": MAP2}
    ok = "PLOIDY=2\nmapfunction =haldane\nCHROMFILE = x.chrom ; trailing words are ignored\n; comment\n\n" \
         "MAPFILE=x.map\nOUTPUT = out\nSEED=-3\nPOPTYPE = F2\nPOPSIZE=12\nMISSING = *\n"
    (lo, lh), (do, dh) = both_logs(host_bin, tmp_path, ok, files)
    assert "parse-only: 15 individuals, 4 founder alleles, 2 chromosomes, 22 loci" in lh
    assert open(do / "out.ped").read() == open(dh / "out.ped").read()
    assert "F2_01\tF1\tF1" in open(dh / "out.ped").read() and "P1\t*\t*" in open(dh / "out.ped").read()
    for k, bad in enumerate(["PLOIDY = 3\n", "POPTYPE = f2\n", "POPSIZE = 0\n", "MAPFUNCTION = MORGAN\n", "NATURALPAIRING = maybe\n",
                             "PAIREDCENTROMERES = 1.5\n", "SEED = 1.5\n", "FOO = 1\n", "PLOIDY =\n", "PLOIDY= 2\n", "=2\n"]):
        sub = tmp_path / ("bad%d" % k)
        sub.mkdir()
        (a, b), _ = both_logs(host_bin, sub, bad + ok, files)
        assert a == b and "Invalid line in parameter file" in b, bad


@pytest.mark.parametrize("case", [
    dict(par=dict(PLOIDY=None), expect="no valid PLOIDY found"),
    dict(par=dict(MAPFUNCTION=None), expect="no valid MAPFUNCTION found"),
    dict(par=dict(CHROMFILE=None), expect="no CHROMFILE specified"),
    dict(par=dict(OUTPUT=None), expect="no OUTPUT specified"),
    dict(par=dict(CHROMFILE="nope.chrom"), expect="CHROMFILE nope.chrom not found or not readable"),
    dict(par=dict(), expect="Either a valid PEDFILE or both POPTYPE and POPSIZE must be specified"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3), files={"x.chrom": "chrom len cen\nc1 100 50\n"}, expect="header line not found"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3), files={"x.chrom": "chromosome length centromere\nc1 100\n"}, expect="insufficient data"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3), files={"x.chrom": "chromosome length centromere\nc1 NA 50\n"}, expect="missing data are not allowed"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3, PLOIDY=4), expect="insufficient data"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3, ALLOWNORECOMBINATION=0), files={"x.chrom": "chromosome length centromere\nc1 60 30\n"},
         expect="minimum chromosome length is 70.0 cM"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3), files={"x.map": "marker chromosome position\nm1 c9 5\n"}, expect="chromName c9 not present"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3), files={"x.map": "marker chromosome position\nm1 c1 100.1\n"}, expect="addLocus: position invalid"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3), files={"x.map": "marker chromosome position\nm1 c1\n"}, expect="too few items"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3, MAPFILE="nope.map"), expect="MAPFILE nope.map not found or not readable"),
    dict(par=dict(PEDFILE="x.ped"), files={"x.ped": "Name Parent1\nA NA NA\n"}, expect="header line not found"),
    dict(par=dict(PEDFILE="x.ped"), files={"x.ped": "Name Parent1 Parent2\nA NA NA\nB A\n"}, expect="incomplete line : B A"),
    dict(par=dict(PEDFILE="x.ped"), files={"x.ped": "Name Parent1 Parent2\nA NA NA\nB A B\n"}, expect="is it's own parent"),
    dict(par=dict(PEDFILE="x.ped"), files={"x.ped": "Name Parent1 Parent2\nA NA NA\nNA A A\n"}, expect="has a missing name"),
    dict(par=dict(PEDFILE="x.ped"), files={"x.ped": "Name Parent1 Parent2\nA NA NA\nB A NA\n"}, expect="one known and one unknown parent"),
    dict(par=dict(PEDFILE="x.ped"), files={"x.ped": "Name Parent1 Parent2\nA NA NA\nB A A\nB A A\n"}, expect="occurs more than once"),
    dict(par=dict(PEDFILE="x.ped"), files={"x.ped": "Name Parent1 Parent2\nA NA NA\nB A Z\n"},
         expect="Z is listed as parent of B but it does not occur in the population"),
    dict(par=dict(PEDFILE="x.ped"), files={"x.ped": "Name Parent1 Parent2\nA NA NA\nB C A\nC B A\n"},
         expect="Individual.setParents: parents may not be null"),  # the rotation places C before its parent B
    # ("circular references" is unreachable: the last row of any cycle always has its parent before it)
    dict(par=dict(PEDFILE="x.ped"), files={"x.ped": "Name Parent1 Parent2\nB C C\nC B B\n"}, expect="parents may not be null"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3, FOUNDERFILE="x.gen"), files={"x.gen": "marker P1_1 P1_2 P2_1\n"},
         expect="must contain one column for the marker name and ploidy columns for each founder"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3, FOUNDERFILE="x.gen"), files={"x.gen": "marker P1_1 P1_2 F1_1_1 F1_1_2\n"},
         expect="F1_1 is not a founder"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3, FOUNDERFILE="x.gen"), files={"x.gen": "marker P1_1 P1_3 P2_1 P2_2\n"},
         expect="P1_3 found but P1_2 expected"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3, FOUNDERFILE="x.gen"), files={"x.gen": "marker P1_1 P1_2 P1_1 P1_2\n"},
         expect="occurs more than once in header"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3, FOUNDERFILE="x.gen"), files={"x.gen": "marker P1_1 P1_2 P2_1 P2_2\nm1 a b a b\n"},
         expect="invalid marker: m1"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3, FOUNDERFILE="x.gen"), files={"x.gen": "marker P1_1 P1_2 P2_1 P2_2\nm0 a b a\n"},
         expect="not all lines have 5 items"),
    dict(par=dict(POPTYPE="F1", POPSIZE=3, FOUNDERFILE="x.gen"), files={"x.gen": "marker P1_1 P1_2 P2_1 P2_2\nm0 a b a b\n"},
         expect="end of file found while reading chromosome c1"),
])
def test_messages_match_the_reference_restatement(host_bin, tmp_path, case):
    files = {"x.chrom": CHROM2, "x.map": MAP2}
    files.update(case.get("files", {}))
    (lo, lh), _ = both_logs(host_bin, tmp_path, base_par(**case["par"]), files)
    assert case["expect"] in lh, lh
    assert lo == lh


def _random_pedigree_text(rng, n, shuffle):
    rows = [("I%04d" % i, "NA", "NA") for i in range(4)]
    for i in range(4, n):
        a, b = rng.integers(0, i, 2)
        rows.append(("I%04d" % i, rows[a][0], rows[b][0]))
    order = np.arange(n)
    if shuffle == "full":
        rng.shuffle(order)
    elif shuffle == "blocks":  # late founders and a few displaced sib blocks: what real files look like
        blocks = np.array_split(order, 9)
        rng.shuffle(blocks)
        order = np.concatenate(blocks)
    return "Name Parent1 Parent2\n" + "".join("%s %s %s\n" % rows[k] for k in order)


@pytest.mark.parametrize("shuffle,seed", [("none", 1), ("blocks", 2), ("blocks", 3), ("full", 4), ("full", 5), ("full", 6)])
def test_pedigree_order_matches_sortpedigree(host_bin, tmp_path, shuffle, seed):
    rng = np.random.default_rng(seed)
    ped = _random_pedigree_text(rng, 120, shuffle)
    files = {"x.chrom": CHROM2, "x.map": MAP2, "x.ped": ped}
    (lo, lh), (do, dh) = both_logs(host_bin, tmp_path, base_par(PEDFILE="x.ped"), files)
    wrote = os.path.exists(do / "out.ped")
    assert wrote == os.path.exists(dh / "out.ped")
    assert wrote == (shuffle != "none")
    if wrote:
        assert open(do / "out.ped").read() == open(dh / "out.ped").read()
    if "Error" not in lo:
        assert ("write file out.ped" in lh) == wrote


def test_large_pedigree_loads_fast(host_bin, tmp_path):
    # 300k individuals, founders listed last: O(N^2) name lookups would take minutes
    n = 300_000
    lines = ["Name Parent1 Parent2"]
    lines += ["L%d F1 F1" % i for i in range(n)]
    lines += ["F1 P1 P2", "P1 NA NA", "P2 NA NA"]
    write(tmp_path, "x.ped", "\n".join(lines) + "\n")
    write(tmp_path, "x.chrom", CHROM2)
    write(tmp_path, "x.map", MAP2)
    write(tmp_path, "x.par", base_par(PEDFILE="x.ped"))
    import time
    t = time.time()
    log = run_host(host_bin, str(tmp_path), "x.par", "--parse-only")
    assert time.time() - t < 20
    assert "parse-only: 300003 individuals" in log
    head = open(tmp_path / "out.ped").read(200).split("\n")
    # the reference's rotation: P1 is the first placeable row, then P2 (F1 and the lines rotate behind)
    assert head[1].split() == ["P1", "NA", "NA"] and head[2].split() == ["P2", "NA", "NA"] and head[3].split()[0] == "F1"


def test_configs3_pedigree_of_seven_million_rows_is_ingested(host_bin, tmp_path):
    """BASELINE.json configs[3] at full size: the RIL chain P1 x P2 -> F1 -> 1 000 000 F2 lines selfed to F8 is a .ped of
    7 000 003 rows (PopulationData.java:244-252 reads it, sortPedigree :309-464 orders it). The file lists the youngest
    generation first and the founders last, every generation in reverse line order: every row has to wait for its
    parents (the reference's rotation takes one pass per generation) and names are looked up far from where they are
    defined. Reading, name resolution, parents-first ordering and writing out.ped back must stay linear per pass."""
    import time
    lines, gens = 1_000_000, 8
    with open(tmp_path / "x.ped", "w") as f:
        f.write("Name\tParent1\tParent2\n")
        for g in range(gens, 2, -1):
            f.write("".join("F%d_%d\tF%d_%d\tF%d_%d\n" % (g, k, g - 1, k, g - 1, k) for k in range(lines - 1, -1, -1)))
        f.write("".join("F2_%d\tF1\tF1\n" % k for k in range(lines - 1, -1, -1)))
        f.write("F1\tP1\tP2\nP1\tNA\tNA\nP2\tNA\tNA\n")
    write(tmp_path, "x.chrom", CHROM2)
    write(tmp_path, "x.map", MAP2)
    write(tmp_path, "x.par", base_par(PEDFILE="x.ped"))
    t = time.time()
    log = run_host(host_bin, str(tmp_path), "x.par", "--parse-only")
    took = time.time() - t
    assert "parse-only: 7000003 individuals" in log
    assert took < 120, took                                   # (measured: about 10 s, of which the 170 MB out.ped about 3)
    with open(tmp_path / "out.ped") as f:
        head = [f.readline().split() for _ in range(5)]
        assert head[1] == ["P1", "NA", "NA"] and head[2] == ["P2", "NA", "NA"] and head[3] == ["F1", "P1", "P2"]
        assert head[4] == ["F2_999999", "F1", "F1"]
        n = 5 + sum(1 for _ in f)
    assert n == 7000004


def test_map_is_position_sorted_with_stable_ties_and_clamped(host_bin, tmp_path):
    # unsorted map with ties and a position a hair outside the chromosome: the .gen row order the
    # reference demands is the sorted one (Main.java:863-867), so a .gen in that order must load
    m = "marker chromosome position\nb c1 50\na c1 10\nc c1 50\nd c1 100.000001\ne c1 0\nz c2 5\n"
    gen_ok = "marker P1_1 P1_2 P2_1 P2_2\n" + "".join("%s x y x y\n" % k for k in "eabcdz")
    gen_bad = "marker P1_1 P1_2 P2_1 P2_2\n" + "".join("%s x y x y\n" % k for k in "eacbdz")
    for gen, ok in ((gen_ok, True), (gen_bad, False)):
        d = tmp_path / ("ok" if ok else "bad")
        d.mkdir()
        files = {"x.chrom": CHROM2, "x.map": m, "x.gen": gen}
        (lo, lh), _ = both_logs(host_bin, d, base_par(POPTYPE="F1", POPSIZE=2, FOUNDERFILE="x.gen"), files)
        assert ("invalid marker" not in lh) == ok
        if not ok:
            assert lo == lh


def test_manual_example_parses(host_bin, tmp_path):
    for f in os.listdir(GOLDEN):
        write(tmp_path, f, open(os.path.join(GOLDEN, f)).read())
    log = run_host(host_bin, str(tmp_path), "example.par", "--parse-only")
    assert "parse-only: 7 individuals, 8 founder alleles, 2 chromosomes, 8 loci, 0 recorded haplostructs" in log


def test_host_sources_do_not_touch_the_oracle():
    host = os.path.join(ROOT, "pedigreesim_b200", "host")
    for f in os.listdir(host):
        src = open(os.path.join(host, f)).read()
        assert "oracle" not in src.replace("no CPU path", "")
