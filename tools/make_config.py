#!/usr/bin/env python
"""Write PedigreeSim input files (.par/.chrom/.map/.gen[/.ped]) for the BASELINE.json configs as
SURVEY.md §8(d) spells them out, at full or reduced size — inputs for `pedigreesim x.par` (the host
program) and for the oracle's CLI twin of the jar.

    python tools/make_config.py <config 1..5> <outdir> [--popsize N] [--loci N] [--chroms N]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write(path, text):
    with open(path, "w") as f:
        f.write(text)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", type=int)
    ap.add_argument("outdir")
    ap.add_argument("--popsize", type=int)
    ap.add_argument("--loci", type=int)
    ap.add_argument("--chroms", type=int)
    ap.add_argument("--generations", type=int, default=8)
    a = ap.parse_args()
    os.makedirs(a.outdir, exist_ok=True)
    d = a.outdir
    par = {"MISSING": "NA", "CHROMFILE": "x.chrom", "MAPFILE": "x.map", "OUTPUT": "out", "SEED": 12345}
    gen = None
    if a.config == 1:
        P, C, loci, pop = 2, 1, a.loci or 50, a.popsize or 200
        lengths = [100.0]
        rows = [("chr1", 100.0, 50.0)]
        par.update(PLOIDY=2, MAPFUNCTION="HALDANE", POPTYPE="F2", POPSIZE=pop, FOUNDERFILE="x.gen")
        gen = lambda l, a_: "AB"[a_ // 2]
    elif a.config == 2:
        import bench
        P, C, loci, pop = 4, a.chroms or 12, a.loci or 2000, a.popsize or 10000
        lengths = [80.0 + 10 * (c % 5) for c in range(C)]
        rows = [("c%02d" % c, lengths[c], 0.4 * lengths[c], 0.3, 0.5) for c in range(C)]
        par.update(PLOIDY=4, MAPFUNCTION="HALDANE", NATURALPAIRING=0, POPTYPE="F1", POPSIZE=pop, FOUNDERFILE="x.gen")
        code = bench.workload(2)["code"] if (C, loci) == (12, 2000) else None
        gen = (lambda l, a_: str(int(code[l, a_]))) if code is not None else (lambda l, a_: str((l * 7 + a_ * 3 + (l >> 3)) & 1))
    elif a.config == 3:
        P, C, loci, pop = 6, a.chroms or 21, a.loci or 10000, a.popsize or 100000
        lengths = [100.0] * C
        rows = [("c%02d" % c, 100.0, 50.0, 0.0, 0.0) for c in range(C)]
        par.update(PLOIDY=6, MAPFUNCTION="HALDANE", NATURALPAIRING=1, POPTYPE="F1", POPSIZE=pop)
    elif a.config == 4:
        P, C, loci, pop = 2, a.chroms or 10, a.loci or 50000, a.popsize or 1000000
        lengths = [150.0] * C
        rows = [("c%02d" % c, 150.0, 75.0) for c in range(C)]
        par.update(PLOIDY=2, MAPFUNCTION="HALDANE", PEDFILE="x.ped")
        with open(os.path.join(d, "x.ped"), "w") as f:
            f.write("Name\tParent1\tParent2\nP1\tNA\tNA\nP2\tNA\tNA\nF1\tP1\tP2\n")
            for g in range(2, a.generations + 1):
                for i in range(1, pop + 1):
                    par_name = "F1" if g == 2 else "F%d_%d" % (g - 1, i)
                    f.write("F%d_%d\t%s\t%s\n" % (g, i, par_name, par_name))
    elif a.config == 5:
        P, C, loci, pop = 4, a.chroms or 5, a.loci or 100, a.popsize or 200
        lengths = [100.0] * C
        rows = [("c%02d" % c, 100.0, 50.0, 0.25, 0.0) for c in range(C)]
        par.update(PLOIDY=4, MAPFUNCTION="HALDANE", NATURALPAIRING=1, POPTYPE="F1", POPSIZE=pop)
    else:
        raise SystemExit("config must be 1..5")
    head = "chromosome\tlength\tcentromere" + ("\tprefPairing\tquadrivalents\n" if P > 2 else "\n")
    write(os.path.join(d, "x.chrom"), head + "".join("\t".join(str(x) for x in r) + "\n" for r in rows))
    names = []
    with open(os.path.join(d, "x.map"), "w") as f:
        f.write("marker\tchromosome\tposition\n")
        for c in range(C):
            for i in range(loci):
                nm = "m%02d_%05d" % (c, i)
                names.append(nm)
                f.write("%s\t%s\t%r\n" % (nm, rows[c][0], lengths[c] * (i + 0.5) / loci))
    if gen is not None:
        founders = ["P1", "P2"]
        with open(os.path.join(d, "x.gen"), "w") as f:
            f.write("marker" + "".join("\t%s_%d" % (fn, h + 1) for fn in founders for h in range(P)) + "\n")
            for l, nm in enumerate(names):
                f.write(nm + "".join("\t" + gen(l, a_) for a_ in range(2 * P)) + "\n")
    write(os.path.join(d, "x.par"), "".join("%s = %s\n" % kv for kv in par.items()))
    print("config %d: ploidy %d, %d chromosomes x %d loci, popsize %d -> %s/x.par" % (a.config, P, C, loci, pop, d))


if __name__ == "__main__":
    main()
