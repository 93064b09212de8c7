#!/usr/bin/env python
"""Inputs of the jar golden cases (tests/golden/make_jar_goldens.sh): BASELINE.json configs[0] (the manual's diploid
example: F1 selfed to a 200-individual F2, one chromosome of 100 cM, 50 loci) and a small tetraploid F1 with
quadrivalents and preferential pairing (configs[1] at reduced size), both with founder genotypes.

    python tests/golden/make_jar_cases.py <outdir>
"""
import os
import sys


def write(path, text):
    with open(path, "w") as f:
        f.write(text)


def case(outdir, name, par, chrom, loci_per_chrom, ploidy, n_founders, kinds):
    d = os.path.join(outdir, name)
    os.makedirs(d, exist_ok=True)
    write(os.path.join(d, "case.par"), "".join("%s = %s\n" % kv for kv in par.items()))
    head = "chromosome\tlength\tcentromere" + ("\tprefPairing\tquadrivalents" if ploidy > 2 else "") + "\n"
    write(os.path.join(d, "case.chrom"), head + "".join("\t".join(str(x) for x in row) + "\n" for row in chrom))
    m, g = ["marker\tchromosome\tposition\n"], None
    names = []
    for row in chrom:
        for l in range(loci_per_chrom):
            nm = "%s_m%03d" % (row[0], l)
            names.append(nm)
            m.append("%s\t%s\t%.4f\n" % (nm, row[0], row[1] * (l + 0.5) / loci_per_chrom))
    write(os.path.join(d, "case.map"), "".join(m))
    founders = ["P%d" % (k + 1) for k in range(n_founders)]
    g = ["marker\t" + "\t".join("%s_%d" % (f, h + 1) for f in founders for h in range(ploidy)) + "\n"]
    for l, nm in enumerate(names):
        g.append(nm + "\t" + "\t".join(kinds[(l * 7 + a * 5 + (l >> 2)) % len(kinds)] for a in range(n_founders * ploidy)) + "\n")
    write(os.path.join(d, "case.gen"), "".join(g))


def main():
    out = sys.argv[1]
    common = {"MISSING": "NA", "CHROMFILE": "case.chrom", "MAPFILE": "case.map", "FOUNDERFILE": "case.gen", "OUTPUT": "out", "SEED": 12345}
    case(out, "configs0_diploid_f2", dict(common, PLOIDY=2, MAPFUNCTION="HALDANE", POPTYPE="F2", POPSIZE=200),
         [("chr1", 100.0, 50.0)], 50, 2, 2, ["A", "B"])
    case(out, "tetraploid_f1", dict(common, PLOIDY=4, MAPFUNCTION="KOSAMBI", NATURALPAIRING=0, POPTYPE="F1", POPSIZE=120, SEED=777),
         [("c1", 100.0, 40.0, 0.3, 0.5), ("c2", 80.0, 30.0, 0.0, 1.0)], 40, 4, 2, ["A", "B", "C"])


if __name__ == "__main__":
    main()
