#!/usr/bin/env python
"""Regenerate tests/golden/manual_example/* from the reference's own manual.

The reference ships no test data; the only fixture-like content is the example input files
printed (tab-separated) in Manual/PedigreeSim_manual.docx, sections "Parameter file",
"Pedigree file", "Chromosome file", "Linkage map file" and "Founder genotypes file", plus the
known-answer Tables 1 and 2. This script extracts them verbatim (tabs preserved). It reads
/root/reference, so it only runs in the authoring container; its outputs are committed.
"""
import json
import os
import re
import sys
import zipfile

DOCX = "/root/reference/Manual/PedigreeSim_manual.docx"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "manual_example")


def paragraphs():
    x = zipfile.ZipFile(DOCX).read("word/document.xml").decode("utf8")
    for p in re.findall(r"<w:p[ >].*?</w:p>", x, flags=re.S):
        p = re.sub(r"<w:pPr>.*?</w:pPr>", "", p, flags=re.S)
        toks = re.findall(r"<w:t(?: [^>]*)?>(.*?)</w:t>|(<w:tab/>)", p, flags=re.S)
        yield "".join(a if a else "\t" for a, b in toks)


def block(lines, first_prefix, n_max=40):
    """Lines from the one starting with first_prefix until the first prose line."""
    i = next(k for k, l in enumerate(lines) if l.startswith(first_prefix))
    out = []
    for l in lines[i:i + n_max]:
        if len(l) > 90 or (l and "\t" not in l and " = " not in l and out):
            break
        out.append(l)
    while out and not out[-1].strip():
        out.pop()
    return out


def main():
    lines = list(paragraphs())
    os.makedirs(OUT, exist_ok=True)
    files = {
        "example.ped": block(lines, "Name\tParent1"),
        "example.chrom": block(lines, "chromosome\tlength"),
        "example.map": block(lines, "marker\tchromosome"),
        "example.gen": block(lines, "marker\tP1_1"),
    }
    i = next(k for k, l in enumerate(lines) if l.startswith("PLOIDY = 4"))
    par = [l.rstrip() for l in lines[i:i + 8]]
    par.append("SEED = 12345")  # added: the manual's example leaves the seed random
    files["example.par"] = par
    for name, ls in files.items():
        with open(os.path.join(OUT, name), "w") as f:
            f.write("\n".join(ls) + "\n")
        print(name, len(ls), "lines")
    # Table 1 (natural pairing; counts of meioses with k quadrivalents out of 10000)
    i = next(k for k, l in enumerate(lines) if l.startswith("Table 1."))
    toks = []
    for l in lines[i + 1:i + 200]:
        if l.startswith("Finally"):
            break
        toks += l.split()
    prefs = [float(x) for x in toks[toks.index("pairing") + 1:toks.index("ploidy")]]
    rest = [t for t in toks[toks.index("quadrivalents") + 1:]]
    table1 = {}
    k = 0
    for ploidy, rows in [(4, 2), (6, 2), (8, 3), (10, 3), (12, 4)]:
        for r in range(rows):
            if rest[k] != str(r):  # the vertically centred ploidy label
                assert int(rest[k]) == ploidy, (rest[k], ploidy)
                k += 1
            assert int(rest[k]) == r, (rest[k], r)
            table1.setdefault(str(ploidy), []).append([int(v) for v in rest[k + 1:k + 6]])
            k += 6
        assert rest[k] == "total" and rest[k + 1:k + 6] == ["10000"] * 5
        k += 6
    with open(os.path.join(HERE, "manual_table1.json"), "w") as f:
        json.dump({"pref_pairing": prefs, "meioses": 10000, "counts_by_ploidy": table1,
                   "source": "Manual/PedigreeSim_manual.docx, Table 1 (NATURALPAIRING = 1)"}, f, indent=1)
    print("table1", table1)


if __name__ == "__main__":
    sys.exit(main())
