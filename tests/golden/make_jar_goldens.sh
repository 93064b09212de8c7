#!/bin/bash
# Records golden files with the REFERENCE'S OWN JAR (needs a JVM: none ships in the build image, so this
# is run by a maintainer on any machine with Java >= 8 and the outputs are committed):
#
#     tests/golden/make_jar_goldens.sh [path/to/PedigreeSim.jar]
#
# For every case directory under tests/golden/jar_cases/ (inputs written below) it runs
#     java -jar PedigreeSim.jar case.par          (compiled/ReadMe_20210413.txt:1-12, Main.java:47-58)
# and keeps what Main.simulate writes (Main.java:358-426): <out>.ped, <out>.hsa, <out>.hsb,
# <out>_meioticconfigs.dat, <out>_founderalleles.dat, <out>_genotypes.dat, <out>_alleledose.dat,
# <out>_allAlleledose.dat  ->  tests/golden/jar/<case>/.
# tests/test_gpu_host.py::test_replay_of_jar_recorded_haplostructs then feeds the jar's .hsa/.hsb to the
# host program (replay mode, HAPLOSTRUCT=...) and requires its four .dat files to equal the jar's byte for
# byte: north_star's "replay mode consumes the reference's own recorded crossover and founder-allele output".
# tests/test_oracle_pins.py::test_oracle_java_mode_reproduces_the_jar requires the oracle's java.util.Random
# mode to reproduce the jar's .hsa/.hsb/_meioticconfigs.dat (draw-for-draw equality).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
JAR="${1:-$HERE/../../baseline/_ref/PedigreeSim.jar}"
command -v java >/dev/null || { echo "no java on PATH"; exit 1; }
[ -f "$JAR" ] || { echo "jar not found: $JAR (python -c 'import __graft_entry__ as g; g.build()' copies it from /root/reference)"; exit 1; }
python "$HERE/make_jar_cases.py" "$HERE/jar_cases"
for d in "$HERE"/jar_cases/*/; do
    name="$(basename "$d")"
    out="$HERE/jar/$name"
    rm -rf "$out"; mkdir -p "$out"
    cp "$d"/* "$out"/
    (cd "$out" && java -jar "$JAR" case.par > jar_stdout.txt)
    echo "recorded $name: $(ls "$out" | wc -l) files"
done
