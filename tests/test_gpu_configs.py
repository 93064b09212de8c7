"""Byte parity of the emission path (K0 + K2) against the CPU oracle at BASELINE.json's OWN
parameters, in the replay form north_star defines: the population the
CUDA library simulated is read back through the C ABI (psim_get_haplostructs), handed to the oracle
exactly as the reference's mode 2 reads `.hsa/.hsb` (Main.readHaploStructFiles, Main.java:1166-1291
-> orc_set_haplostructs), and the oracle's restatement of Main.writeGenotypesFile /
writeAlleledoseFile (Main.java:903-1022) over HaploStruct.getFounderAt (HaploStruct.java:107-142)
must reproduce every emitted byte.

configs[1] runs at FULL size through the exact call bench.py times (psim_emit with
PSIM_MEM_DEVICE targets, 128-byte padded pitch, `emit_into`); configs[2..4] run at their own model
parameters with the individual counts cut to what the oracle does in seconds on the box's cores.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import _cases
import _oracle

pytestmark = pytest.mark.gpu


def _replay_oracle(g, ploidy, lengths, centro, p0, p1, off, pos, codes=None, replicate=None):
    """The oracle in the reference's replay mode, holding the haplostructs the GPU context holds."""
    so, fo, po = g.get_haplostructs(replicate=replicate) if replicate is not None else g.get_haplostructs()
    o = _oracle.Oracle(ploidy)
    o.set_chromosomes(lengths, centro)
    o.set_pedigree(p0, p1)
    o.set_haplostructs(0, so, fo, po)
    o.set_map(off, pos)
    if codes is not None:
        o.set_allele_codes(codes)
    return o


def _oracle_tables(o, off, which, ind_first=0, ind_count=None, kinds=(_oracle.EMIT_FOUNDER, _oracle.EMIT_DOSE)):
    """Oracle tables of all loci, one chromosome per host thread (orc_emit only reads the population
    and ctypes releases the GIL)."""
    C = len(off) - 1
    n_thr = max(1, min(C, len(os.sched_getaffinity(0))))
    out = {}
    with ThreadPoolExecutor(n_thr) as ex:
        for kind in kinds:
            parts = list(ex.map(lambda c: o.emit(kind, which=which, ind_first=ind_first, ind_count=ind_count,
                                                 loc_first=int(off[c]), loc_count=int(off[c + 1] - off[c])), range(C)))
            out[kind] = np.concatenate(parts, axis=0)
    return out


def test_configs1_full_size_through_the_bench_call_matches_the_oracle():
    """configs[1]: autotetraploid F1 of 10 000, NATURALPAIRING=0, quadrivalents 0.5, prefPairing 0.3,
    12 chromosomes x 2 000 loci, biallelic .gen; founder table + MAX_ALLELE dose, device targets."""
    import torch
    import bench
    import pedigreesim_b200 as ps
    w = bench.workload()
    P, L, N = w["P"], w["L"], w["N"]
    g = ps.PsimContext(P, seed=bench.SEED, replicate_first=3, **w["model"])
    g.set_chromosomes(w["lengths"], w["centro"], w["pref"], w["fq"])
    g.set_pedigree(w["p0"], w["p1"])
    g.set_map(w["off"], w["pos"])
    g.set_allele_codes(w["code"])
    g.simulate()
    fp = (N * P + 127) // 128 * 128
    dp = (N + 127) // 128 * 128
    d_founder = torch.full((L, fp), 0xEE, dtype=torch.uint8, device="cuda")
    d_dose = torch.full((L, dp), 0xEE, dtype=torch.uint8, device="cuda")
    tgt = ps.EmitTargets()
    tgt.founder, tgt.founder_pitch, tgt.founder_bytes = d_founder.data_ptr(), fp, 1
    tgt.dose, tgt.dose_pitch, tgt.dose_which = d_dose.data_ptr(), dp, ps.DOSE_MAX_ALLELE
    tgt.memory = ps.MEM_DEVICE
    g.emit_into(tgt, 0, N, 0, L)
    torch.cuda.synchronize()
    F = d_founder.cpu().numpy()
    D = d_dose.cpu().numpy()
    assert (F[:, N * P:] == 0xEE).all() and (D[:, N:] == 0xEE).all()      # the padding is not touched
    o = _replay_oracle(g, P, w["lengths"], w["centro"], w["p0"], w["p1"], w["off"], w["pos"], w["code"])
    ref = _oracle_tables(o, w["off"], _oracle.MAX_ALLELE)
    assert np.array_equal(F[:, :N * P], ref[_oracle.EMIT_FOUNDER])
    assert np.array_equal(D[:, :N], ref[_oracle.EMIT_DOSE])
    st = g.stats()
    assert st["emit_kernel_launches"] == 1                                  # one launch of the row-image kernel


def test_configs2_hexaploid_natural_pairing_matches_the_oracle():
    """configs[2]: hexaploid F1, NATURALPAIRING=1, prefPairing 0, 21 chromosomes of 100 cM (centromere
    50), no .gen (founder alleles + dose of founder allele 0); 2 000 offspring x 2 000 loci / chr."""
    import pedigreesim_b200 as ps
    P, C, n, loci = 6, 21, 2000, 2000
    lengths, centro = [1.0] * C, [0.5] * C
    p0, p1 = _cases.std_pedigree("F1", n)
    off, pos = _cases.uniform_map(lengths, [loci] * C)
    g = ps.PsimContext(P, seed=12345, natural_pairing=True)
    g.set_chromosomes(lengths, centro, [0.0] * C, [0.0] * C)
    g.set_pedigree(p0, p1)
    g.simulate()
    g.set_map(off, pos)
    quad, _ = g.get_meiotic_configs(first=2)
    assert quad.max() == 1 and 0.87 < (quad == 1).mean() < 0.93            # manual Table 1, ploidy 6, pref 0: 8 990 of 10 000
    got = g.emit(founder=True, dose=True, dose_which=0)
    o = _replay_oracle(g, P, lengths, centro, p0, p1, off, pos)
    ref = _oracle_tables(o, off, 0)
    assert np.array_equal(got["founder"], ref[_oracle.EMIT_FOUNDER])
    assert np.array_equal(got["dose"], ref[_oracle.EMIT_DOSE])


def test_configs3_ril_chain_f8_matches_the_oracle():
    """configs[3]: diploid RIL chain P1 x P2 -> F1 -> F2 ... F8 by selfing, 10 chromosomes of 150 cM;
    2 000 lines x 5 000 loci / chr, the F8 generation emitted."""
    import pedigreesim_b200 as ps
    P, C, lines, loci = 2, 10, 2000, 5000
    lengths, centro = [1.5] * C, [0.75] * C
    p0, p1 = _cases.ril_pedigree(lines, 8)
    N = len(p0)
    assert N == 3 + 7 * lines
    off, pos = _cases.uniform_map(lengths, [loci] * C)
    g = ps.PsimContext(P, seed=12345)
    g.set_chromosomes(lengths, centro)
    g.set_pedigree(p0, p1)
    g.simulate()
    g.set_map(off, pos)
    first = N - lines
    got = g.emit(founder=True, dose=True, dose_which=0, ind_first=first, ind_count=lines)
    o = _replay_oracle(g, P, lengths, centro, p0, p1, off, pos)
    ref = _oracle_tables(o, off, 0, ind_first=first, ind_count=lines)
    assert np.array_equal(got["founder"], ref[_oracle.EMIT_FOUNDER])
    assert np.array_equal(got["dose"], ref[_oracle.EMIT_DOSE])
    het = (got["founder"].reshape(-1, lines, 2)[:, :, 0] != got["founder"].reshape(-1, lines, 2)[:, :, 1]).mean()
    assert 0.002 < het < 0.02                                                # F8: (1/2)^7 = 0.0078 heterozygous


def test_configs4_replicate_sweep_matches_the_oracle():
    """configs[4]: independent replicate tetraploid pedigrees side by side in one context (F1 of 200,
    natural pairing, prefPairing 0.25, 5 chromosomes x 100 loci), 64 replicates; every replicate's
    column block must be what the oracle emits for that replicate's population."""
    import pedigreesim_b200 as ps
    P, C, n, loci, R = 4, 5, 200, 100, 64
    lengths, centro = [1.0] * C, [0.5] * C
    p0, p1 = _cases.std_pedigree("F1", n)
    N = len(p0)
    off, pos = _cases.uniform_map(lengths, [loci] * C)
    g = ps.PsimContext(P, seed=12345, natural_pairing=True, replicate_first=1000, replicate_count=R)
    g.set_chromosomes(lengths, centro, [0.25] * C, [0.0] * C)
    g.set_pedigree(p0, p1)
    g.simulate()
    g.set_map(off, pos)
    got = g.emit(founder=True, dose=True, dose_which=0)
    assert got["founder"].shape == (C * loci, R * N * P)
    Fr = got["founder"].reshape(C * loci, R, N * P)
    Dr = got["dose"].reshape(C * loci, R, N)
    for r in range(R):
        o = _replay_oracle(g, P, lengths, centro, p0, p1, off, pos, replicate=1000 + r)
        assert np.array_equal(Fr[:, r], o.emit(_oracle.EMIT_FOUNDER)), r
        assert np.array_equal(Dr[:, r], o.emit(_oracle.EMIT_DOSE, which=0)), r
    assert len({Fr[:, r].tobytes() for r in range(R)}) == R                # the replicates differ
