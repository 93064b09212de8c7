"""ctypes binding of the CPU oracle (oracle/libpsim_oracle.so) — test infrastructure only.

The method names mirror pedigreesim_b200.PsimContext so parity tests drive both identically.
Arrays use the .hsa order [individual][chromosome][homolog].
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libpsim_oracle.so")
CLI = os.path.join(ORACLE_DIR, "psim_oracle_cli")

RNG_JAVA, RNG_PHILOX = 0, 1
MIN_ALLELE, MAX_ALLELE = -1, -2
EMIT_FOUNDER, EMIT_ALLELE, EMIT_DOSE, EMIT_FOUNDER_U16 = 0, 1, 2, 3


def build():
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("psim_oracle.cpp", "popio.hpp", "meiosis.hpp", "rng.hpp", "cli.cpp")]
    stale = not (os.path.exists(LIB) and os.path.exists(CLI)) or any(
        os.path.getmtime(s) > min(os.path.getmtime(LIB), os.path.getmtime(CLI)) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp, i32, i64, u32, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_double
        L.orc_create.restype = vp
        L.orc_create.argtypes = [i32, i32, i32, i32, dbl, dbl, i32, i64, u32]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_unreduced_gametes.argtypes = [vp, i32]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [vp]
        L.orc_set_chromosomes.argtypes = [vp, i32, vp, vp, vp, vp]
        L.orc_set_pedigree.argtypes = [vp, i64, vp, vp]
        L.orc_set_haplostructs.argtypes = [vp, i64, i64, vp, vp, vp]
        L.orc_simulate.argtypes = [vp, i32]
        L.orc_segment_count.restype = i64
        L.orc_segment_count.argtypes = [vp, i64, i64]
        L.orc_get_haplostructs.argtypes = [vp, i64, i64, vp, vp, vp]
        L.orc_get_meiotic_configs.argtypes = [vp, i64, i64, vp, vp]
        L.orc_set_map.argtypes = [vp, vp, vp]
        L.orc_set_allele_codes.argtypes = [vp, i32, vp]
        L.orc_emit.argtypes = [vp, i32, i32, i64, i64, i64, i64, vp, i64]
        L.orc_get_stats.argtypes = [vp, vp]
        L.orc_run_parameter_file.argtypes = [C.c_char_p, i32, C.c_char_p, i64]
        L.orc_java_next_int32.restype = i32
        L.orc_java_next_int32.argtypes = [i64, i32]
        L.orc_java_next_int.restype = i32
        L.orc_java_next_int.argtypes = [i64, i32, i32]
        L.orc_java_next_double.restype = dbl
        L.orc_java_next_double.argtypes = [i64, i32]
        L.orc_philox_block.argtypes = [vp, vp, vp]
        L.orc_philox_stream.argtypes = [u32, u32, u32, u32, u32, u32, i32, vp]
        L.orc_calc_recomb_dist.restype = dbl
        L.orc_calc_recomb_dist.argtypes = [dbl]
        L.orc_java_double_to_string.argtypes = [dbl, C.c_char_p, i32]
        L.orc_ran_gamma.restype = dbl
        L.orc_ran_gamma.argtypes = [i64, dbl, dbl, i32]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """CPU restatement of the reference's meiosis + emission path."""

    def __init__(self, ploidy, kosambi=False, natural_pairing=True, allow_no_recomb=True,
                 parallel_multivalents=0.0, paired_centromeres=0.0, rng=RNG_PHILOX, seed=12345, replicate=0,
                 unreduced_gametes=0):
        self.L = lib()
        self.P = ploidy
        self.h = self.L.orc_create(ploidy, int(kosambi), int(natural_pairing), int(allow_no_recomb),
                                   parallel_multivalents, paired_centromeres, rng, seed, replicate)
        if unreduced_gametes:
            self._chk(self.L.orc_set_unreduced_gametes(self.h, unreduced_gametes))
        self.C = 0
        self.N = 0

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    __del__ = close

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.orc_last_error(self.h).decode())

    def set_chromosomes(self, length_M, centromere_M, pref_pairing=None, frac_quad=None):
        ln = np.ascontiguousarray(length_M, dtype=np.float64)
        ce = np.ascontiguousarray(centromere_M, dtype=np.float64)
        pp = np.ascontiguousarray(pref_pairing if pref_pairing is not None else np.zeros_like(ln), dtype=np.float64)
        fq = np.ascontiguousarray(frac_quad if frac_quad is not None else np.zeros_like(ln), dtype=np.float64)
        self.C = len(ln)
        self._chk(self.L.orc_set_chromosomes(self.h, self.C, _p(ln), _p(ce), _p(pp), _p(fq)))

    def set_pedigree(self, parent0, parent1):
        p0 = np.ascontiguousarray(parent0, dtype=np.int32)
        p1 = np.ascontiguousarray(parent1, dtype=np.int32)
        self.N = len(p0)
        self._chk(self.L.orc_set_pedigree(self.h, self.N, _p(p0), _p(p1)))

    def set_haplostructs(self, first, seg_off, founder, pos):
        so = np.ascontiguousarray(seg_off, dtype=np.int64)
        fo = np.ascontiguousarray(founder, dtype=np.int32)
        po = np.ascontiguousarray(pos, dtype=np.float64)
        count = (len(so) - 1) // (self.C * self.P)
        self._chk(self.L.orc_set_haplostructs(self.h, first, count, _p(so), _p(fo), _p(po)))

    def simulate(self, maxfounderallele=-1):
        self._chk(self.L.orc_simulate(self.h, maxfounderallele))

    def get_haplostructs(self, first=0, count=None):
        count = self.N - first if count is None else count
        n = self.L.orc_segment_count(self.h, first, count)
        so = np.zeros(count * self.C * self.P + 1, dtype=np.int64)
        fo = np.zeros(n, dtype=np.int32)
        po = np.zeros(n, dtype=np.float64)
        self._chk(self.L.orc_get_haplostructs(self.h, first, count, _p(so), _p(fo), _p(po)))
        return so, fo, po

    def get_meiotic_configs(self, first=0, count=None):
        count = self.N - first if count is None else count
        quad = np.zeros((count, self.C, 2), dtype=np.int32)
        seq = np.zeros((count, self.C, 2, self.P), dtype=np.int32)
        self._chk(self.L.orc_get_meiotic_configs(self.h, first, count, _p(quad), _p(seq)))
        return quad, seq

    def set_map(self, locus_off, pos_M):
        lo = np.ascontiguousarray(locus_off, dtype=np.int64)
        po = np.ascontiguousarray(pos_M, dtype=np.float64)
        self.Ltot = int(lo[-1])
        self._chk(self.L.orc_set_map(self.h, _p(lo), _p(po)))

    def set_allele_codes(self, code):
        code = np.ascontiguousarray(code, dtype=np.uint8)
        self.A = code.shape[1]
        self._chk(self.L.orc_set_allele_codes(self.h, self.A, _p(code)))

    def emit(self, kind, which=0, ind_first=0, ind_count=None, loc_first=0, loc_count=None):
        ind_count = self.N - ind_first if ind_count is None else ind_count
        loc_count = self.Ltot - loc_first if loc_count is None else loc_count
        cols = ind_count if kind == EMIT_DOSE else ind_count * self.P
        dt = np.uint16 if kind == EMIT_FOUNDER_U16 else np.uint8
        out = np.zeros((loc_count, cols), dtype=dt)
        self._chk(self.L.orc_emit(self.h, kind, which, ind_first, ind_count, loc_first, loc_count, _p(out),
                                  cols * out.itemsize))
        return out

    def stats(self):
        out = np.zeros(11, dtype=np.int64)
        self.L.orc_get_stats(self.h, _p(out))
        return out


def run_parameter_file(path, rng=RNG_JAVA, cwd=None):
    """Whole-program restatement; relative file names resolve against cwd like the jar."""
    args = [CLI, path] + (["--philox"] if rng == RNG_PHILOX else [])
    build()
    return subprocess.run(args, cwd=cwd, check=True, capture_output=True, text=True).stdout
