"""ctypes binding of libpsim.so (include/psim.h).

There is deliberately no fallback: if the CUDA library is missing or no device is present the
calls raise. Nothing here imports the CPU oracle.
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libpsim.so")
if os.environ.get("PSIM_TUNING_LIB"):   # experiments under tools/ only: a differently built library (profiles/README.md)
    LIB_PATH = os.environ["PSIM_TUNING_LIB"]

PSIM_OK, PSIM_EINVAL, PSIM_ECUDA, PSIM_ECAPACITY, PSIM_ESTATE = 0, 1, 2, 3, 4
DOSE_MIN_ALLELE, DOSE_MAX_ALLELE = -1, -2
MEM_HOST, MEM_DEVICE = 0, 1
EMIT_ASYNC, EMIT_PACKED4 = 0x1, 0x2
PATH_AUTO, PATH_ROWS, PATH_TILES, PATH_GENERIC = 0, 1, 2, 3
PARTITION_BLOCK, PARTITION_LINEAGE = 0, 1
UNIQUE_ID_BYTES = 128


def emit_path(p):
    return p << 8


class PsimError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("ploidy", C.c_int32), ("chiasma_interference", C.c_int32), ("natural_pairing", C.c_int32),
        ("allow_no_recomb", C.c_int32), ("parallel_multivalents", C.c_double), ("paired_centromeres", C.c_double),
        ("seed", C.c_int64), ("replicate_first", C.c_uint32), ("replicate_count", C.c_uint32),
        ("device", C.c_int32), ("unreduced_gametes", C.c_int32),
    ]


class EmitTargets(C.Structure):
    _fields_ = [
        ("founder", C.c_void_p), ("founder_pitch", C.c_int64), ("founder_bytes", C.c_int32), ("memory", C.c_int32),
        ("allele", C.c_void_p), ("allele_pitch", C.c_int64),
        ("dose", C.c_void_p), ("dose_pitch", C.c_int64), ("dose_which", C.c_int32), ("flags", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_meioses", C.c_int64), ("n_units", C.c_int64), ("n_segments", C.c_int64), ("kernel_launches", C.c_int64),
        ("ms_simulate", C.c_double), ("ms_emit_kernel", C.c_double), ("ms_emit_total", C.c_double),
        ("emit_kernel_launches", C.c_int64), ("ms_emit_plan", C.c_double),
    ]


class TextTargets(C.Structure):
    _fields_ = [("founder", C.c_void_p), ("founder_capacity", C.c_int64), ("founder_bytes", C.c_int64),
                ("allele", C.c_void_p), ("allele_capacity", C.c_int64), ("allele_bytes", C.c_int64),
                ("dose", C.c_void_p), ("dose_capacity", C.c_int64), ("dose_bytes", C.c_int64),
                ("dose_which", C.c_int32), ("reserved", C.c_int32)]


class GameteStats(C.Structure):
    _fields_ = [("n_gametes", C.c_int64), ("allele_count", C.c_void_p), ("dose_count", C.c_void_p),
                ("homozygous_count", C.c_void_p), ("recomb_count", C.c_void_p), ("segments_hist", C.c_void_p),
                ("quadrivalent_hist", C.c_void_p)]


class ShardStats(C.Structure):
    _fields_ = [("collectives", C.c_int64), ("bytes_sent", C.c_int64), ("segments_received", C.c_int64), ("ms_total", C.c_double)]


class DeviceHS(C.Structure):
    _fields_ = [("n_homologs", C.c_int64), ("n_segments", C.c_int64), ("seg_off", C.c_void_p),
                ("founder", C.c_void_p), ("pos", C.c_void_p)]


EXPORTS = [
    "psim_abi_version", "psim_create", "psim_destroy", "psim_last_error", "psim_set_chromosomes", "psim_set_pedigree",
    "psim_set_haplostructs", "psim_simulate", "psim_segment_count", "psim_get_haplostructs", "psim_get_meiotic_configs",
    "psim_set_map", "psim_set_allele_codes", "psim_emit", "psim_all_dose_rows", "psim_emit_all_dose", "psim_get_stats",
    "psim_export_haplostructs_device", "psim_import_haplostructs_device", "psim_reset", "psim_set_stream", "psim_simulate_subset",
    "psim_alloc_host", "psim_free_host", "psim_set_locus_names", "psim_set_allele_names", "psim_emit_text",
    "psim_gamete_stats_run", "psim_emit_all_dose_text", "psim_simulate_async", "psim_wait",
    "psim_shard_unique_id", "psim_shard_init", "psim_shard_group_create", "psim_shard_group_destroy", "psim_shard_init_local",
    "psim_simulate_sharded", "psim_shard_owner", "psim_shard_gather", "psim_shard_get_stats", "psim_shard_plan",
]

_lib = None


def load():
    """Load libpsim.so. Raises if it has not been built (python -m pedigreesim_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PsimError(PSIM_ESTATE, "libpsim.so is not built: run `python -m pedigreesim_b200.build` "
                                     "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
    L.psim_abi_version.restype = i32
    L.psim_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.psim_destroy.argtypes = [vp]
    L.psim_destroy.restype = None
    L.psim_last_error.argtypes = [vp]
    L.psim_last_error.restype = C.c_char_p
    L.psim_set_chromosomes.argtypes = [vp, i32, vp, vp, vp, vp]
    L.psim_set_pedigree.argtypes = [vp, i64, vp, vp]
    L.psim_set_haplostructs.argtypes = [vp, u32, i64, i64, vp, vp, vp]
    L.psim_simulate.argtypes = [vp, i32]
    L.psim_simulate_subset.argtypes = [vp, i32, vp, i64]
    L.psim_simulate_async.argtypes = [vp, i32]
    L.psim_wait.argtypes = [vp]
    L.psim_segment_count.argtypes = [vp, u32, i64, i64, C.POINTER(i64)]
    L.psim_get_haplostructs.argtypes = [vp, u32, i64, i64, vp, vp, vp]
    L.psim_get_meiotic_configs.argtypes = [vp, u32, i64, i64, vp, vp]
    L.psim_set_map.argtypes = [vp, vp, vp]
    L.psim_set_allele_codes.argtypes = [vp, i32, vp]
    L.psim_emit.argtypes = [vp, i64, i64, i64, i64, C.POINTER(EmitTargets)]
    L.psim_all_dose_rows.argtypes = [vp, i64, i64, vp]
    L.psim_emit_all_dose.argtypes = [vp, i64, i64, i64, i64, vp, i64, i32]
    L.psim_reset.argtypes = [vp, i64, u32]
    L.psim_set_stream.argtypes = [vp, C.c_uint64]
    L.psim_set_locus_names.argtypes = [vp, vp, vp]
    L.psim_set_allele_names.argtypes = [vp, vp, vp, vp]
    L.psim_emit_text.argtypes = [vp, i64, i64, i64, i64, C.POINTER(TextTargets)]
    L.psim_gamete_stats_run.argtypes = [vp, i64, i64, i32, C.POINTER(GameteStats)]
    L.psim_emit_all_dose_text.argtypes = [vp, i64, i64, i64, i64, vp, i64, C.POINTER(i64)]
    L.psim_alloc_host.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.psim_free_host.argtypes = [vp, vp]
    L.psim_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.psim_export_haplostructs_device.argtypes = [vp, i64, i64, C.POINTER(DeviceHS)]
    L.psim_import_haplostructs_device.argtypes = [vp, i64, i64, C.POINTER(DeviceHS)]
    L.psim_shard_unique_id.argtypes = [vp]
    L.psim_shard_init.argtypes = [vp, i32, i32, vp]
    L.psim_shard_group_create.argtypes = [i32, C.POINTER(vp)]
    L.psim_shard_group_destroy.argtypes = [vp]
    L.psim_shard_group_destroy.restype = None
    L.psim_shard_init_local.argtypes = [vp, i32, vp]
    L.psim_simulate_sharded.argtypes = [vp, i32, i32]
    L.psim_shard_owner.argtypes = [vp, vp]
    L.psim_shard_gather.argtypes = [vp, vp, i64]
    L.psim_shard_get_stats.argtypes = [vp, C.POINTER(ShardStats)]
    L.psim_shard_plan.argtypes = [i64, vp, vp, vp, i32, i32, vp, vp, vp]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class PsimContext:
    """One libpsim context = the device-resident population(s) of one host thread."""

    def __init__(self, ploidy, kosambi=False, natural_pairing=True, allow_no_recomb=True, parallel_multivalents=0.0,
                 paired_centromeres=0.0, seed=12345, replicate_first=0, replicate_count=1, device=0, unreduced_gametes=0):
        self.L = load()
        self.P = ploidy
        self.unreduced = unreduced_gametes   # 2NGAMETES: 0, 1 = FDR, 2 = SDR
        self.R = replicate_count
        self.replicate_first = replicate_first
        cfg = Config(ploidy, int(kosambi), int(natural_pairing), int(allow_no_recomb), parallel_multivalents,
                     paired_centromeres, seed, replicate_first, replicate_count, device, unreduced_gametes)
        h = C.c_void_p()
        rc = self.L.psim_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise PsimError(rc, self.L.psim_last_error(None).decode())
        self.h = h
        self.C = 0
        self.N = 0
        self.Ltot = 0
        self.A = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.psim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc != 0:
            raise PsimError(rc, self.L.psim_last_error(self.h).decode())

    # --- inputs
    def set_chromosomes(self, length_M, centromere_M, pref_pairing=None, frac_quad=None):
        ln = np.ascontiguousarray(length_M, dtype=np.float64)
        ce = np.ascontiguousarray(centromere_M, dtype=np.float64)
        pp = None if pref_pairing is None else np.ascontiguousarray(pref_pairing, dtype=np.float64)
        fq = None if frac_quad is None else np.ascontiguousarray(frac_quad, dtype=np.float64)
        self.C = len(ln)
        self._chk(self.L.psim_set_chromosomes(self.h, self.C, _p(ln), _p(ce), _p(pp), _p(fq)))

    def set_pedigree(self, parent0, parent1):
        p0 = np.ascontiguousarray(parent0, dtype=np.int32)
        p1 = np.ascontiguousarray(parent1, dtype=np.int32)
        self.N = len(p0)
        self._chk(self.L.psim_set_pedigree(self.h, self.N, _p(p0), _p(p1)))

    def set_haplostructs(self, first, seg_off, founder, pos, replicate=None):
        so = np.ascontiguousarray(seg_off, dtype=np.int64)
        fo = np.ascontiguousarray(founder, dtype=np.int32)
        po = np.ascontiguousarray(pos, dtype=np.float64)
        count = (len(so) - 1) // (self.C * self.P)
        rep = self.replicate_first if replicate is None else replicate
        self._chk(self.L.psim_set_haplostructs(self.h, rep, first, count, _p(so), _p(fo), _p(po)))

    def set_map(self, locus_off, pos_M):
        lo = np.ascontiguousarray(locus_off, dtype=np.int64)
        po = np.ascontiguousarray(pos_M, dtype=np.float64)
        self._chk(self.L.psim_set_map(self.h, _p(lo), _p(po)))
        self.Ltot = int(lo[-1])
        self.A = 0

    def set_allele_codes(self, code):
        if code is None:
            self._chk(self.L.psim_set_allele_codes(self.h, 0, None))
            self.A = 0
            return
        code = np.ascontiguousarray(code, dtype=np.uint8)
        self.A = code.shape[1]
        self._chk(self.L.psim_set_allele_codes(self.h, self.A, _p(code)))

    # --- seam A
    def simulate(self, maxfounderallele=-1):
        self._chk(self.L.psim_simulate(self.h, maxfounderallele))

    def simulate_async(self, maxfounderallele=-1):
        """Enqueue the simulation and return; wait() (or the next call that needs the population) completes it."""
        self._chk(self.L.psim_simulate_async(self.h, maxfounderallele))

    def wait(self):
        self._chk(self.L.psim_wait(self.h))

    def simulate_subset(self, individuals, maxfounderallele=-1):
        ind = np.ascontiguousarray(individuals, dtype=np.int32)
        self._chk(self.L.psim_simulate_subset(self.h, maxfounderallele, _p(ind), len(ind)))

    def get_haplostructs(self, first=0, count=None, replicate=None):
        count = self.N - first if count is None else count
        rep = self.replicate_first if replicate is None else replicate
        n = C.c_int64()
        self._chk(self.L.psim_segment_count(self.h, rep, first, count, C.byref(n)))
        so = np.zeros(count * self.C * self.P + 1, dtype=np.int64)
        fo = np.zeros(n.value, dtype=np.int32)
        po = np.zeros(n.value, dtype=np.float64)
        self._chk(self.L.psim_get_haplostructs(self.h, rep, first, count, _p(so), _p(fo), _p(po)))
        return so, fo, po

    def get_meiotic_configs(self, first=0, count=None, replicate=None):
        count = self.N - first if count is None else count
        rep = self.replicate_first if replicate is None else replicate
        quad = np.zeros((count, self.C, 2), dtype=np.int32)
        seq = np.zeros((count, self.C, 2, self.P), dtype=np.int32)
        self._chk(self.L.psim_get_meiotic_configs(self.h, rep, first, count, _p(quad), _p(seq)))
        return quad, seq

    # --- seam B
    def emit(self, founder=False, allele=False, dose=False, dose_which=0, ind_first=0, ind_count=None, loc_first=0,
             loc_count=None, founder_bytes=1, packed=False, path=PATH_AUTO):
        """Emit into fresh host arrays; returns a dict of the requested tables ([loci][columns]).
        packed: two 4-bit cells per byte (EMIT_PACKED4), rows of ceil(columns / 2) bytes."""
        ind_count = self.N - ind_first if ind_count is None else ind_count
        loc_count = self.Ltot - loc_first if loc_count is None else loc_count
        q = self.R * ind_count
        out = {}
        t = EmitTargets()
        t.memory = MEM_HOST
        t.dose_which = dose_which
        t.founder_bytes = founder_bytes
        t.flags = (EMIT_PACKED4 if packed else 0) | emit_path(path)
        if founder:
            cols = (q * self.P + 1) // 2 if packed else q * self.P
            out["founder"] = np.zeros((loc_count, cols), dtype=np.uint8 if founder_bytes == 1 else np.uint16)
            t.founder = out["founder"].ctypes.data
            t.founder_pitch = cols * founder_bytes
        if allele:
            out["allele"] = np.zeros((loc_count, q * self.P), dtype=np.uint8)
            t.allele = out["allele"].ctypes.data
            t.allele_pitch = q * self.P
        if dose:
            cols = (q + 1) // 2 if packed else q
            out["dose"] = np.zeros((loc_count, cols), dtype=np.uint8)
            t.dose = out["dose"].ctypes.data
            t.dose_pitch = cols
        self._chk(self.L.psim_emit(self.h, ind_first, ind_count, loc_first, loc_count, C.byref(t)))
        return out

    def emit_into(self, targets, ind_first, ind_count, loc_first, loc_count):
        """Emit into caller-provided memory described by an EmitTargets (host or device pointers)."""
        self._chk(self.L.psim_emit(self.h, ind_first, ind_count, loc_first, loc_count, C.byref(targets)))

    def all_dose_rows(self, loc_first=0, loc_count=None):
        loc_count = self.Ltot - loc_first if loc_count is None else loc_count
        ro = np.zeros(loc_count + 1, dtype=np.int64)
        self._chk(self.L.psim_all_dose_rows(self.h, loc_first, loc_count, _p(ro)))
        return ro

    def emit_all_dose(self, ind_first=0, ind_count=None, loc_first=0, loc_count=None):
        ind_count = self.N - ind_first if ind_count is None else ind_count
        loc_count = self.Ltot - loc_first if loc_count is None else loc_count
        ro = self.all_dose_rows(loc_first, loc_count)
        q = self.R * ind_count
        out = np.zeros((int(ro[-1]), q), dtype=np.uint8)
        self._chk(self.L.psim_emit_all_dose(self.h, ind_first, ind_count, loc_first, loc_count, _p(out), q, MEM_HOST))
        return ro, out

    @staticmethod
    def _csr(strings):
        b = [x.encode() if isinstance(x, str) else bytes(x) for x in strings]
        off = np.zeros(len(b) + 1, dtype=np.int64)
        np.cumsum([len(x) for x in b], out=off[1:])
        return off, np.frombuffer(b"".join(b) + b"\0", dtype=np.uint8).copy()

    def set_locus_names(self, names):
        off, chars = self._csr(names)
        self._chk(self.L.psim_set_locus_names(self.h, _p(off), _p(chars)))

    def set_allele_names(self, names_per_locus):
        """names_per_locus[l] = the locus' sorted unique allele names (name k <-> code k)."""
        first = np.zeros(len(names_per_locus) + 1, dtype=np.int64)
        np.cumsum([len(x) for x in names_per_locus], out=first[1:])
        off, chars = self._csr([n for x in names_per_locus for n in x])
        self._chk(self.L.psim_set_allele_names(self.h, _p(first), _p(off), _p(chars)))

    def emit_text(self, founder=False, allele=False, dose=False, dose_which=0, ind_first=0, ind_count=None, loc_first=0,
                  loc_count=None, capacity=None):
        """Text lines of the requested tables (bytes objects); capacity defaults to a safe upper bound."""
        ind_count = self.N - ind_first if ind_count is None else ind_count
        loc_count = self.Ltot - loc_first if loc_count is None else loc_count
        q = self.R * ind_count
        cap = capacity if capacity is not None else loc_count * (q * self.P * 8 + 4096)
        t = TextTargets()
        t.dose_which = dose_which
        bufs = {}
        for name, want in (("founder", founder), ("allele", allele), ("dose", dose)):
            if want:
                bufs[name] = np.zeros(cap, dtype=np.uint8)
                setattr(t, name, bufs[name].ctypes.data)
                setattr(t, name + "_capacity", cap)
        self._chk(self.L.psim_emit_text(self.h, ind_first, ind_count, loc_first, loc_count, C.byref(t)))
        return {k: bytes(v[:getattr(t, k + "_bytes")]) for k, v in bufs.items()}

    def emit_all_dose_text(self, ind_first=0, ind_count=None, loc_first=0, loc_count=None, capacity=None):
        ind_count = self.N - ind_first if ind_count is None else ind_count
        loc_count = self.Ltot - loc_first if loc_count is None else loc_count
        rows = int(self.all_dose_rows(loc_first, loc_count)[-1])
        cap = capacity if capacity is not None else rows * (self.R * ind_count * 3 + 4096)
        buf = np.zeros(cap, dtype=np.uint8)
        n = C.c_int64()
        self._chk(self.L.psim_emit_all_dose_text(self.h, ind_first, ind_count, loc_first, loc_count, _p(buf), cap, C.byref(n)))
        return bytes(buf[:n.value])

    def gamete_stats(self, chrom, n_loci, n_alleles, ind_first=0, ind_count=None, recomb=True):
        """TEST-mode counters of one chromosome over the gametes of the given individuals (dict of arrays)."""
        ind_count = self.N - ind_first if ind_count is None else ind_count
        nd = (self.P if self.unreduced else self.P // 2) + 1   # homologs per gamete + 1
        out = {"allele_count": np.zeros((n_alleles, n_loci), np.uint32), "dose_count": np.zeros((n_alleles, n_loci, nd), np.uint32),
               "homozygous_count": np.zeros(n_loci, np.uint32), "segments_hist": np.zeros(64, np.uint32),
               "quadrivalent_hist": np.zeros(self.P // 4 + 1, np.uint32)}
        if recomb:
            out["recomb_count"] = np.zeros((n_loci, n_loci), np.uint32)
        gs = GameteStats()
        for k, v in out.items():
            setattr(gs, k, v.ctypes.data)
        self._chk(self.L.psim_gamete_stats_run(self.h, ind_first, ind_count, chrom, C.byref(gs)))
        out["n_gametes"] = gs.n_gametes
        return out

    def reset(self, seed, replicate_first=0):
        self.replicate_first = replicate_first
        self._chk(self.L.psim_reset(self.h, seed, replicate_first))

    def set_stream(self, cuda_stream):
        self._chk(self.L.psim_set_stream(self.h, int(cuda_stream)))

    def stats(self):
        s = Stats()
        self._chk(self.L.psim_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def export_device(self, first, count):
        d = DeviceHS()
        self._chk(self.L.psim_export_haplostructs_device(self.h, first, count, C.byref(d)))
        return d

    def import_device(self, first, count, d):
        self._chk(self.L.psim_import_haplostructs_device(self.h, first, count, C.byref(d)))

    # ---- one pedigree over several GPUs (psim_simulate_sharded)
    def shard_init(self, rank, world, unique_id):
        """One process per GPU over NCCL; `unique_id` = shard_unique_id() of rank 0, handed to every rank."""
        uid = np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        assert len(uid) == UNIQUE_ID_BYTES
        self._chk(self.L.psim_shard_init(self.h, rank, world, _p(uid)))

    def shard_init_local(self, rank, group):
        """Several contexts of one process (ShardGroup), each driven by its own thread."""
        self._chk(self.L.psim_shard_init_local(self.h, rank, group.h))
        self._group = group   # (the group outlives the contexts that joined it)

    def simulate_sharded(self, maxfounderallele=-1, partition=PARTITION_BLOCK):
        self._chk(self.L.psim_simulate_sharded(self.h, maxfounderallele, partition))

    def shard_owner(self):
        out = np.empty(self.N, np.int32)
        self._chk(self.L.psim_shard_owner(self.h, _p(out)))
        return out

    def shard_gather(self, individuals):
        ind = np.ascontiguousarray(individuals, dtype=np.int32)
        self._chk(self.L.psim_shard_gather(self.h, _p(ind), len(ind)))

    def shard_stats(self):
        s = ShardStats()
        self._chk(self.L.psim_shard_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in ShardStats._fields_}


def shard_unique_id():
    """The NCCL unique id of a new communicator (rank 0 makes it, every rank passes it to shard_init)."""
    L = load()
    uid = np.zeros(UNIQUE_ID_BYTES, np.uint8)
    rc = L.psim_shard_unique_id(_p(uid))
    if rc != PSIM_OK:
        raise PsimError(rc, (L.psim_last_error(None) or b"").decode())
    return uid.tobytes()


class ShardGroup:
    """Rendezvous of the contexts of one process that share a pedigree (psim_shard_init_local)."""

    def __init__(self, world):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.psim_shard_group_create(world, C.byref(h))
        if rc != PSIM_OK:
            raise PsimError(rc, "psim_shard_group_create failed")
        self.h, self.world = h, world

    def close(self):
        if self.h:
            self.L.psim_shard_group_destroy(self.h)
            self.h = None


def shard_plan(parent0, parent1, world, partition, has_haplostruct=None):
    """(level, owner, share) of every individual: the partition psim_simulate_sharded uses (pure host)."""
    L = load()
    p0 = np.ascontiguousarray(parent0, dtype=np.int32)
    p1 = np.ascontiguousarray(parent1, dtype=np.int32)
    n = len(p0)
    has = None if has_haplostruct is None else np.ascontiguousarray(has_haplostruct, dtype=np.uint8)
    level, owner, share = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.uint8)
    rc = L.psim_shard_plan(n, _p(p0), _p(p1), None if has is None else _p(has), world, partition, _p(level), _p(owner), _p(share))
    if rc != PSIM_OK:
        raise PsimError(rc, "psim_shard_plan: invalid pedigree or arguments")
    return level, owner, share.astype(bool)
