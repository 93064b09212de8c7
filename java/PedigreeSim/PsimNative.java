/*
 * PsimNative — Panama FFM (java.lang.foreign, JDK 22+) binding of libpsim (include/psim.h).
 *
 * NOT COMPILED IN THE AUTHORING IMAGE (no javac / JVM there); it is the reference-side stub a
 * PedigreeSim maintainer adds next to Main.java. It lives in package PedigreeSim so that it can
 * read the package-private fields the two seams use (Individual.parconfig, Genotype.ChromConfig).
 *
 * Seam A: Main.simulateIndividuals(popdata, maxfounderallele)   (Main.java:1504-1520)
 * Seam B: Main.writeGenotypesFile / writeAlleledoseFile         (Main.java:903-1022)
 */
package PedigreeSim;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import java.util.ArrayList;

import static java.lang.foreign.ValueLayout.*;

final class PsimNative implements AutoCloseable {
    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(
            System.getProperty("psim.library", "libpsim.so"), Arena.global());

    private static MethodHandle fn(String name, FunctionDescriptor d) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(), d);
    }

    // struct psim_config (include/psim.h)
    static final StructLayout CONFIG = MemoryLayout.structLayout(
            JAVA_INT.withName("ploidy"), JAVA_INT.withName("chiasma_interference"),
            JAVA_INT.withName("natural_pairing"), JAVA_INT.withName("allow_no_recomb"),
            JAVA_DOUBLE.withName("parallel_multivalents"), JAVA_DOUBLE.withName("paired_centromeres"),
            JAVA_LONG.withName("seed"), JAVA_INT.withName("replicate_first"), JAVA_INT.withName("replicate_count"),
            JAVA_INT.withName("device"), JAVA_INT.withName("unreduced_gametes"));
    // struct psim_emit_targets
    static final StructLayout TARGETS = MemoryLayout.structLayout(
            ADDRESS.withName("founder"), JAVA_LONG.withName("founder_pitch"), JAVA_INT.withName("founder_bytes"),
            JAVA_INT.withName("memory"), ADDRESS.withName("allele"), JAVA_LONG.withName("allele_pitch"),
            ADDRESS.withName("dose"), JAVA_LONG.withName("dose_pitch"), JAVA_INT.withName("dose_which"),
            JAVA_INT.withName("reserved"));

    private static final MethodHandle CREATE = fn("psim_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle DESTROY = fn("psim_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle LAST_ERROR = fn("psim_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    private static final MethodHandle SET_CHROMOSOMES = fn("psim_set_chromosomes",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle SET_PEDIGREE = fn("psim_set_pedigree",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS));
    private static final MethodHandle SET_HAPLOSTRUCTS = fn("psim_set_haplostructs",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle SIMULATE = fn("psim_simulate", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    private static final MethodHandle SEGMENT_COUNT = fn("psim_segment_count",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS));
    private static final MethodHandle GET_HAPLOSTRUCTS = fn("psim_get_haplostructs",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle GET_CONFIGS = fn("psim_get_meiotic_configs",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS));
    private static final MethodHandle SET_MAP = fn("psim_set_map", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle SET_CODES = fn("psim_set_allele_codes",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    // finished text lines of the table files: int psim_set_locus_names(ctx, int64_t* off, char* chars),
    // int psim_set_allele_names(ctx, int64_t* first, int64_t* off, char* chars), int psim_emit_text(ctx, i0, ni, l0, nl, psim_text_targets*)
    private static final MethodHandle SET_LOCUS_NAMES = fn("psim_set_locus_names", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle SET_ALLELE_NAMES = fn("psim_set_allele_names", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle EMIT_TEXT = fn("psim_emit_text",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_LONG, JAVA_LONG, ADDRESS));
    // TEST-mode counters on the device: int psim_gamete_stats_run(ctx, ind_first, ind_count, chrom, psim_gamete_stats*)
    private static final MethodHandle GAMETE_STATS = fn("psim_gamete_stats_run",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_INT, ADDRESS));
    private static final MethodHandle EMIT_ALL_DOSE_TEXT = fn("psim_emit_all_dose_text",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_LONG, JAVA_LONG, ADDRESS, JAVA_LONG, ADDRESS));
    // page-locked target tables: int psim_alloc_host(psim_ctx*, size_t, void**), int psim_free_host(psim_ctx*, void*)
    private static final MethodHandle ALLOC_HOST = fn("psim_alloc_host", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
    private static final MethodHandle FREE_HOST = fn("psim_free_host", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle EMIT = fn("psim_emit",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_LONG, JAVA_LONG, ADDRESS));

    // several GPUs in one JVM: one PsimNative (context) and one Java thread per GPU (include/psim.h "psim_simulate_sharded")
    private static final MethodHandle SHARD_GROUP_CREATE = fn("psim_shard_group_create", FunctionDescriptor.of(JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle SHARD_GROUP_DESTROY = fn("psim_shard_group_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle SHARD_INIT_LOCAL = fn("psim_shard_init_local", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    private static final MethodHandle SIMULATE_SHARDED = fn("psim_simulate_sharded", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT));
    private static final MethodHandle SHARD_OWNER = fn("psim_shard_owner", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle SHARD_GATHER = fn("psim_shard_gather", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
    // a host thread that keeps several contexts busy (psim_simulate_async returns once the work is enqueued)
    private static final MethodHandle SIMULATE_ASYNC = fn("psim_simulate_async", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    private static final MethodHandle WAIT = fn("psim_wait", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final int PARTITION_BLOCK = 0, PARTITION_LINEAGE = 1;

    private final Arena arena = Arena.ofShared();   // (a sharded run touches the context from its own thread)
    private final MemorySegment ctx;
    private final PopulationData popdata;

    PsimNative(PopulationData popdata, int device) throws Exception {
        this.popdata = popdata;
        MemorySegment cfg = arena.allocate(CONFIG);
        cfg.set(JAVA_INT, 0, popdata.ploidy);
        cfg.set(JAVA_INT, 4, popdata.chiasmaInterference ? 1 : 0);
        cfg.set(JAVA_INT, 8, popdata.naturalPairing ? 1 : 0);
        cfg.set(JAVA_INT, 12, popdata.allowNoRecomb ? 1 : 0);
        cfg.set(JAVA_DOUBLE, 16, popdata.parallelMultivalents);
        cfg.set(JAVA_DOUBLE, 24, popdata.pairedCentromeres);
        cfg.set(JAVA_LONG, 32, popdata.randomSeed);
        cfg.set(JAVA_INT, 40, 0);   // replicate_first
        cfg.set(JAVA_INT, 44, 1);   // replicate_count
        cfg.set(JAVA_INT, 48, device);
        MemorySegment out = arena.allocate(ADDRESS);
        int rc = (int) invoke(CREATE, cfg, out);
        if (rc != 0) throw new Exception(errorText(MemorySegment.NULL));
        ctx = out.get(ADDRESS, 0);
        // chromosome table: Morgan, exactly what Main.readChromosomeFile stored (Main.java:619-630)
        int C = popdata.chromCount();
        double[] len = new double[C], cen = new double[C], pref = new double[C], fq = new double[C];
        for (int c = 0; c < C; c++) {
            Chromosome ch = popdata.getChrom(c);
            len[c] = ch.getTailPos();          // head is 0 unless the centromere lies outside
            cen[c] = ch.getCentromerePos();
            if (ch instanceof TetraploidChromosome t) { pref[c] = t.getPrefPairingProb(); fq[c] = t.getFracQuadrivalents(); }
        }
        check((int) invoke(SET_CHROMOSOMES, ctx, C, arena.allocateFrom(JAVA_DOUBLE, len), arena.allocateFrom(JAVA_DOUBLE, cen),
                arena.allocateFrom(JAVA_DOUBLE, pref), arena.allocateFrom(JAVA_DOUBLE, fq)));
    }

    private static Object invoke(MethodHandle h, Object... a) throws Exception {
        try { return h.invokeWithArguments(a); } catch (Throwable t) { throw new Exception(t); }
    }

    private String errorText(MemorySegment c) throws Exception {
        MemorySegment p = (MemorySegment) invoke(LAST_ERROR, c);
        return p.reinterpret(4096).getString(0);
    }

    /** The reference throws Exception(message), prints it and returns (Main.java:106-109,446-448). */
    private void check(int rc) throws Exception {
        if (rc != 0) throw new Exception(errorText(ctx));
    }

    /** Seam A — drop-in body of Main.simulateIndividuals (Main.java:1504-1520). */
    void simulateIndividuals(int maxfounderallele) throws Exception {
        ArrayList<Individual> inds = popdata.getIndividual();   // parents-first (PopulationData.sortPedigree)
        int n = inds.size(), C = popdata.chromCount(), P = popdata.ploidy;
        int[] p0 = new int[n], p1 = new int[n];
        for (int i = 0; i < n; i++) {
            Individual ind = inds.get(i);
            p0[i] = ind.isFounder() ? -1 : inds.indexOf(ind.getParents()[0]);
            p1[i] = ind.isFounder() ? -1 : inds.indexOf(ind.getParents()[1]);
        }
        check((int) invoke(SET_PEDIGREE, ctx, (long) n, arena.allocateFrom(JAVA_INT, p0), arena.allocateFrom(JAVA_INT, p1)));
        // individuals that already carry a HaploStruct (HAPLOSTRUCT= given: Main.java:317-330) are preset
        for (int i = 0; i < n; i++) if (inds.get(i).getAllHaploStruct() != null) uploadHaploStruct(i, inds.get(i));
        check((int) invoke(SIMULATE, ctx, maxfounderallele));
        for (int i = 0; i < n; i++) if (inds.get(i).getAllHaploStruct() == null) downloadIndividual(i, inds.get(i));
    }

    /** Seam A over the GPUs of the box: `ranks[r]` is a PsimNative on device r, all made from the same PopulationData.
     *  Every rank runs on its own thread and makes the same calls; with PARTITION_BLOCK every rank ends with the whole
     *  population and rank 0 fills in the Individuals, with PARTITION_LINEAGE the last generations are gathered first. */
    static void simulateIndividualsSharded(PsimNative[] ranks, int maxfounderallele, int partition) throws Exception {
        MemorySegment out = ranks[0].arena.allocate(ADDRESS);
        if ((int) invoke(SHARD_GROUP_CREATE, ranks.length, out) != 0) throw new Exception("psim_shard_group_create failed");
        MemorySegment group = out.get(ADDRESS, 0);
        Exception[] failed = new Exception[ranks.length];
        Thread[] th = new Thread[ranks.length];
        for (int r = 0; r < ranks.length; r++) {
            final int rank = r;
            th[r] = new Thread(() -> {
                try {
                    PsimNative g = ranks[rank];
                    g.setPedigreeAndPresets();
                    g.check((int) invoke(SHARD_INIT_LOCAL, g.ctx, rank, group));
                    g.check((int) invoke(SIMULATE_SHARDED, g.ctx, maxfounderallele, partition));
                    if (partition == PARTITION_LINEAGE) {   // bring everybody to every rank (or emit per rank instead: psim_shard_owner)
                        int n = g.popdata.getIndividual().size();
                        int[] all = new int[n];
                        for (int i = 0; i < n; i++) all[i] = i;
                        g.check((int) invoke(SHARD_GATHER, g.ctx, g.arena.allocateFrom(JAVA_INT, all), (long) n));
                    }
                } catch (Exception e) { failed[rank] = e; }
            });
            th[r].start();
        }
        for (Thread t : th) t.join();
        for (Exception e : failed) if (e != null) throw e;
        ArrayList<Individual> inds = ranks[0].popdata.getIndividual();
        for (int i = 0; i < inds.size(); i++) if (inds.get(i).getAllHaploStruct() == null) ranks[0].downloadIndividual(i, inds.get(i));
        invoke(SHARD_GROUP_DESTROY, group);
    }

    /** The first half of simulateIndividuals: pedigree and recorded haplostructs into the context. */
    private void setPedigreeAndPresets() throws Exception {
        ArrayList<Individual> inds = popdata.getIndividual();
        int n = inds.size();
        int[] p0 = new int[n], p1 = new int[n];
        for (int i = 0; i < n; i++) {
            Individual ind = inds.get(i);
            p0[i] = ind.isFounder() ? -1 : inds.indexOf(ind.getParents()[0]);
            p1[i] = ind.isFounder() ? -1 : inds.indexOf(ind.getParents()[1]);
        }
        check((int) invoke(SET_PEDIGREE, ctx, (long) n, arena.allocateFrom(JAVA_INT, p0), arena.allocateFrom(JAVA_INT, p1)));
        for (int i = 0; i < n; i++) if (inds.get(i).getAllHaploStruct() != null) uploadHaploStruct(i, inds.get(i));
    }

    private void uploadHaploStruct(int i, Individual ind) throws Exception {
        int C = popdata.chromCount(), P = popdata.ploidy;
        long[] off = new long[C * P + 1];
        ArrayList<Integer> f = new ArrayList<>();
        ArrayList<Double> x = new ArrayList<>();
        for (int c = 0; c < C; c++) for (int h = 0; h < P; h++) {
            HaploStruct hs = ind.getHaploStruct(c, h);
            f.addAll(hs.getFounder());
            x.addAll(hs.getRecombPos());
            off[c * P + h + 1] = f.size();
        }
        check((int) invoke(SET_HAPLOSTRUCTS, ctx, 0, (long) i, 1L, arena.allocateFrom(JAVA_LONG, off),
                arena.allocateFrom(JAVA_INT, f.stream().mapToInt(Integer::intValue).toArray()),
                arena.allocateFrom(JAVA_DOUBLE, x.stream().mapToDouble(Double::doubleValue).toArray())));
    }

    private void downloadIndividual(int i, Individual ind) throws Exception {
        int C = popdata.chromCount(), P = popdata.ploidy;
        MemorySegment nseg = arena.allocate(JAVA_LONG);
        check((int) invoke(SEGMENT_COUNT, ctx, 0, (long) i, 1L, nseg));
        long ns = nseg.get(JAVA_LONG, 0);
        MemorySegment off = arena.allocate(JAVA_LONG, C * P + 1), f = arena.allocate(JAVA_INT, ns), x = arena.allocate(JAVA_DOUBLE, ns);
        check((int) invoke(GET_HAPLOSTRUCTS, ctx, 0, (long) i, 1L, off, f, x));
        HaploStruct[][] hs = new HaploStruct[C][P];
        for (int c = 0; c < C; c++) for (int h = 0; h < P; h++) {
            long b = off.getAtIndex(JAVA_LONG, c * P + h), e = off.getAtIndex(JAVA_LONG, c * P + h + 1);
            HaploStruct s = new HaploStruct(popdata.getChrom(c), f.getAtIndex(JAVA_INT, b));
            for (long k = b + 1; k < e; k++) s.addSegment(x.getAtIndex(JAVA_DOUBLE, k), f.getAtIndex(JAVA_INT, k));
            hs[c][h] = s;
        }
        ind.setHaploStruct(hs);
        if (!ind.isFounder()) {     // Individual.parconfig, written later by Main.writeMeioticConfigsFile
            MemorySegment quad = arena.allocate(JAVA_INT, C * 2L), seq = arena.allocate(JAVA_INT, C * 2L * P);
            check((int) invoke(GET_CONFIGS, ctx, 0, (long) i, 1L, quad, seq));
            ind.parconfig = new Genotype.ChromConfig[2][C];
            for (int c = 0; c < C; c++) for (int par = 0; par < 2; par++) {
                Genotype.ChromConfig cc = new Genotype.ChromConfig(P);
                cc.quad = quad.getAtIndex(JAVA_INT, c * 2L + par);
                for (int h = 0; h < P; h++) cc.chromseq[h] = seq.getAtIndex(JAVA_INT, (c * 2L + par) * P + h);
                ind.parconfig[par][c] = cc;
            }
        }
    }

    /** Seam B — the matrices behind writeGenotypesFile(founders=true) and writeAlleledoseFile
     *  (Main.java:903-1022), locus-major like the files; the caller keeps its PrintWriter loops
     *  and reads cells from the returned segments instead of calling getFounderAt per cell. */
    MemorySegment[] emit(boolean withCodes, byte[] codes, int nAlleles, int doseWhich) throws Exception {
        int C = popdata.chromCount(), P = popdata.ploidy, n = popdata.indivCount();
        long[] off = new long[C + 1];
        ArrayList<Double> pos = new ArrayList<>();
        for (int c = 0; c < C; c++) {
            for (Locus l : popdata.getChrom(c).getLocus()) pos.add(l.getPosition());
            off[c + 1] = pos.size();
        }
        long L = off[C];
        check((int) invoke(SET_MAP, ctx, arena.allocateFrom(JAVA_LONG, off),
                arena.allocateFrom(JAVA_DOUBLE, pos.stream().mapToDouble(Double::doubleValue).toArray())));
        if (withCodes) check((int) invoke(SET_CODES, ctx, nAlleles, arena.allocateFrom(JAVA_BYTE, codes)));
        MemorySegment founder = arena.allocate(L * n * P), dose = arena.allocate(L * n);
        MemorySegment allele = withCodes ? arena.allocate(L * n * P) : MemorySegment.NULL;
        MemorySegment t = arena.allocate(TARGETS);
        t.set(ADDRESS, 0, founder); t.set(JAVA_LONG, 8, (long) n * P); t.set(JAVA_INT, 16, 1); t.set(JAVA_INT, 20, 0);
        t.set(ADDRESS, 24, allele); t.set(JAVA_LONG, 32, (long) n * P);
        t.set(ADDRESS, 40, dose); t.set(JAVA_LONG, 48, (long) n); t.set(JAVA_INT, 56, doseWhich);
        check((int) invoke(EMIT, ctx, 0L, (long) n, 0L, L, t));
        return new MemorySegment[] {founder, allele, dose};
    }

    @Override public void close() throws Exception { invoke(DESTROY, ctx); arena.close(); }
}
