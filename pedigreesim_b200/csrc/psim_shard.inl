// One pedigree over the GPUs of a box, inside the library (include/psim.h "psim_simulate_sharded";
// included at the end of psim.cu). The dependency is the reference's own: an individual needs the
// haplostructs of its parents (Main.simulateIndividuals, Main.java:1504-1520), so the generations of
// the pedigree are simulated in turn, every rank its share, and after a generation the ranks hand
// each other the new haplostructs that somebody else needs.
//   plan       shard_plan(): generations, owner rank and "shared" mark of every individual (pure host)
//   exchange   shard_exchange(): pack (k_pack_counts, scan, k_pack_segments) -> sizes -> transfers
//              straight into the receivers' slabs -> k_unpack_desc; one host wait per exchange
//   transport  NCCL (libnccl.so.2 loaded at run time; grouped ncclBroadcast = all-gather of ragged
//              parts) for one process per GPU; a rendezvous of threads for several contexts in one
//              process (peer copies), which is also how a single-GPU box tests the whole path
#include <dlfcn.h>

#include <chrono>
#include <condition_variable>
#include <mutex>

struct psim_shard_group {
    int world = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    uint64_t phase = 0;
    int64_t value[PSIM_SHARD_MAX_WORLD] = {};
    const void* src[4 * PSIM_SHARD_MAX_WORLD] = {};
    int device[PSIM_SHARD_MAX_WORLD] = {};
};

namespace {

// ------------------------------------------------------------------------------------ the plan
struct ShardPlan {
    std::vector<int32_t> level, owner;
    std::vector<uint8_t> share;
    int32_t max_level = 0;
    std::vector<int32_t> order;       // individuals of level 1, 2, ... (ascending inside a level)
    std::vector<int64_t> level_off;   // [max_level + 2]
};

// level = 1 + max(level of the parents); founders and individuals that have a haplostruct are level 0.
// PSIM_PARTITION_BLOCK: a generation in `world` contiguous blocks, everybody shared.
// PSIM_PARTITION_LINEAGE: a child goes to the rank of its first owned parent; a child of parents that are
// everywhere is free, and the free ones are dealt out in blocks, lightest rank first. A parent with many
// children in one generation (the single F1 plant of a RIL chain) is shared instead of followed, and so is
// a second parent that lives on another rank. A shared individual is everywhere after its generation.
int shard_plan(int64_t n, const int32_t* p0, const int32_t* p1, const uint8_t* has_hs, int world, int partition, ShardPlan& pl, std::string& err) {
    pl.level.assign((size_t)n, 0);
    pl.owner.assign((size_t)n, -1);
    pl.share.assign((size_t)n, 0);
    pl.max_level = 0;
    auto settled = [&](int64_t i) { return p0[i] < 0 || (has_hs && has_hs[i]); };
    for (int64_t i = 0; i < n; i++) {
        if (settled(i)) continue;
        if (p0[i] >= i || p1[i] >= i || p1[i] < 0) { err = "pedigree is not in parents-first order"; return PSIM_EINVAL; }
        pl.level[i] = 1 + std::max(pl.level[p0[i]], pl.level[p1[i]]);
        pl.max_level = std::max(pl.max_level, pl.level[i]);
    }
    pl.level_off.assign((size_t)pl.max_level + 2, 0);
    for (int64_t i = 0; i < n; i++) if (pl.level[i] > 0) pl.level_off[pl.level[i] + 1]++;
    for (int lv = 1; lv <= pl.max_level + 1; lv++) pl.level_off[lv] += pl.level_off[lv - 1];
    pl.order.assign((size_t)pl.level_off[pl.max_level + 1], 0);
    {
        std::vector<int64_t> fill(pl.level_off.begin(), pl.level_off.end());
        for (int64_t i = 0; i < n; i++) if (pl.level[i] > 0) pl.order[(size_t)fill[pl.level[i]]++] = (int32_t)i;
    }
    auto deal_blocks = [&](const int32_t* items, int64_t cnt, std::vector<int64_t>* load) {
        const int64_t base = cnt / world, extra = cnt % world;
        int64_t at = 0;
        for (int r = 0; r < world; r++) {
            const int64_t k = base + (r < extra ? 1 : 0);
            for (int64_t q = 0; q < k; q++) pl.owner[items[at + q]] = r;
            if (load) (*load)[r] += k;
            at += k;
        }
    };
    if (partition == PSIM_PARTITION_BLOCK) {
        for (int lv = 1; lv <= pl.max_level; lv++) {
            deal_blocks(pl.order.data() + pl.level_off[lv], pl.level_off[lv + 1] - pl.level_off[lv], nullptr);
            for (int64_t k = pl.level_off[lv]; k < pl.level_off[lv + 1]; k++) pl.share[pl.order[(size_t)k]] = 1;
        }
        return PSIM_OK;
    }
    std::vector<int64_t> load((size_t)world, 0);
    std::vector<int32_t> kids((size_t)n, 0), free_list;
    for (int lv = 1; lv <= pl.max_level; lv++) {
        const int32_t* inds = pl.order.data() + pl.level_off[lv];
        const int64_t cnt = pl.level_off[lv + 1] - pl.level_off[lv];
        const double heavy = std::max(2.0, (double)cnt / (4.0 * world));
        for (int side = 0; side < 2; side++) {   // (the first parents' marks count when the second parents are looked at)
            const int32_t* par = side == 0 ? p0 : p1;
            for (int64_t k = 0; k < cnt; k++) { const int32_t q = par[inds[k]]; if (pl.owner[q] >= 0 && !pl.share[q]) kids[q]++; }
            for (int64_t k = 0; k < cnt; k++) { const int32_t q = par[inds[k]]; if (kids[q] > heavy) pl.share[q] = 1; }
            for (int64_t k = 0; k < cnt; k++) kids[par[inds[k]]] = 0;
        }
        free_list.clear();
        for (int64_t k = 0; k < cnt; k++) {
            const int32_t i = inds[k];
            const int32_t a = pl.share[p0[i]] ? -1 : pl.owner[p0[i]], b = pl.share[p1[i]] ? -1 : pl.owner[p1[i]];
            const int32_t own = a >= 0 ? a : b;
            if (own < 0) { free_list.push_back(i); continue; }
            pl.owner[i] = own;
            load[own]++;
        }
        {
            // the other parent lives elsewhere: it is shared (it belongs to an earlier generation, and the whole plan
            // is made before the first generation runs). Collected first and marked afterwards: the marks of this
            // generation do not move its own individuals.
            std::vector<int32_t> clash;
            for (int64_t k = 0; k < cnt; k++) {
                const int32_t i = inds[k];
                const int32_t a = pl.share[p0[i]] ? -1 : pl.owner[p0[i]], b = pl.share[p1[i]] ? -1 : pl.owner[p1[i]];
                if (a >= 0 && b >= 0 && a != b) clash.push_back(p1[i]);
            }
            for (int32_t q : clash) pl.share[q] = 1;
        }
        if (free_list.empty()) continue;
        const int64_t lo = *std::min_element(load.begin(), load.end()), hi = *std::max_element(load.begin(), load.end());
        if (lo == hi) { deal_blocks(free_list.data(), (int64_t)free_list.size(), &load); continue; }
        int64_t sum = 0;
        for (int64_t v : load) sum += v;
        const double target = (double)(sum + (int64_t)free_list.size()) / world;
        std::vector<int> by_load((size_t)world);
        std::iota(by_load.begin(), by_load.end(), 0);
        std::stable_sort(by_load.begin(), by_load.end(), [&](int x, int y) { return load[x] < load[y]; });
        int64_t at = 0;
        for (int r : by_load) {
            const int64_t want = (int64_t)std::max(0.0, std::nearbyint(target - (double)load[r]));
            const int64_t k = std::min<int64_t>((int64_t)free_list.size() - at, want);
            for (int64_t q = 0; q < k; q++) pl.owner[free_list[(size_t)(at + q)]] = r;
            load[r] += k;
            at += k;
        }
        if (at < (int64_t)free_list.size()) {   // rounding leftovers
            const int r = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            for (int64_t q = at; q < (int64_t)free_list.size(); q++) pl.owner[free_list[(size_t)q]] = r;
            load[r] += (int64_t)free_list.size() - at;
        }
    }
    return PSIM_OK;
}

// ------------------------------------------------------------------------------------ transports
struct XferOp { int root; const void* src; void* dst; int64_t bytes; };   // src is valid on the root, dst on the others

struct ShardTransport {
    virtual ~ShardTransport() {}
    // the device value *mine_d of every rank (stream-ordered on c->stream), on the host when the call returns
    virtual int sizes(psim_ctx_impl* c, const int64_t* mine_d, int64_t* all_h) = 0;
    // every op: `bytes` bytes from the root's src to dst on every other rank (stream-ordered on c->stream)
    virtual int xfer(psim_ctx_impl* c, const XferOp* ops, int n) = 0;
    // `bytes` bytes of every rank's send to slot r of recv on every OTHER rank (the own slot may or may not be written)
    virtual int allgather(psim_ctx_impl* c, const void* send, void* recv, int64_t bytes) = 0;
};

struct Nccl {
    typedef struct ncclComm* comm_t;
    struct unique_id { char internal[128]; };
    enum { kUint8 = 1, kInt64 = 4 };
    void* handle = nullptr;
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
    bool load() {
        if (handle) return true;
        // (a process that already runs NCCL - torch.distributed - gets that very library: same soname)
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) { error = std::string("libnccl.so.2 cannot be loaded: ") + dlerror(); return false; }
        auto sym = [&](const char* name) { void* p = dlsym(handle, name); if (!p) error = std::string("libnccl has no ") + name; return p; };
        GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        AllGather = (decltype(AllGather))sym("ncclAllGather");
        Broadcast = (decltype(Broadcast))sym("ncclBroadcast");
        Send = (decltype(Send))sym("ncclSend");
        Recv = (decltype(Recv))sym("ncclRecv");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllGather || !Broadcast || !Send || !Recv || !GroupStart || !GroupEnd || !GetErrorString) {
            handle = nullptr;
            return false;
        }
        return true;
    }
};
Nccl g_nccl;
std::mutex g_nccl_mutex;

#define NCCL_TRY(call)                                                                                      \
    do {                                                                                                    \
        const int r__ = (call);                                                                             \
        if (r__ != 0) return fail(c, PSIM_ECUDA, std::string(#call) + ": " + g_nccl.GetErrorString(r__));   \
    } while (0)

struct NcclTransport : ShardTransport {
    Nccl::comm_t comm = nullptr;
    int rank = 0, world = 1;
    int64_t* all_d = nullptr;
    int64_t* all_h = nullptr;       // mapped page-locked
    int64_t* all_hd = nullptr;
    ~NcclTransport() override {
        if (comm) g_nccl.CommDestroy(comm);
        cudaFree(all_d);
        if (all_h) cudaFreeHost(all_h);
    }
    int sizes(psim_ctx_impl* c, const int64_t* mine_d, int64_t* out) override {
        NCCL_TRY(g_nccl.AllGather(mine_d, all_d, 1, Nccl::kInt64, comm, c->stream));
        k_publish_i64<<<1, 32, 0, c->stream>>>(all_hd, all_d, world);
        c->stats.kernel_launches++;
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (int r = 0; r < world; r++) out[r] = all_h[r];
        return PSIM_OK;
    }
    int allgather(psim_ctx_impl* c, const void* send, void* recv, int64_t bytes) override {
        if (bytes <= 0) return PSIM_OK;
        NCCL_TRY(g_nccl.AllGather(send, recv, (size_t)bytes, Nccl::kUint8, comm, c->stream));
        return PSIM_OK;
    }
    int xfer(psim_ctx_impl* c, const XferOp* ops, int n) override {
        // one group of ncclBroadcast, a part each (tuning builds, PSIM_SHARD_SENDRECV=1: point-to-point transfers instead,
        // the root of a part to every other rank - 4 % faster on 2 GPUs, 16 % slower on 8)
        static const int env_bcast = tuning_int("PSIM_SHARD_SENDRECV") ? 0 : 1;
        NCCL_TRY(g_nccl.GroupStart());
        for (int k = 0; k < n; k++) {
            if (ops[k].bytes <= 0) continue;
            int r = 0;
            if (env_bcast) {
                void* const dst = ops[k].root == rank ? const_cast<void*>(ops[k].src) : ops[k].dst;   // (in place on the root)
                r = g_nccl.Broadcast(ops[k].src, dst, (size_t)ops[k].bytes, Nccl::kUint8, ops[k].root, comm, c->stream);
            } else if (ops[k].root == rank) {
                for (int p = 0; p < world && r == 0; p++)
                    if (p != rank) r = g_nccl.Send(ops[k].src, (size_t)ops[k].bytes, Nccl::kUint8, p, comm, c->stream);
            } else {
                r = g_nccl.Recv(ops[k].dst, (size_t)ops[k].bytes, Nccl::kUint8, ops[k].root, comm, c->stream);
            }
            if (r != 0) { g_nccl.GroupEnd(); return fail(c, PSIM_ECUDA, std::string("NCCL transfer: ") + g_nccl.GetErrorString(r)); }
        }
        NCCL_TRY(g_nccl.GroupEnd());
        return PSIM_OK;
    }
};

struct LocalTransport : ShardTransport {
    psim_shard_group* g = nullptr;
    int rank = 0;
    int barrier(psim_ctx_impl* c) {
        std::unique_lock<std::mutex> lk(g->m);
        const uint64_t ph = g->phase;
        if (++g->arrived == g->world) { g->arrived = 0; g->phase++; g->cv.notify_all(); return PSIM_OK; }
        if (!g->cv.wait_for(lk, std::chrono::seconds(120), [&] { return g->phase != ph; }))
            return fail(c, PSIM_ESTATE, "shard group: a rank did not arrive (every context of the group must make the same calls)");
        return PSIM_OK;
    }
    int sizes(psim_ctx_impl* c, const int64_t* mine_d, int64_t* out) override {
        int64_t v = 0;
        CUDA_TRY(cudaMemcpyAsync(&v, mine_d, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        g->value[rank] = v;
        if (int rc = barrier(c)) return rc;
        for (int r = 0; r < g->world; r++) out[r] = g->value[r];
        return barrier(c);
    }
    int allgather(psim_ctx_impl* c, const void* send, void* recv, int64_t bytes) override {
        CUDA_TRY(cudaStreamSynchronize(c->stream));   // what this rank sends is complete
        g->device[rank] = c->cfg.device;
        g->src[rank] = send;
        if (int rc = barrier(c)) return rc;
        for (int r = 0; r < g->world; r++)
            if (r != rank && bytes > 0)
                CUDA_TRY(cudaMemcpyPeerAsync((uint8_t*)recv + (size_t)r * bytes, c->cfg.device, g->src[r], g->device[r], (size_t)bytes, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        return barrier(c);
    }
    int xfer(psim_ctx_impl* c, const XferOp* ops, int n) override {
        if (n > 4 * PSIM_SHARD_MAX_WORLD) return fail(c, PSIM_EINVAL, "too many transfers in one exchange");
        CUDA_TRY(cudaStreamSynchronize(c->stream));   // what this rank sends is complete
        g->device[rank] = c->cfg.device;
        for (int k = 0; k < n; k++) if (ops[k].root == rank) g->src[k] = ops[k].src;
        if (int rc = barrier(c)) return rc;
        for (int k = 0; k < n; k++)
            if (ops[k].root != rank && ops[k].bytes > 0)
                CUDA_TRY(cudaMemcpyPeerAsync(ops[k].dst, c->cfg.device, g->src[k], g->device[ops[k].root], (size_t)ops[k].bytes, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        return barrier(c);   // (nobody overwrites what it sent before everybody has read it)
    }
};

struct ShardState {
    int rank = 0, world = 1;
    ShardTransport* tr = nullptr;
    ShardPlan plan;
    bool plan_cached = false;
    uint64_t plan_ped_version = 0;
    int plan_partition = -1;
    std::vector<int32_t> lists_h;       // per (level, rank): the individuals that rank shares after that level
    std::vector<int64_t> list_off;      // [(max_level + 1) * world + 1]
    std::vector<uint8_t> level_direct;  // per level: every rank shares all it simulates (its new segments travel straight out of the slab)
    psim_ctx_impl::SimPlan mine;        // per level: the individuals this rank simulates, as psim_simulate's generation lists
    std::vector<int32_t> owner;         // after the last run: rank that holds an individual, -1 = everybody
    std::vector<int32_t> owner_after;   // ... as the plan leaves it
    bool owner_current = false;         // owner == owner_after
    psim_ctx_impl::Buf lists_d, cnt, send, offs;
    psim_shard_stats stats{};
    ~ShardState() {
        delete tr;
        cudaFree(lists_d.p); cudaFree(cnt.p); cudaFree(send.p); cudaFree(offs.p);
    }
};

void shard_free(psim_ctx_impl* c) {
    if (c->shard && c->lev_plan == &c->shard->mine) c->lev_plan = nullptr;
    delete c->shard;
    c->shard = nullptr;
}

inline int64_t align8(int64_t x) { return (x + 7) / 8 * 8; }

// The ranks hand each other the haplostructs of the listed individuals: lists[list_off[r] .. list_off[r + 1])
// (device: lists_d, host: lists_h) is what rank r contributes and every other rank receives.
// Every rank shares ALL it has just simulated (the block partition): the new segments are the tail of the rank's slab
// from `from` on and travel straight out of it - no pack pass, no scan - as three all-gathers (the collective NCCL does
// best: NVLS / ring over NVSwitch) of equal, padded parts: the homologs' descriptors relative to `from`, the founders
// and the positions, the last two straight into the receivers' slabs (rank r's part at slot r; the own slot is a
// duplicate that nobody reads, the price of the plain collective).
// local_rc != 0: this rank failed in the generation; it still takes part in the size exchange (with -1) so that every
// rank leaves the call with an error instead of waiting for it in a collective.
int shard_exchange_all(psim_ctx_impl* c, ShardState* S, const int32_t* lists_d, const int32_t* lists_h, const int64_t* list_off, int64_t from, int local_rc) {
    const int world = S->world, rank = S->rank, C = c->C, P = c->P;
    const int64_t hp = (int64_t)C * P;
    int64_t nh_max = 0;
    for (int r = 0; r < world; r++) nh_max = std::max(nh_max, (list_off[r + 1] - list_off[r]) * hp);
    if (nh_max == 0 || world == 1) return PSIM_OK;
    const int64_t nh = (list_off[rank + 1] - list_off[rank]) * hp;
    const int64_t mine = local_rc ? -1 : nh > 0 ? c->slab_used - from : 0;
    if (int rc = ensure_buf(c, S->cnt, 2 * sizeof(int64_t))) return rc;
    CUDA_TRY(cudaMemcpyAsync(S->cnt.p, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));   // (pageable: staged before it returns)
    int64_t nseg[PSIM_SHARD_MAX_WORLD];
    if (int rc = S->tr->sizes(c, (const int64_t*)S->cnt.p, nseg)) return rc;
    S->stats.collectives++;
    if (local_rc) return local_rc;
    for (int r = 0; r < world; r++)
        if (nseg[r] < 0) return fail(c, PSIM_ESTATE, "rank " + std::to_string(r) + " of the sharded pedigree failed in this generation");
    int64_t seg_max = 0, in_seg = 0;
    for (int r = 0; r < world; r++) { seg_max = std::max(seg_max, nseg[r]); if (r != rank) in_seg += nseg[r]; }
    // slots behind the widest possible send window, so that no rank's window overlaps what it receives
    const int64_t slots_at = from + std::max(seg_max, c->slab_used - from);
    if (int rc = ensure_slab(c, slots_at + (int64_t)world * seg_max)) return rc;
    if (int rc = ensure_buf(c, S->send, (size_t)(8 * nh_max))) return rc;
    if (int rc = ensure_buf(c, S->offs, (size_t)(8 * nh_max * world))) return rc;
    if (nh > 0) {
        k_pack_desc<<<grid_for(nh, 256), 256, 0, c->stream>>>(nh, lists_d + list_off[rank], C, c->N, P, c->desc_d, from, (uint64_t*)S->send.p);
        c->stats.kernel_launches++;
    }
    if (int rc = S->tr->allgather(c, S->send.p, S->offs.p, 8 * nh_max)) return rc;
    if (int rc = S->tr->allgather(c, c->seg_founder_d + from, c->seg_founder_d + slots_at, 4 * seg_max)) return rc;
    if (int rc = S->tr->allgather(c, c->seg_pos_d + from, c->seg_pos_d + slots_at, 8 * seg_max)) return rc;
    S->stats.collectives++;
    for (int r = 0; r < world; r++) {
        const int64_t nh_r = (list_off[r + 1] - list_off[r]) * hp;
        if (r == rank || nh_r == 0) continue;
        k_unpack_desc_rel<<<grid_for(nh_r, 256), 256, 0, c->stream>>>(nh_r, lists_d + list_off[r], C, c->N, P, (const uint64_t*)((uint8_t*)S->offs.p + (size_t)r * 8 * nh_max),
                                                                      slots_at + (int64_t)r * seg_max, c->desc_d);
        c->stats.kernel_launches++;
        for (int64_t k = list_off[r]; k < list_off[r + 1]; k++) c->hs_reps[lists_h[k]] = c->R;
    }
    CUDA_TRY(cudaGetLastError());
    c->slab_used = slots_at + (int64_t)world * seg_max;
    c->stats.n_segments = c->slab_used;
    c->runs_valid = false;
    c->fresh = false;
    S->stats.bytes_sent += nh > 0 ? 8 * nh + 12 * nseg[rank] : 0;
    S->stats.segments_received += in_seg;
    return PSIM_OK;
}

// The ranks hand each other the haplostructs of the listed individuals: lists[list_off[r] .. list_off[r + 1])
// (device: lists_d, host: lists_h) is what rank r contributes and every other rank receives.
int shard_exchange(psim_ctx_impl* c, ShardState* S, const int32_t* lists_d, const int32_t* lists_h, const int64_t* list_off, int local_rc = 0) {
    const int world = S->world, rank = S->rank, C = c->C, P = c->P;
    const int64_t hp = (int64_t)C * P;
    int64_t any = 0;
    for (int r = 0; r < world; r++) any += list_off[r + 1] - list_off[r];
    if (any == 0 || world == 1) return PSIM_OK;
    // ---- pack what this rank contributes
    const int64_t nh = (list_off[rank + 1] - list_off[rank]) * hp;
    if (int rc = ensure_buf(c, S->cnt, (size_t)(nh + 1) * 2 * sizeof(int64_t))) return rc;
    int64_t* const cnt = (int64_t*)S->cnt.p;
    int64_t* const off = cnt + nh + 1;
    k_pack_counts<<<grid_for(nh + 1, 256), 256, 0, c->stream>>>(nh, lists_d + list_off[rank], C, c->N, P, c->desc_d, cnt);
    c->stats.kernel_launches++;
    if (int rc = exclusive_scan_i64(c, cnt, off, nh + 1)) return rc;
    if (local_rc) {
        const int64_t minus = -1;
        CUDA_TRY(cudaMemcpyAsync(off + nh, &minus, sizeof(minus), cudaMemcpyHostToDevice, c->stream));
    }
    int64_t nseg[PSIM_SHARD_MAX_WORLD];
    if (int rc = S->tr->sizes(c, off + nh, nseg)) return rc;
    S->stats.collectives++;
    if (local_rc) return local_rc;
    for (int r = 0; r < world; r++)
        if (nseg[r] < 0) return fail(c, PSIM_ESTATE, "rank " + std::to_string(r) + " of the sharded pedigree failed in this generation");
    if (nseg[rank] >= (1LL << 32)) return fail(c, PSIM_ECAPACITY, "more than 2^32 segments in one exchange");
    const int64_t off_bytes = align8(4 * (nh + 1)), fo_bytes = align8(4 * nseg[rank]);
    if (int rc = ensure_buf(c, S->send, (size_t)(off_bytes + fo_bytes + 8 * nseg[rank]))) return rc;
    uint8_t* const send = (uint8_t*)S->send.p;
    if (nh > 0) {
        k_pack_segments<<<grid_for(nh + 1, 256), 256, 0, c->stream>>>(nh, lists_d + list_off[rank], C, c->N, P, c->desc_d, off, c->seg_pos_d, c->seg_founder_d,
                                                                      (uint32_t*)send, (int32_t*)(send + off_bytes), (double*)(send + off_bytes + fo_bytes));
        c->stats.kernel_launches++;
    }
    // ---- where the others' parts go: offsets into a receive buffer, segments straight into the slab
    int64_t in_seg = 0, in_off_bytes = 0;
    for (int r = 0; r < world; r++) {
        if (r == rank) continue;
        in_seg += nseg[r];
        in_off_bytes += align8(4 * ((list_off[r + 1] - list_off[r]) * hp + 1));
    }
    if (int rc = ensure_slab(c, c->slab_used + in_seg)) return rc;
    if (int rc = ensure_buf(c, S->offs, (size_t)in_off_bytes)) return rc;
    XferOp ops[3 * PSIM_SHARD_MAX_WORLD];
    int n_ops = 0;
    int64_t slab_at = c->slab_used, offs_at = 0;
    int64_t base[PSIM_SHARD_MAX_WORLD], offs_pos[PSIM_SHARD_MAX_WORLD];
    for (int r = 0; r < world; r++) {
        const int64_t nh_r = (list_off[r + 1] - list_off[r]) * hp;
        if (nh_r == 0) continue;
        const int64_t ob = align8(4 * (nh_r + 1));
        if (r == rank) {
            ops[n_ops++] = {r, send, nullptr, ob};
            ops[n_ops++] = {r, send + off_bytes, nullptr, 4 * nseg[r]};
            ops[n_ops++] = {r, send + off_bytes + fo_bytes, nullptr, 8 * nseg[r]};
            continue;
        }
        base[r] = slab_at; offs_pos[r] = offs_at;
        ops[n_ops++] = {r, nullptr, (uint8_t*)S->offs.p + offs_at, ob};
        ops[n_ops++] = {r, nullptr, c->seg_founder_d + slab_at, 4 * nseg[r]};
        ops[n_ops++] = {r, nullptr, c->seg_pos_d + slab_at, 8 * nseg[r]};
        slab_at += nseg[r]; offs_at += ob;
    }
    if (int rc = S->tr->xfer(c, ops, n_ops)) return rc;
    S->stats.collectives++;
    for (int r = 0; r < world; r++) {
        const int64_t nh_r = (list_off[r + 1] - list_off[r]) * hp;
        if (r == rank || nh_r == 0) continue;
        k_unpack_desc<<<grid_for(nh_r, 256), 256, 0, c->stream>>>(nh_r, lists_d + list_off[r], C, c->N, P, (const uint32_t*)((uint8_t*)S->offs.p + offs_pos[r]),
                                                                  base[r], c->desc_d);
        c->stats.kernel_launches++;
        for (int64_t k = list_off[r]; k < list_off[r + 1]; k++) c->hs_reps[lists_h[k]] = c->R;
    }
    CUDA_TRY(cudaGetLastError());
    c->slab_used = slab_at;
    c->stats.n_segments = c->slab_used;
    c->runs_valid = false;
    c->fresh = false;
    S->stats.bytes_sent += nh > 0 ? off_bytes + 12 * nseg[rank] : 0;
    S->stats.segments_received += in_seg;
    return PSIM_OK;
}

int shard_ready(psim_ctx_impl* c) {
    if (!c->shard || !c->shard->tr) return fail(c, PSIM_ESTATE, "psim_shard_init (or psim_shard_init_local) must be called first");
    if (c->R != 1) return fail(c, PSIM_EINVAL, "a sharded pedigree needs replicate_count == 1 (replicates are sharded by replicate id, without any exchange)");
    if (c->N == 0) return fail(c, PSIM_ESTATE, "psim_set_pedigree must be called first");
    return PSIM_OK;
}

int shard_attach(psim_ctx_impl* c, int rank, int world, ShardTransport* tr) {
    shard_free(c);
    c->shard = new ShardState();
    c->shard->rank = rank; c->shard->world = world; c->shard->tr = tr;
    return PSIM_OK;
}

}  // namespace

extern "C" {

int psim_shard_plan(int64_t n_ind, const int32_t* parent0, const int32_t* parent1, const uint8_t* has_haplostruct, int32_t world, int32_t partition,
                    int32_t* level, int32_t* owner, uint8_t* share) {
    if (n_ind <= 0 || !parent0 || !parent1 || world < 1 || world > PSIM_SHARD_MAX_WORLD) return PSIM_EINVAL;
    if (partition != PSIM_PARTITION_BLOCK && partition != PSIM_PARTITION_LINEAGE) return PSIM_EINVAL;
    for (int64_t i = 0; i < n_ind; i++)
        if (parent0[i] >= n_ind || parent1[i] >= n_ind || (parent0[i] < 0) != (parent1[i] < 0)) return PSIM_EINVAL;
    ShardPlan pl;
    std::string err;
    if (int rc = shard_plan(n_ind, parent0, parent1, has_haplostruct, world, partition, pl, err)) return rc;
    if (level) std::copy(pl.level.begin(), pl.level.end(), level);
    if (owner) std::copy(pl.owner.begin(), pl.owner.end(), owner);
    if (share) std::copy(pl.share.begin(), pl.share.end(), share);
    return PSIM_OK;
}

int psim_shard_unique_id(uint8_t* id) {
    if (!id) return PSIM_EINVAL;
    std::lock_guard<std::mutex> lk(g_nccl_mutex);
    if (!g_nccl.load()) { g_create_error = g_nccl.error; return PSIM_ECUDA; }
    Nccl::unique_id u;
    const int r = g_nccl.GetUniqueId(&u);
    if (r != 0) { g_create_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return PSIM_ECUDA; }
    static_assert(sizeof(u) == PSIM_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, sizeof(u));
    return PSIM_OK;
}

int psim_shard_init(psim_ctx* ctx, int32_t rank, int32_t world, const uint8_t* id) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    if (!id || world < 1 || world > PSIM_SHARD_MAX_WORLD || rank < 0 || rank >= world) return fail(c, PSIM_EINVAL, "psim_shard_init: bad rank / world / id");
    cudaSetDevice(c->cfg.device);
    if (int rc = settle(c)) return rc;
    {
        std::lock_guard<std::mutex> lk(g_nccl_mutex);
        if (!g_nccl.load()) return fail(c, PSIM_ECUDA, g_nccl.error);
    }
    NcclTransport* tr = new NcclTransport();
    tr->rank = rank; tr->world = world;
    Nccl::unique_id u;
    memcpy(&u, id, sizeof(u));
    const int r = g_nccl.CommInitRank(&tr->comm, world, u, rank);
    if (r != 0) { tr->comm = nullptr; delete tr; return fail(c, PSIM_ECUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)); }
    if (cudaMalloc((void**)&tr->all_d, PSIM_SHARD_MAX_WORLD * sizeof(int64_t)) != cudaSuccess ||
        cudaHostAlloc((void**)&tr->all_h, PSIM_SHARD_MAX_WORLD * sizeof(int64_t), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&tr->all_hd, tr->all_h, 0) != cudaSuccess) {
        delete tr;
        return fail(c, PSIM_ECUDA, "psim_shard_init: allocation failed");
    }
    return shard_attach(c, rank, world, tr);
}

int psim_shard_group_create(int32_t world, psim_shard_group** out) {
    if (!out || world < 1 || world > PSIM_SHARD_MAX_WORLD) return PSIM_EINVAL;
    *out = new psim_shard_group();
    (*out)->world = world;
    return PSIM_OK;
}

void psim_shard_group_destroy(psim_shard_group* group) { delete group; }

int psim_shard_init_local(psim_ctx* ctx, int32_t rank, psim_shard_group* group) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    if (!group || rank < 0 || rank >= group->world) return fail(c, PSIM_EINVAL, "psim_shard_init_local: bad rank / group");
    cudaSetDevice(c->cfg.device);
    if (int rc = settle(c)) return rc;
    LocalTransport* tr = new LocalTransport();
    tr->g = group; tr->rank = rank;
    return shard_attach(c, rank, group->world, tr);
}

int psim_simulate_sharded(psim_ctx* ctx, int32_t max_founder_allele, int32_t partition) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
    if (int rc = settle(c)) return rc;
    if (int rc = shard_ready(c)) return rc;
    if (partition != PSIM_PARTITION_BLOCK && partition != PSIM_PARTITION_LINEAGE) return fail(c, PSIM_EINVAL, "unknown partition");
    ShardState* S = c->shard;
    const auto t0 = std::chrono::steady_clock::now();
    const int world = S->world, rank = S->rank;
    const int64_t N = c->N;
    // ---- the plan: the same on every rank (same pedigree, same individuals simulated so far); kept across a
    // psim_reset loop like the generation lists of psim_simulate
    const bool fresh = c->fresh;
    if (!(fresh && S->plan_cached && S->plan_ped_version == c->ped_version && S->plan_partition == partition)) {
        std::vector<uint8_t> has((size_t)N);
        for (int64_t i = 0; i < N; i++) has[(size_t)i] = c->hs_reps[i] != 0;
        std::string err;
        if (int rc = shard_plan(N, c->par0_h.data(), c->par1_h.data(), has.data(), world, partition, S->plan, err)) return fail(c, rc, err);
        const ShardPlan& pl = S->plan;
        const int levels = pl.max_level;
        S->list_off.assign((size_t)(levels + 1) * world + 1, 0);
        S->lists_h.clear();
        S->mine = psim_ctx_impl::SimPlan();
        S->mine.level_off.assign((size_t)levels + 2, 0);
        S->level_direct.assign((size_t)levels + 1, 0);
        for (int lv = 1; lv <= levels; lv++) {
            bool all_shared = true;
            for (int64_t k = pl.level_off[lv]; k < pl.level_off[lv + 1] && all_shared; k++) all_shared = pl.share[pl.order[(size_t)k]] != 0;
            S->level_direct[(size_t)lv] = all_shared;
            for (int r = 0; r < world; r++) {
                S->list_off[(size_t)lv * world + r] = (int64_t)S->lists_h.size();
                for (int64_t k = pl.level_off[lv]; k < pl.level_off[lv + 1]; k++) {
                    const int32_t i = pl.order[(size_t)k];
                    if (pl.owner[i] == r && pl.share[i]) S->lists_h.push_back(i);
                }
            }
            S->mine.level_off[lv] = (int64_t)S->mine.order.size();
            for (int64_t k = pl.level_off[lv]; k < pl.level_off[lv + 1]; k++)
                if (pl.owner[pl.order[(size_t)k]] == rank) S->mine.order.push_back(pl.order[(size_t)k]);
        }
        S->list_off[(size_t)(levels + 1) * world] = (int64_t)S->lists_h.size();
        for (int r = 0; r < world; r++) S->list_off[(size_t)r] = 0;   // (level 0: nothing)
        S->mine.level_off[(size_t)levels + 1] = (int64_t)S->mine.order.size();
        S->mine.max_level = levels;
        S->mine.valid = true;
        S->owner_current = false;
        S->owner_after.assign((size_t)N, -1);   // who holds an individual after the run: its owner, everybody if it was shared
        for (int64_t i = 0; i < N; i++)
            if (pl.level[(size_t)i] > 0 && !pl.share[(size_t)i]) S->owner_after[(size_t)i] = pl.owner[(size_t)i];
        if (c->lev_plan == &S->mine) c->lev_plan = nullptr;   // (a new list: the device copy is stale)
        if (int rc = ensure_buf(c, S->lists_d, S->lists_h.size() * sizeof(int32_t))) return rc;
        if (!S->lists_h.empty())
            CUDA_TRY(cudaMemcpyAsync(S->lists_d.p, S->lists_h.data(), S->lists_h.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        S->plan_cached = fresh; S->plan_ped_version = c->ped_version; S->plan_partition = partition;
    }
    const ShardPlan& pl = S->plan;
    S->stats = psim_shard_stats{};
    // ---- founders: every rank initialises them (counter-based streams: no exchange needed)
    static const int32_t none = 0;
    if (int rc = simulate_impl(c, max_founder_allele, &none, 0)) return rc;
    // (the generation lists of this rank were made with the plan and stay on the device across a psim_reset loop: a
    // generation costs its kernel, not a walk over its individuals on the host)
    for (int lv = 1; lv <= pl.max_level; lv++) {
        const int64_t slab_before = c->slab_used;
        int rc_lv = PSIM_OK;
        std::string err_lv;
        if (S->mine.level_off[lv + 1] > S->mine.level_off[lv]) {
            rc_lv = sim_start(c, S->mine, !S->mine.on_device, lv, lv, false);
            if (!rc_lv) rc_lv = sim_finish(c);
            if (rc_lv) err_lv = c->err;
        }
        const int64_t* lo = S->list_off.data() + (size_t)lv * world;
        int rc_x;
        if (S->level_direct[(size_t)lv]) rc_x = shard_exchange_all(c, S, (const int32_t*)S->lists_d.p, S->lists_h.data(), lo, slab_before, rc_lv);
        else rc_x = shard_exchange(c, S, (const int32_t*)S->lists_d.p, S->lists_h.data(), lo, rc_lv);
        if (rc_lv) return fail(c, rc_lv, err_lv);   // (the other ranks have been told in the size exchange, if there was one)
        if (rc_x) return rc_x;
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (!S->owner_current) { S->owner = S->owner_after; S->owner_current = true; }   // (psim_shard_gather changes its copy)
    S->stats.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return PSIM_OK;
}

int psim_shard_gather(psim_ctx* ctx, const int32_t* individuals, int64_t count) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
    if (int rc = settle(c)) return rc;
    if (int rc = shard_ready(c)) return rc;
    ShardState* S = c->shard;
    if (count > 0 && !individuals) return fail(c, PSIM_EINVAL, "null individual list");
    if ((int64_t)S->owner.size() != c->N) return fail(c, PSIM_ESTATE, "psim_simulate_sharded must be called first");
    std::vector<int32_t> lists;
    std::vector<int64_t> off((size_t)S->world + 1, 0);
    for (int r = 0; r < S->world; r++) {
        off[(size_t)r] = (int64_t)lists.size();
        for (int64_t k = 0; k < count; k++) {
            const int32_t i = individuals[k];
            if (i < 0 || i >= c->N) return fail(c, PSIM_EINVAL, "individual index out of bounds");
            if (S->owner[(size_t)i] == r) lists.push_back(i);
        }
    }
    off[(size_t)S->world] = (int64_t)lists.size();
    if (lists.empty()) return PSIM_OK;
    psim_ctx_impl::Buf tmp;
    if (int rc = ensure_buf(c, tmp, lists.size() * sizeof(int32_t))) return rc;
    int rc = PSIM_OK;
    if (cudaMemcpyAsync(tmp.p, lists.data(), lists.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
        rc = fail(c, PSIM_ECUDA, "psim_shard_gather: upload failed");
    if (!rc) rc = shard_exchange(c, S, (const int32_t*)tmp.p, lists.data(), off.data());
    cudaStreamSynchronize(c->stream);
    cudaFree(tmp.p);
    if (!rc) { for (int32_t i : lists) S->owner[(size_t)i] = -1; S->owner_current = false; }
    return rc;
}

int psim_shard_owner(psim_ctx* ctx, int32_t* owner) {
    psim_ctx_impl* c = CTX;
    if (!c || !owner) return PSIM_EINVAL;
    if (!c->shard || (int64_t)c->shard->owner.size() != c->N) return fail(c, PSIM_ESTATE, "psim_simulate_sharded must be called first");
    std::copy(c->shard->owner.begin(), c->shard->owner.end(), owner);
    return PSIM_OK;
}

int psim_shard_get_stats(psim_ctx* ctx, psim_shard_stats* out) {
    psim_ctx_impl* c = CTX;
    if (!c || !out) return PSIM_EINVAL;
    if (!c->shard) return fail(c, PSIM_ESTATE, "psim_shard_init (or psim_shard_init_local) must be called first");
    *out = c->shard->stats;
    return PSIM_OK;
}

}  // extern "C"
