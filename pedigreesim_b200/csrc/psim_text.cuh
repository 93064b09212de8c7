// K3 — table rows as text
// (part of libpsim: included by psim.cu inside namespace psim)
#pragma once

// ------------------------------------------------------------------------------------------
// K3 — the reference's table files as text, formatted on the device (Main.writeGenotypesFile /
// writeAlleledoseFile, Main.java:903-1022: `label { '\t' token } '\n'` per locus, one
// PrintWriter.print per cell in the reference). One block per table row: the row's cells are read
// into shared memory a chunk at a time, every thread sizes (pass 1) or writes (pass 2) the tokens of
// its run of consecutive cells at the offset a block scan gives it, and the chunk's text leaves
// shared memory with coalesced stores. Pass 1 only measures; a device scan of the row lengths
// places the rows back to back.
constexpr int TEXT_THREADS = 256;

struct TextArgs {
    const uint8_t* cells; int64_t pitch; int cell_bytes;   // device table [rows][pitch], u8 or u16 cells
    int64_t n_cells;               // tokens per row
    int64_t row0_locus;            // global locus of table row 0
    int n_rows;
    const int64_t* label_off;      // [L+1] locus names (psim_set_locus_names)
    const char* labels;
    int kind;                      // 0 = decimal numbers, 1 = the locus' allele names by code
    const int64_t* code_off;       // kind 1: [L+1] first name of each locus
    const int64_t* name_off;       //         [names+1]
    const char* names;
    int max_tok;                   // longest token including its tab
    int chunk;                     // cells per chunk, a multiple of TEXT_THREADS
    // _allAlleledose rows: several rows per locus, labelled "name TAB allele"
    const int32_t* row_locus;      // [n_rows] locus of the row relative to row0_locus (null: row r is locus row0_locus + r)
    const int32_t* row_sub;        // [n_rows] which allele of the locus the row counts
    int sub_kind;                  // 0 none, 1 the founder allele number, 2 the locus' allele name number row_sub
    int64_t* row_len;              // pass 1 out [n_rows]
    const int64_t* row_off;        // pass 2 in  [n_rows+1]
    char* text;                    // pass 2 out
};

template <bool WRITE>
__global__ void __launch_bounds__(TEXT_THREADS) k_text_rows(const TextArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_warp[TEXT_THREADS / 32];
    __shared__ uint32_t s_nlen[256], s_nbeg[256];
    const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
    uint8_t* const s_cells = smem;                                   // chunk * cell_bytes
    char* const s_text = reinterpret_cast<char*>(smem + (size_t)a.chunk * a.cell_bytes);   // chunk * max_tok (pass 2)
    const int cpt = a.chunk / TEXT_THREADS;
    for (int r = blockIdx.x; r < a.n_rows; r += gridDim.x) {
        const int64_t gl = a.row0_locus + (a.row_locus ? a.row_locus[r] : r);
        const uint8_t* row = a.cells + (int64_t)r * a.pitch;
        const int64_t lab0 = a.label_off[gl];
        int lab_n = (int)(a.label_off[gl + 1] - lab0);
        char* out = WRITE ? a.text + a.row_off[r] : nullptr;
        if (WRITE)
            for (int i = tid; i < lab_n; i += TEXT_THREADS) out[i] = a.labels[lab0 + i];
        if (a.sub_kind) {   // "TAB allele" after the locus name
            const uint32_t sub = (uint32_t)a.row_sub[r];
            if (a.sub_kind == 1) {
                const int nd = sub < 10 ? 1 : sub < 100 ? 2 : sub < 1000 ? 3 : sub < 10000 ? 4 : 5;
                if (WRITE && tid == 0) {
                    out[lab_n] = '\t';
                    uint32_t v = sub;
                    for (int k = nd - 1; k >= 0; k--) { out[lab_n + 1 + k] = (char)('0' + v % 10); v /= 10; }
                }
                lab_n += 1 + nd;
            } else {
                const int64_t nb = a.code_off[gl] + sub;
                const int nl = (int)(a.name_off[nb + 1] - a.name_off[nb]);
                if (WRITE) {
                    if (tid == 0) out[lab_n] = '\t';
                    for (int i = tid; i < nl; i += TEXT_THREADS) out[lab_n + 1 + i] = a.names[a.name_off[nb] + i];
                }
                lab_n += 1 + nl;
            }
        }
        int64_t base = 0;   // first name of the locus
        if (a.kind == 1) {
            base = a.code_off[gl];
            const int k = (int)(a.code_off[gl + 1] - base);
            __syncthreads();
            for (int i = tid; i < k && i < 256; i += TEXT_THREADS) {
                s_nbeg[i] = (uint32_t)(a.name_off[base + i] - a.name_off[base]);
                s_nlen[i] = (uint32_t)(a.name_off[base + i + 1] - a.name_off[base + i]);
            }
        }
        int64_t running = lab_n;
        for (int64_t c0 = 0; c0 < a.n_cells; c0 += a.chunk) {
            const int n = (int)min((int64_t)a.chunk, a.n_cells - c0);
            __syncthreads();   // previous chunk's text has been copied out; name table loaded
            {   // the chunk's cells, coalesced (rows start 4-byte aligned: staging pitches are multiples of 128)
                const int nbytes = n * a.cell_bytes;
                const uint8_t* src = row + c0 * a.cell_bytes;
                for (int i = tid * 4; i < (nbytes & ~3); i += TEXT_THREADS * 4)
                    *reinterpret_cast<uint32_t*>(s_cells + i) = *reinterpret_cast<const uint32_t*>(src + i);
                if (tid < (nbytes & 3)) s_cells[(nbytes & ~3) + tid] = src[(nbytes & ~3) + tid];
            }
            __syncthreads();
            const int i0 = tid * cpt, i1 = min(i0 + cpt, n);
            auto cell = [&](int i) -> uint32_t {
                return a.cell_bytes == 1 ? (uint32_t)s_cells[i] : (uint32_t)reinterpret_cast<const uint16_t*>(s_cells)[i];
            };
            auto tok_len = [&](uint32_t v) -> uint32_t {
                if (a.kind == 1) return 1 + s_nlen[v & 255];
                return v < 10 ? 2 : v < 100 ? 3 : v < 1000 ? 4 : v < 10000 ? 5 : 6;
            };
            uint32_t len = 0;
            for (int i = i0; i < i1; i++) len += tok_len(cell(i));
            // block exclusive scan of the per-thread lengths
            uint32_t x = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
            if (lane == 31) s_warp[wq] = x;
            __syncthreads();
            uint32_t wbase = 0, total = 0;
#pragma unroll
            for (int w = 0; w < TEXT_THREADS / 32; w++) { const uint32_t v = s_warp[w]; if (w < wq) wbase += v; total += v; }
            if (WRITE) {
                char* p = s_text + wbase + x - len;
                for (int i = i0; i < i1; i++) {
                    const uint32_t v = cell(i);
                    *p++ = '\t';
                    if (a.kind == 1) {
                        const char* nm = a.names + a.name_off[base] + s_nbeg[v & 255];
                        const uint32_t nl = s_nlen[v & 255];
                        for (uint32_t q = 0; q < nl; q++) *p++ = nm[q];
                    } else {
                        if (v >= 10000) *p++ = (char)('0' + v / 10000);
                        if (v >= 1000) *p++ = (char)('0' + v / 1000 % 10);
                        if (v >= 100) *p++ = (char)('0' + v / 100 % 10);
                        if (v >= 10) *p++ = (char)('0' + v / 10 % 10);
                        *p++ = (char)('0' + v % 10);
                    }
                }
                __syncthreads();
                char* dst = out + running;
                // coalesced copy: head bytes up to 4-byte alignment of the destination, then words
                const uint32_t mis = (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3);
                const uint32_t head = min(mis, total);
                if (tid < head) dst[tid] = s_text[tid];
                const uint32_t words = (total - head) >> 2;
                for (uint32_t i = tid; i < words; i += TEXT_THREADS) {
                    const char* s = s_text + head + 4 * i;   // shared memory side may be misaligned: assemble bytes
                    const uint32_t w = (uint32_t)(uint8_t)s[0] | ((uint32_t)(uint8_t)s[1] << 8) | ((uint32_t)(uint8_t)s[2] << 16) | ((uint32_t)(uint8_t)s[3] << 24);
                    *reinterpret_cast<uint32_t*>(dst + head + 4 * i) = w;
                }
                const uint32_t done = head + 4 * words;
                if (tid < total - done) dst[done + tid] = s_text[done + tid];
            }
            running += total;
        }
        if (tid == 0) {
            if (WRITE) out[running] = '\n';
            else a.row_len[r] = running + 1;
        }
    }
}

