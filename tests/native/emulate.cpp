// tests/native/emulate.cpp -- TEST INFRASTRUCTURE.  Runs the product's device logic (colbwt_core.cuh: build_row,
// lane_step, slow_reposition; pack.cpp: pack_read_2bit) one lane at a time on the CPU so that the row packing and
// the traversal state machine can be checked against the oracle without a GPU.  Never shipped, never linked into
// libcolbwt_b200.so.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <type_traits>
#include <vector>

#include "../../col_bwt_b200/csrc/colbwt_core.cuh"
#include "../../col_bwt_b200/csrc/tasks.h"
#include "../../col_bwt_b200/csrc/fastx.h"

namespace colbwt {
bool pack_read_2bit(const uint8_t *seq, uint64_t len, uint32_t *words);
void pack_slice(const uint8_t *seqs, const uint64_t *off, uint64_t r0, uint64_t r1, uint64_t r_base, uint64_t base0,
                uint64_t seq_end, uint32_t *words, uint64_t w, ReadMeta *meta, std::vector<uint64_t> &irr);
}
using namespace colbwt;

struct EmuTable {
    std::vector<Row> rows;
    std::vector<uint64_t> hot, cold;
    uint32_t max_len = 0;
    std::vector<uint8_t> ch8;
    std::vector<uint64_t> idx, thr;
    std::vector<uint32_t> char_rows, char_start;
    uint8_t code_lut[256];
    TableView view;
    uint64_t slow_rows = 0;
    uint32_t flags = 0;
};

// Long-read path: plan chunk tasks (tasks.h) exactly as the driver does, run every task as an independent lane in
// REVERSE scheduling order (the result must not depend on the order), then verify / repair the chains (fixup_chain).
// Returns the number of chunks that had to be re-traversed; *n_tasks_out = chunk tasks run.
template <typename PmlT, bool NARROW>
static uint64_t emu_split_impl(EmuTable *t, const uint8_t *seqs, const uint64_t *off, uint64_t n_reads, void *pml, uint8_t *cid,
                               const SplitParams &sp, uint64_t *n_tasks_out)
{
    BatchView bv{};
    bv.pml = pml;
    bv.cid = cid;
    bv.bytes = seqs;
    uint32_t stage_buf[64] = {0};
    const Stage sg{stage_buf, 1, 0};
    std::vector<uint32_t> words;
    std::vector<uint64_t> word_off(n_reads, 0);
    std::vector<char> is_packed(n_reads, 0), in_plan(n_reads, 0);
    uint64_t nw = 0;
    for (uint64_t i = 0; i < n_reads; ++i) { word_off[i] = nw; nw += ((off[i + 1] - off[i] + 15) >> 4); }
    words.assign(nw + 4, 0);
    TaskPlan plan;
    for (uint64_t i = 0; i < n_reads; ++i) {
        const uint64_t len = off[i + 1] - off[i];
        if (!len) continue;
        is_packed[i] = pack_read_2bit(seqs + off[i], len, words.data() + word_off[i]);
        in_plan[i] = plan.add(off[i], is_packed[i] ? (uint32_t)word_off[i] : (uint32_t)off[i], (uint32_t)len, is_packed[i], sp);   // split, or whole among the tasks
    }
    plan.finish();
    std::vector<ChainState> ss(plan.by_slot.size() + 1), es(plan.by_slot.size() + 1);
    bv.words = words.data();
    bv.tasks = plan.tasks.data();
    bv.by_slot = plan.by_slot.data();
    bv.start_state = ss.data();
    bv.end_state = es.data();
    bv.chains = plan.chains.data();
    bv.n_tasks = plan.n_tasks;
    bv.n_tasks_b = plan.n_tasks_b;
    bv.n_chains = (uint32_t)plan.chains.size();
    for (size_t k = plan.tasks.size(); k-- > 0;) {
        Lane<PmlT> L;
        if (k < plan.n_tasks) {
            lane_begin_task<true>(L, t->view, bv, plan.tasks[k]);
            lane_run<true, NARROW>(L, sg, t->view, bv, t->code_lut);
        } else {
            lane_begin_task<false>(L, t->view, bv, plan.tasks[k]);
            lane_run<false, NARROW>(L, sg, t->view, bv, t->code_lut);
        }
    }
    for (uint64_t i = 0; i < n_reads; ++i) {   // whole reads
        const uint64_t len = off[i + 1] - off[i];
        if (!len || in_plan[i]) continue;
        Lane<PmlT> L;
        ReadMeta m{off[i], (uint32_t)len, is_packed[i] ? (uint32_t)word_off[i] : (uint32_t)off[i]};
        if (is_packed[i]) { lane_begin<true>(L, t->view, bv, m); lane_run<true, NARROW>(L, sg, t->view, bv, t->code_lut); }
        else { lane_begin<false>(L, t->view, bv, m); lane_run<false, NARROW>(L, sg, t->view, bv, t->code_lut); }
    }
    uint64_t redone = 0;
    for (const ChainDesc &c : plan.chains)
        redone += c.packed ? fixup_chain<true, NARROW, PmlT>(sg, t->view, bv, c, t->code_lut) : fixup_chain<false, NARROW, PmlT>(sg, t->view, bv, c, t->code_lut);
    *n_tasks_out = plan.tasks.size();
    return redone;
}


extern "C" {

EmuTable *emu_build(uint64_t n, uint64_t r, const uint8_t *ch, const uint64_t *idx, const uint32_t *interval,
                    const uint16_t *offset, const uint8_t *col_id, const uint64_t *thr)
{
    EmuTable *t = new EmuTable;
    t->ch8.assign(ch, ch + r);
    t->idx.assign(idx, idx + r);
    t->thr.assign(thr, thr + r);
    BuildView b{ch, idx, thr, interval, offset, col_id, n, (uint32_t)r};
    t->rows.resize(r);
    for (uint64_t k = 0; k < r; ++k) {
        uint32_t f = 0;
        t->rows[k] = build_row(b, (uint32_t)k, &f);
        t->flags |= f;
        if (f & BUILD_FLAG_SLOW) ++t->slow_rows;
        t->max_len = std::max(t->max_len, row_len(t->rows[k]));
    }
    t->hot.resize(r);
    t->cold.resize(r);
    for (uint64_t k = 0; k < r; ++k) {
        t->hot[k] = hot_from_row(t->rows[k]);
        t->cold[k] = cold_from_row(t->rows[k]);
    }
    t->char_rows.resize(r);
    std::iota(t->char_rows.begin(), t->char_rows.end(), 0u);
    std::stable_sort(t->char_rows.begin(), t->char_rows.end(), [&](uint32_t a, uint32_t c) { return ch[a] < ch[c]; });
    t->char_start.assign(257, 0);
    for (uint64_t k = 0; k < r; ++k) ++t->char_start[ch[k] + 1];
    for (int c = 0; c < 256; ++c) t->char_start[c + 1] += t->char_start[c];
    for (int c = 0; c < 256; ++c) {
        const int pc = primary_code((uint8_t)c);
        const bool present = t->char_start[c + 1] > t->char_start[c];
        t->code_lut[c] = pc >= 0 ? (uint8_t)pc : (present ? CODE_OTHER : CODE_ABSENT);
    }
    t->view = TableView{t->rows.data(), t->hot.data(), t->cold.data(), t->ch8.data(), t->idx.data(), t->thr.data(), t->char_rows.data(), t->char_start.data(),
                        n, (uint32_t)r, (uint32_t)(n - idx[r - 1])};
    return t;
}

void emu_free(EmuTable *t) { delete t; }
// Slice packer used by the streaming pipeline: returns the number of irregular reads (their indices in irr_out);
// meta_out gets 4 u64 per read: out_off, len, in_off(word), 0.
uint64_t emu_pack_slice(const uint8_t *seqs, const uint64_t *off, uint64_t n_reads, uint32_t *words, uint64_t *meta_out, uint64_t *irr_out)
{
    std::vector<ReadMeta> meta(n_reads);
    std::vector<uint64_t> irr;
    pack_slice(seqs, off, 0, n_reads, 0, off[0], off[n_reads], words, 0, meta.data(), irr);
    for (uint64_t i = 0; i < n_reads; ++i) {
        meta_out[4 * i] = meta[i].out_off;
        meta_out[4 * i + 1] = meta[i].len;
        meta_out[4 * i + 2] = meta[i].in_off;
        meta_out[4 * i + 3] = 0;
    }
    for (size_t k = 0; k < irr.size(); ++k) irr_out[k] = irr[k];
    return irr.size();
}
uint64_t emu_query_split(EmuTable *t, const uint8_t *seqs, const uint64_t *off, uint64_t n_reads, void *pml, int pml_width, uint8_t *cid,
                         uint32_t chunk, uint32_t warm, uint32_t min_len, int narrow, uint64_t *n_tasks_out)
{
    SplitParams sp;
    sp.chunk = chunk;
    sp.warm = warm;
    sp.min_len = std::max(min_len, 2 * chunk);
    sp.mode = 1;
    if (pml_width == 1) return narrow ? emu_split_impl<uint8_t, true>(t, seqs, off, n_reads, pml, cid, sp, n_tasks_out) : emu_split_impl<uint8_t, false>(t, seqs, off, n_reads, pml, cid, sp, n_tasks_out);
    if (pml_width == 2) return narrow ? emu_split_impl<uint16_t, true>(t, seqs, off, n_reads, pml, cid, sp, n_tasks_out) : emu_split_impl<uint16_t, false>(t, seqs, off, n_reads, pml, cid, sp, n_tasks_out);
    return narrow ? emu_split_impl<uint32_t, true>(t, seqs, off, n_reads, pml, cid, sp, n_tasks_out) : emu_split_impl<uint32_t, false>(t, seqs, off, n_reads, pml, cid, sp, n_tasks_out);
}

// The CLI's FASTA/FASTQ(.gz) reader (fastx.h): parses `path` in batches of at most max_bases / max_reads and returns the
// total number of records; sequences are concatenated into seq_out, offsets[n+1], ids joined by '\n' into ids_out.
uint64_t emu_fastx(const char *path, uint64_t max_bases, uint64_t max_reads, uint8_t *seq_out, uint64_t seq_cap, uint64_t *off_out,
                   uint64_t off_cap, char *ids_out, uint64_t ids_cap, uint64_t *n_batches)
{
    FastxReader rd(path);
    if (!rd.ok()) return UINT64_MAX;
    std::vector<uint8_t> seqs;
    std::vector<uint64_t> off;
    std::vector<std::string> ids;
    uint64_t n = 0, bases = 0, idp = 0, batches = 0;
    off_out[0] = 0;
    while (rd.next_batch(seqs, off, ids, max_bases, max_reads)) {
        ++batches;
        for (size_t i = 0; i < ids.size(); ++i) {
            const uint64_t len = off[i + 1] - off[i];
            if (bases + len > seq_cap || n + 2 > off_cap || idp + ids[i].size() + 1 > ids_cap) return UINT64_MAX - 1;
            memcpy(seq_out + bases, seqs.data() + off[i], len);
            bases += len;
            off_out[++n] = bases;
            memcpy(ids_out + idp, ids[i].data(), ids[i].size());
            idp += ids[i].size();
            ids_out[idp++] = '\n';
        }
    }
    ids_out[idp] = 0;
    *n_batches = batches;
    return n;
}

// Chunk planner (tasks.h) for one read: writes lo, hi, top, slot per task (by_slot order); returns the task count.
uint32_t emu_plan(uint32_t len, uint32_t chunk, uint32_t warm, uint32_t *out, uint32_t cap)
{
    SplitParams sp;
    sp.chunk = chunk;
    sp.warm = warm;
    TaskPlan plan;
    plan.add_read(1000, 7, len, true, sp);
    plan.finish();
    for (size_t i = 0; i < plan.by_slot.size() && i < cap; ++i) {
        out[4 * i] = plan.by_slot[i].lo;
        out[4 * i + 1] = plan.by_slot[i].hi;
        out[4 * i + 2] = plan.by_slot[i].top;
        out[4 * i + 3] = plan.by_slot[i].slot;
    }
    return (uint32_t)plan.by_slot.size();
}

// Batch geometry of the long-read path (tasks.h: SplitParams::adapted): out = chunk, warm, min_len, whole_min.
void emu_adapted(uint64_t total_bases, uint64_t lanes, uint32_t *out)
{
    const SplitParams sp = SplitParams().adapted(total_bases, lanes);
    out[0] = sp.chunk;
    out[1] = sp.warm;
    out[2] = sp.min_len;
    out[3] = sp.whole_min();
}

// Streaming chunks of colbwt_query (tasks.h: chunk_geometry + plan_chunks): writes r0 of every chunk (and n_reads at the end)
// to out; returns the number of chunks.  caps[0..1] = largest chunk in bases / in reads (what the staging is sized for).
uint64_t emu_plan_chunks(const uint64_t *off, uint64_t n_reads, uint64_t staged_bytes_per_base, uint64_t *out, uint64_t cap, uint64_t *caps)
{
    uint64_t max_len = 0;
    for (uint64_t i = 0; i < n_reads; ++i) max_len = std::max(max_len, off[i + 1] - off[i]);
    Geometry g = chunk_geometry(off[n_reads] - off[0], n_reads, (uint32_t)max_len, staged_bytes_per_base);
    std::vector<Chunk> chunks;
    plan_chunks(off, n_reads, g, chunks);
    for (size_t i = 0; i < chunks.size() && i < cap; ++i) out[i] = chunks[i].r0;
    if (chunks.size() < cap) out[chunks.size()] = chunks.empty() ? 0 : chunks.back().r1;
    caps[0] = g.chunk_bases;
    caps[1] = g.chunk_reads;
    return chunks.size();
}

int emu_choose_mode(int rule, uint32_t allowed, const double *rate, int n_modes, int large_call)
{
    return choose_mode(rule, allowed, rate, n_modes, large_call != 0);
}
int emu_pack(const uint8_t *seq, uint64_t len, uint32_t *words) { return pack_read_2bit(seq, len, words) ? 1 : 0; }
uint64_t emu_slow_rows(EmuTable *t) { return t->slow_rows; }
uint32_t emu_flags(EmuTable *t) { return t->flags; }
void emu_rows(EmuTable *t, void *out) { memcpy(out, t->rows.data(), t->rows.size() * sizeof(Row)); }

// force_bytes: run every read through the byte (general) path.  pml_width 2 or 4.  Returns lane iterations.
uint32_t emu_max_len(EmuTable *t) { return t->max_len; }

// Experiment support: the sequence of rows gathered (wide layout, packed path) for all reads, -1 between reads.
static std::vector<int64_t> g_trace;
static bool g_trace_on = false;
void emu_trace(int on) { g_trace_on = on != 0; g_trace.clear(); }
uint64_t emu_trace_size() { return g_trace.size(); }
void emu_trace_copy(int64_t *out) { memcpy(out, g_trace.data(), g_trace.size() * 8); }

// force_bytes bit 1: use the narrow (hot/cold) layout.
uint64_t emu_query(EmuTable *t, const uint8_t *seqs, const uint64_t *off, uint64_t n_reads, void *pml, int pml_width,
                   uint8_t *cid, int force_bytes)
{
    const bool narrow = (force_bytes & 2) != 0;
    const bool defer = (force_bytes & 4) != 0;   // bit 2: post flush requests and serve them word by word (flush_word), as the warp does
    const bool intrip = (force_bytes & 8) != 0;  // bit 3: in-trip resolution of same-line neighbours (the kernel variant for tables >= 2 GiB)
    force_bytes &= 1;
    uint64_t iters = 0;
    uint32_t stage_buf[64] = {0};
    const Stage sg{stage_buf, 1, defer ? 5u : 0u};   // deferred mode also exercises the bank rotation
    BatchView bv{};
    bv.pml = pml;
    bv.cid = cid;
    bv.bytes = seqs;
    std::vector<uint32_t> words;
    for (uint64_t i = 0; i < n_reads; ++i) {
        const uint64_t len = off[i + 1] - off[i];
        if (!len) continue;
        words.assign(((len + 15) >> 4) + 2, 0);
        bool packed = !force_bytes && pack_read_2bit(seqs + off[i], len, words.data());
        bv.words = words.data();
        ReadMeta m{off[i], (uint32_t)len, packed ? 0u : (uint32_t)off[i]};
        auto run = [&](auto &L, auto packed_tag) {
            constexpr bool P = decltype(packed_tag)::value;
            lane_begin<P>(L, t->view, bv, m);
            using PmlT = std::remove_reference_t<decltype(L.plen_type())>;
            while (L.state != LANE_IDLE) {
                if (narrow) {
                    const uint64_t *base = ((L.state & 7u) == LANE_COLD) ? t->view.cold : t->view.hot;
                    if (defer) lane_step_narrow<P, true>(L, sg, t->view, bv, base[L.addr], t->code_lut);
                    else lane_step_narrow<P>(L, sg, t->view, bv, base[L.addr], t->code_lut);
                } else {
                    if (g_trace_on) g_trace.push_back(L.addr);
                    if (intrip) lane_step<P, true, true>(L, sg, t->view, bv, ld_row(t->view.rows + L.addr), t->code_lut);   // as k_traverse<..., INTRIP> calls it
                    else if (defer) lane_step<P, true>(L, sg, t->view, bv, ld_row(t->view.rows + L.addr), t->code_lut);
                    else lane_step<P>(L, sg, t->view, bv, ld_row(t->view.rows + L.addr), t->code_lut);
                }
                if (L.flush) {   // what k_traverse's warp does after the step, one "thread" at a time
                    const uint64_t g = L.out_base + L.j;
                    for (uint32_t w = 0; w < 64; ++w)
                        flush_word<PmlT>(stage_buf, sg.swz, w, bv, g & ~(uint64_t)(stage_block<PmlT>() - 1), (uint32_t)(g & (stage_block<PmlT>() - 1)), L.flush - 1);
                    L.flush = 0;
                }
                ++iters;
            }
            if (g_trace_on) g_trace.push_back(-1);
        };
        if (pml_width == 1) {
            Lane<uint8_t> L;
            if (packed) run(L, std::true_type{}); else run(L, std::false_type{});
        } else if (pml_width == 2) {
            Lane<uint16_t> L;
            if (packed) run(L, std::true_type{}); else run(L, std::false_type{});
        } else {
            Lane<uint32_t> L;
            if (packed) run(L, std::true_type{}); else run(L, std::false_type{});
        }
    }
    return iters;
}

} // extern "C"
