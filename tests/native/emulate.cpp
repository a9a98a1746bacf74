// tests/native/emulate.cpp -- TEST INFRASTRUCTURE.  Runs the product's device logic (colbwt_core.cuh: build_row,
// lane_step, slow_reposition; pack.cpp: pack_read_2bit) one lane at a time on the CPU so that the row packing and
// the traversal state machine can be checked against the oracle without a GPU.  Never shipped, never linked into
// libcolbwt_b200.so.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../col_bwt_b200/csrc/colbwt_core.cuh"

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
    force_bytes &= 1;
    uint64_t iters = 0;
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
            while (L.state != LANE_IDLE) {
                if (narrow) {
                    const uint64_t *base = ((L.state & 7u) == LANE_COLD) ? t->view.cold : t->view.hot;
                    lane_step_narrow<P>(L, t->view, bv, base[L.addr], t->code_lut);
                } else {
                    if (g_trace_on) g_trace.push_back(L.addr);
                    lane_step<P>(L, t->view, bv, ld_row(t->view.rows + L.addr), t->code_lut);
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
