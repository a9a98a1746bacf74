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
}
using namespace colbwt;

struct EmuTable {
    std::vector<Row> rows;
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
    t->view = TableView{t->rows.data(), t->ch8.data(), t->idx.data(), t->thr.data(), t->char_rows.data(), t->char_start.data(),
                        n, (uint32_t)r, (uint32_t)(n - idx[r - 1])};
    return t;
}

void emu_free(EmuTable *t) { delete t; }
int emu_pack(const uint8_t *seq, uint64_t len, uint32_t *words) { return pack_read_2bit(seq, len, words) ? 1 : 0; }
uint64_t emu_slow_rows(EmuTable *t) { return t->slow_rows; }
uint32_t emu_flags(EmuTable *t) { return t->flags; }
void emu_rows(EmuTable *t, void *out) { memcpy(out, t->rows.data(), t->rows.size() * sizeof(Row)); }

// force_bytes: run every read through the byte (general) path.  pml_width 2 or 4.  Returns lane iterations.
uint64_t emu_query(EmuTable *t, const uint8_t *seqs, const uint64_t *off, uint64_t n_reads, void *pml, int pml_width,
                   uint8_t *cid, int force_bytes)
{
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
            Policies pol;
            lane_begin<P, 0>(L, t->view, bv, m, pol);
            while (L.state != LANE_IDLE) {
                lane_step<P, 0>(L, t->view, bv, ld_row<0>(t->view.rows + L.addr, pol), t->code_lut, pol);
                ++iters;
            }
        };
        if (pml_width == 2) {
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
