// tasks.h -- host-side planning of chunk tasks for long reads (see ChunkTask in colbwt_core.cuh).  Shared by the
// driver (query.cu) and the host emulation used by the CPU tests.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "colbwt_core.cuh"

namespace colbwt {

struct SplitParams {
    uint32_t min_len = 8192;   // reads at least this long are split ...
    uint32_t chunk = 4096;     // ... into chunks of about this many bases ...
    uint32_t warm = 512;       // ... each started this many bases early from the initial state
    int mode = -1;             // -1 auto (split when the longest read would dominate the batch), 0 never, 1 always
    bool pinned_chunk = false, pinned_warm = false, pinned_min = false;   // set by the environment: adapt() leaves them alone
    static SplitParams from_env()
    {
        SplitParams p;
        if (const char *e = getenv("COLBWT_SPLIT")) p.mode = atoi(e);
        if (const char *e = getenv("COLBWT_SPLIT_CHUNK")) { p.chunk = std::max(8, atoi(e)); p.pinned_chunk = true; }
        if (const char *e = getenv("COLBWT_SPLIT_WARM")) { p.warm = std::max(0, atoi(e)); p.pinned_warm = true; }
        if (const char *e = getenv("COLBWT_SPLIT_MIN")) { p.min_len = std::max(16, atoi(e)); p.pinned_min = true; }
        p.min_len = std::max(p.min_len, 2 * p.chunk);
        return p;
    }
    // `lanes` reads are in flight at a time; splitting pays when one read's serial chain is a large share of the
    // time the whole batch would need at full parallelism.
    bool wanted(uint64_t max_len, uint64_t total_bases, uint64_t lanes) const
    {
        if (mode == 0 || max_len < min_len) return false;
        if (mode == 1) return true;
        return max_len * lanes > total_bases / 4;
    }
    // Geometry for one batch.  A lane runs one task at a time, so the batch ends when the lane with the most work ends:
    // tasks must be SHORT against a lane's share of the batch (total_bases / lanes), or the last tasks -- and every whole
    // read nearly as long as a chunk -- run while most lanes idle.  Measured on configs[2] with 125 k reads per GPU
    // (8.2 kbases per lane, profiles/r2/r2_longread_sweep_c3.log): chunk 4096 -> 52.7 ms, 3072 -> 41.7, 2048 -> 41.9
    // although the shorter chunks traverse 12 % more bases; with the whole reads ordered among the tasks (add_whole) 4096 ->
    // 39.9, 2048 -> 39.3, 1536 -> 36.4, 1024 -> 40.6 (r2_longread_sweep2.log).  So: about six tasks per lane, chunk in [1024, 4096], warm-up
    // 256 for chunks up to 3072 (enough at the error rates that make long reads diverge; a chunk whose warm-up was too
    // short is re-traversed by k_fixup, never wrong), and everything longer than 1.5 chunks is cut.
    SplitParams adapted(uint64_t total_bases, uint64_t lanes) const
    {
        SplitParams p = *this;
        if (!pinned_chunk) {
            const uint64_t share = total_bases / std::max<uint64_t>(1, lanes * 6);
            p.chunk = (uint32_t)std::min<uint64_t>(4096, std::max<uint64_t>(1024, (share + 255) & ~255ull));
        }
        if (!pinned_warm) p.warm = p.chunk <= 3072 ? 256 : 512;
        if (!pinned_min) p.min_len = std::max<uint32_t>(16, p.chunk + p.chunk / 2);
        else p.min_len = std::max(p.min_len, 2 * p.chunk);
        return p;
    }
    // Reads this long, but not long enough to be cut, still take part in the longest-first order of the chunk tasks.
    uint32_t whole_min() const { return std::max<uint32_t>(64, chunk / 8); }
};

// How colbwt_query runs a call that could go several ways (query.cu: where the reads are packed, how the results cross
// the link).  rate[m] = bases/s of the last large call run in mode m (0 = not measured yet); `allowed` = bit mask of the
// modes this call can use; `rule` = the starting guess.  Large calls: the rule first, then every other allowed mode once,
// then the fastest -- the rule keeps its place unless another mode is more than 5 % faster.  Small calls follow what
// the large ones found.  No periodic re-trial: a process that shares the host with seven others (one rank per GPU) would
// only sample noise.
inline int choose_mode(int rule, uint32_t allowed, const double *rate, int n_modes, bool large_call)
{
    if (!(allowed & (1u << rule))) {
        // the rule is not available to this call: keep what it can of the rule -- its transport (mode >> 1) first, then its
        // packing bit (mode & 1) -- else the lowest allowed mode
        int pick = -1;
        for (int m = 0; m < n_modes && pick < 0; ++m)
            if ((allowed & (1u << m)) && (m >> 1) == (rule >> 1)) pick = m;
        for (int m = 0; m < n_modes && pick < 0; ++m)
            if ((allowed & (1u << m)) && (m & 1) == (rule & 1)) pick = m;
        for (int m = 0; m < n_modes && pick < 0; ++m)
            if (allowed & (1u << m)) pick = m;
        rule = pick < 0 ? 0 : pick;
    }
    if (large_call) {
        if (rate[rule] == 0) return rule;
        for (int m = 0; m < n_modes; ++m)
            if ((allowed & (1u << m)) && rate[m] == 0) return m;
    }
    int best = rule;
    if (rate[rule] > 0)
        for (int m = 0; m < n_modes; ++m)
            if ((allowed & (1u << m)) && rate[m] > 1.05 * rate[rule] && rate[m] > rate[best]) best = m;
    return best;
}

// Stable order by decreasing work.  A batch of long reads is ~1e6 tasks of 32 bytes and the feeder thread plans every chunk
// while the GPU waits for the next one: a comparison sort took ~20 ms per 200 k tasks, the counting sort (the key is a length
// of a few thousand) takes one pass.
inline void sort_longest_first(std::vector<ChunkTask> &v)
{
    if (v.size() < 2) return;
    uint32_t mx = 0;
    for (const ChunkTask &t : v) mx = std::max(mx, t.len);
    if ((uint64_t)mx > 8 * (uint64_t)v.size() + (1u << 16)) {   // few tasks, huge keys: not worth a histogram
        std::stable_sort(v.begin(), v.end(), [](const ChunkTask &x, const ChunkTask &y) { return x.len > y.len; });
        return;
    }
    std::vector<uint32_t> at((size_t)mx + 2, 0);
    for (const ChunkTask &t : v) ++at[(size_t)(mx - t.len) + 1];
    for (size_t i = 1; i < at.size(); ++i) at[i] += at[i - 1];
    std::vector<ChunkTask> out(v.size());
    for (const ChunkTask &t : v) out[at[(size_t)(mx - t.len)]++] = t;
    v.swap(out);
}

struct TaskPlan {
    std::vector<ChunkTask> tasks;     // scheduling order: packed tasks (longest first), then byte tasks (longest first)
    std::vector<ChunkTask> by_slot;   // chunk tasks of the split reads in slot order
    std::vector<ChunkTask> wholes;    // whole reads scheduled among the chunk tasks (slot = NO_SLOT: nothing to verify)
    std::vector<char> wholes_packed;
    std::vector<ChainDesc> chains;
    uint32_t n_tasks = 0, n_tasks_b = 0;
    void clear()
    {
        tasks.clear();
        by_slot.clear();
        wholes.clear();
        wholes_packed.clear();
        chains.clear();
        n_tasks = n_tasks_b = 0;
    }
    bool empty() const { return by_slot.empty() && wholes.empty(); }
    // Cut one read into chunk tasks, top chunk first.
    void add_read(uint64_t out_off, uint32_t in_off, uint32_t len, bool packed, const SplitParams &sp)
    {
        const uint32_t k = (len + sp.chunk - 1) / sp.chunk;
        const uint32_t size = (len + k - 1) / k;
        ChainDesc c{(uint32_t)by_slot.size(), 0, packed ? 1u : 0u, 0};
        for (uint32_t hi = len; hi > 0;) {
            const uint32_t lo = hi > size ? hi - size : 0;
            const uint32_t top = (hi == len) ? hi : std::min<uint32_t>(len, hi + sp.warm);
            by_slot.push_back(ChunkTask{out_off, in_off, lo, hi, top, (uint32_t)by_slot.size(), top - lo});
            ++c.n_chunks;
            hi = lo;
        }
        chains.push_back(c);
    }
    // A read that is traversed in one piece but takes its place in the longest-first order of the tasks: with the chunk
    // tasks first and the whole reads after them, a read just short of the splitting length started when the chunk tasks
    // were done and ran alone for the length of two chunks.
    void add_whole(uint64_t out_off, uint32_t in_off, uint32_t len, bool packed)
    {
        wholes.push_back(ChunkTask{out_off, in_off, 0, len, len, NO_SLOT, len});
        wholes_packed.push_back(packed ? 1 : 0);
    }
    // What the driver does with every read once a batch is being split: true = the read became one or more tasks.
    bool add(uint64_t out_off, uint32_t in_off, uint32_t len, bool packed, const SplitParams &sp)
    {
        if (len >= sp.min_len) add_read(out_off, in_off, len, packed, sp);
        else if (len >= sp.whole_min()) add_whole(out_off, in_off, len, packed);
        else return false;
        return true;
    }
    void finish()
    {
        std::vector<ChunkTask> p, b;
        for (const ChainDesc &c : chains)
            for (uint32_t i = 0; i < c.n_chunks; ++i) (c.packed ? p : b).push_back(by_slot[c.first_slot + i]);
        for (size_t i = 0; i < wholes.size(); ++i) (wholes_packed[i] ? p : b).push_back(wholes[i]);
        sort_longest_first(p);
        sort_longest_first(b);
        n_tasks = (uint32_t)p.size();
        n_tasks_b = (uint32_t)b.size();
        tasks = std::move(p);
        tasks.insert(tasks.end(), b.begin(), b.end());
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// Streaming chunks of colbwt_query (query.cu): reads [r0, r1) of the call go through the pipeline together.
// ---------------------------------------------------------------------------------------------------------------------
struct Chunk { uint64_t r0, r1; };
struct Geometry { uint64_t chunk_bases = 0, chunk_reads = 0, grow_bases = 0, min_reads = 0; uint32_t max_len = 0; };

// staged_bytes_per_base: bytes of pinned staging a base of output needs when the destination cannot be DMA-ed into
// directly (0 otherwise); bounds the chunk so that one slot's staging stays under 256 MB.
inline Geometry chunk_geometry(uint64_t total_bases, uint64_t n_reads, uint32_t max_len, uint64_t staged_bytes_per_base)
{
    Geometry g;
    g.max_len = max_len;
    uint64_t chunk_bases = 96ull << 20;   // measured on C2 (profiles/r1/e2e_chunk_sweep.log): 16/24/32/48/96/128/192 M -> 83.6/85.0/81.1/76.3/71.8-73.2/72.0/73.1 ms
    const char *env = getenv("COLBWT_CHUNK_BASES");
    if (env) chunk_bases = std::max<uint64_t>(1024, strtoull(env, nullptr, 10));
    // Long reads: a chunk must still hold enough reads to occupy the lanes (a lane works on one read or chunk task at a
    // time and several chunks are in flight), so a chunk that reaches chunk_bases with fewer than 32 Ki reads keeps growing,
    // up to 512 Mbases.  Decided chunk by chunk (plan_chunks), not from the batch's mean read length: in a mixed batch
    // (configs[4]: 12.4 M short reads followed by 125 k long ones) the long reads would otherwise be cut into 96-Mbase
    // chunks of 9.6 k reads each, every one of them as slow as its longest serial chain.
    uint64_t grow_bases = env ? chunk_bases : (512ull << 20);
    if (staged_bytes_per_base && !env) {
        chunk_bases = std::min<uint64_t>(chunk_bases, (256ull << 20) / staged_bytes_per_base);
        grow_bases = std::min<uint64_t>(grow_bases, (256ull << 20) / staged_bytes_per_base);
    }
    chunk_bases = std::max<uint64_t>(chunk_bases, max_len);
    chunk_bases = std::min<uint64_t>(chunk_bases, std::max<uint64_t>(total_bases, 16));
    g.chunk_bases = chunk_bases;
    g.grow_bases = std::min<uint64_t>(std::max(grow_bases, chunk_bases), std::max<uint64_t>(total_bases, 16));
    g.min_reads = 32768;
    // a chunk also ends after chunk_bases/32 reads, which bounds the meta staging (16 B per read) for very short reads
    g.chunk_reads = std::min<uint64_t>(std::max<uint64_t>(chunk_bases / 32, 1024), n_reads);
    return g;
}

// Cuts the batch into chunks and leaves the capacities the staging needs (largest chunk in bases and in reads) in g.
inline void plan_chunks(const uint64_t *off, uint64_t n_reads, Geometry &g, std::vector<Chunk> &chunks)
{
    chunks.clear();
    uint64_t cap_bases = 16, cap_reads = 1;
    for (uint64_t r0 = 0; r0 < n_reads;) {   // at most chunk_bases bases (grow_bases while short of min_reads reads) and chunk_reads reads, at least one read
        uint64_t r1 = (uint64_t)(std::upper_bound(off + r0, off + n_reads + 1, off[r0] + g.chunk_bases) - off) - 1;
        if (r1 - r0 < g.min_reads && g.grow_bases > g.chunk_bases) {
            const uint64_t far = (uint64_t)(std::upper_bound(off + r0, off + n_reads + 1, off[r0] + g.grow_bases) - off) - 1;
            r1 = std::max(r1, std::min(far, r0 + g.min_reads));
        }
        r1 = std::min(r1, r0 + g.chunk_reads);
        if (r1 <= r0) r1 = r0 + 1;
        chunks.push_back(Chunk{r0, r1});
        cap_bases = std::max(cap_bases, off[r1] - off[r0]);
        cap_reads = std::max(cap_reads, r1 - r0);
        r0 = r1;
    }
    g.chunk_bases = cap_bases;
    g.chunk_reads = cap_reads;
}

} // namespace colbwt
