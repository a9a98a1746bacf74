// tasks.h -- host-side planning of chunk tasks for long reads (see ChunkTask in colbwt_core.cuh).  Shared by the
// driver (query.cu) and the host emulation used by the CPU tests.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "colbwt_core.cuh"

namespace colbwt {

struct SplitParams {
    uint32_t min_len = 8192;   // reads at least this long are split ...
    uint32_t chunk = 4096;     // ... into chunks of about this many bases ...
    uint32_t warm = 512;       // ... each started this many bases early from the initial state
    int mode = -1;             // -1 auto (split when the longest read would dominate the batch), 0 never, 1 always
    static SplitParams from_env()
    {
        SplitParams p;
        if (const char *e = getenv("COLBWT_SPLIT")) p.mode = atoi(e);
        if (const char *e = getenv("COLBWT_SPLIT_CHUNK")) p.chunk = std::max(8, atoi(e));
        if (const char *e = getenv("COLBWT_SPLIT_WARM")) p.warm = std::max(0, atoi(e));
        if (const char *e = getenv("COLBWT_SPLIT_MIN")) p.min_len = std::max(16, atoi(e));
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
        // the rule is not available to this call: keep what it can of the rule (its low bits), else the lowest allowed mode
        int pick = -1;
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

struct TaskPlan {
    std::vector<ChunkTask> tasks;     // scheduling order: packed tasks (longest first), then byte tasks (longest first)
    std::vector<ChunkTask> by_slot;   // slot order
    std::vector<ChainDesc> chains;
    uint32_t n_tasks = 0, n_tasks_b = 0;
    void clear()
    {
        tasks.clear();
        by_slot.clear();
        chains.clear();
        n_tasks = n_tasks_b = 0;
    }
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
    void finish()
    {
        std::vector<ChunkTask> p, b;
        for (const ChainDesc &c : chains)
            for (uint32_t i = 0; i < c.n_chunks; ++i) (c.packed ? p : b).push_back(by_slot[c.first_slot + i]);
        auto longer = [](const ChunkTask &x, const ChunkTask &y) { return x.len > y.len; };
        std::stable_sort(p.begin(), p.end(), longer);
        std::stable_sort(b.begin(), b.end(), longer);
        n_tasks = (uint32_t)p.size();
        n_tasks_b = (uint32_t)b.size();
        tasks = std::move(p);
        tasks.insert(tasks.end(), b.begin(), b.end());
    }
};

} // namespace colbwt
