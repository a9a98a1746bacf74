// traverse.cu -- the query hot path: backward move-structure traversal producing PML + chain id per base.
//
// Replaces col_pml::_query_pml / threshold_step (include/col_bwt.hpp:498-574) and LF_table::LF
// (include/ds/LF_table.hpp:251-262) for a whole batch of reads.
//
// Execution model (B200, 148 SMs): a persistent grid of sm_count * CTAS_PER_SM CTAs.  Every lane owns one read
// at a time and is a small state machine (colbwt_core.cuh: lane_step): each trip round the loop issues exactly
// ONE 128-bit read-only row gather per lane -- the LF destination, a fast-forward neighbour or a reposition
// target, whichever that lane needs -- so divergent lanes never serialise extra gathers behind each other and the
// SM keeps (resident warps x 32) independent gathers in flight to cover DRAM/L2 latency.  Lanes that finish a read
// pull the next one from a global cursor (one warp-aggregated atomic), so reads of any length mix freely.
// Outputs are staged per lane in shared memory (64 / 32 / 16 positions for 8 / 16 / 32-bit PML, sized so that the
// carve-out leaves the L1 its 124-156 KB) and written by the whole warp as bursts of up to 64 bytes per array;
// reads long enough to dominate a batch are cut into speculative chunk tasks and repaired by k_fixup (colbwt_core.cuh).
#include <atomic>

#include "internal.h"

namespace colbwt {

constexpr int TRAVERSE_THREADS = 256;

template <bool PACKED, typename PmlT, int CTAS, bool NARROW, bool INTRIP>
__global__ void __launch_bounds__(TRAVERSE_THREADS, CTAS)
k_traverse(const TableView t, const BatchView bv, const uint8_t *__restrict__ code_lut_g, unsigned long long *cursor)
{
    extern __shared__ uint32_t stage_mem[];                     // stage_words<PmlT>() words per thread: its output block
    const uint8_t *code_lut = code_lut_g;                       // byte reads only; 256 bytes that stay in the L1 (no static
                                                                // shared memory: it would push the carve-out up a step)
    constexpr uint32_t STAGE_PITCH = stage_words<PmlT>();
    const uint32_t lane = threadIdx.x & 31;
    // each lane's block is contiguous (the warp flushes it with 32 consecutive words) and rotated by the lane number
    const Stage sg{stage_mem + threadIdx.x * STAGE_PITCH, 1, lane & stage_swz_mask<PmlT>()};
    const uint32_t *warp_stage = stage_mem + (threadIdx.x - lane) * STAGE_PITCH;
    const ReadMeta *meta = PACKED ? bv.meta : bv.meta_b;
    const ChunkTask *tasks = PACKED ? bv.tasks : bv.tasks + bv.n_tasks;
    const uint32_t n_tasks = PACKED ? bv.n_tasks : bv.n_tasks_b;       // chunk tasks of split reads go first (longest work)
    const uint64_t count = (uint64_t)n_tasks + (PACKED ? bv.n_packed : (bv.n_bytes_dev ? __ldg(bv.n_bytes_dev) : bv.n_bytes));
    Lane<PmlT> L;
    bool exhausted = false;
    for (;;) {
        // ---- refill idle lanes: one atomic per warp --------------------------------------------------------
        const uint32_t want = __ballot_sync(0xffffffffu, L.state == LANE_IDLE && !exhausted);
        if (want) {
            const int leader = __ffs(want) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(want));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (want & (1u << lane)) {
                const unsigned long long i = base + __popc(want & ((1u << lane) - 1u));
                if (i < n_tasks) {
                    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(tasks + i));
                    const uint4 b = __ldg(reinterpret_cast<const uint4 *>(tasks + i) + 1);
                    const ChunkTask k{(uint64_t)a.x | ((uint64_t)a.y << 32), a.z, a.w, b.x, b.y, b.z, b.w};
                    lane_begin_task<PACKED>(L, t, bv, k);
                } else if (i < count) {
                    const uint4 mv = __ldg(reinterpret_cast<const uint4 *>(meta + (i - n_tasks)));
                    ReadMeta m;
                    m.out_off = (uint64_t)mv.x | ((uint64_t)mv.y << 32);
                    m.len = mv.z;
                    m.in_off = mv.w;
                    if (m.len) lane_begin<PACKED>(L, t, bv, m);   // zero-length / irregular / split read: nothing to do here
                } else {
                    exhausted = true;
                }
            }
        }
        if (__all_sync(0xffffffffu, exhausted && L.state == LANE_IDLE)) break;
        // ---- one gather per active lane ------------------------------------------------------------------------
        if (L.state != LANE_IDLE) {
            if (NARROW) {
                const uint64_t *base = ((L.state & 7u) == LANE_COLD) ? t.cold : t.hot;
                const uint64_t w = ld_row64(base + L.addr);
                lane_step_narrow<PACKED, true>(L, sg, t, bv, w, code_lut);
            } else {
                const Row row = ld_row(t.rows + L.addr);
                lane_step<PACKED, true, INTRIP>(L, sg, t, bv, row, code_lut);
            }
        }
        // ---- completed output blocks: the warp writes them together (colbwt_core.cuh: flush_word) -------------------
        uint32_t fl = __ballot_sync(0xffffffffu, L.flush != 0);
        if (fl) {
            const uint64_t g = L.out_base + L.j;                // position of the base just emitted (lowest slot of the block)
            const uint32_t mine = (uint32_t)(g & (stage_block<PmlT>() - 1)) | (L.flush << 8);
            const uint64_t my_blk = g & ~(uint64_t)(stage_block<PmlT>() - 1);
            L.flush = 0;
            __syncwarp();
            do {
                const int src = __ffs(fl) - 1;
                fl &= fl - 1;
                const uint64_t blk = __shfl_sync(0xffffffffu, my_blk, src);
                const uint32_t lh = __shfl_sync(0xffffffffu, mine, src);
                const uint32_t *block = warp_stage + src * STAGE_PITCH;
#pragma unroll
                for (uint32_t w = 0; w < stage_words<PmlT>(); w += 32)
                    flush_word<PmlT>(block, (uint32_t)src & stage_swz_mask<PmlT>(), w + lane, bv, blk, lh & 255u, (lh >> 8) - 1u);
            } while (fl);
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Device-side read packing (used by colbwt_query when the caller's sequence buffer is pinned and the reads are short):
// the raw bytes of a chunk are copied to the device as they are and packed here, so the host touches no sequence
// byte at all.  Same encoding and same irregular-read rule as pack.cpp.  One warp per read.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_reads(const uint8_t *__restrict__ bytes, ReadMeta *meta, uint32_t n_reads, uint32_t *words,
                                                     ReadMeta *meta_b, uint32_t *n_irregular)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_reads) return;
    ReadMeta m = meta[warp];                    // host filled out_off (= byte offset of the read in `bytes`), len, in_off (= word offset)
    const uint8_t *p = bytes + m.out_off;
    bool bad = false;
    for (uint32_t w = lane; w * 16 < m.len; w += 32) {
        const uint32_t e = min(16u, m.len - w * 16);
        uint32_t v = 0;
        for (uint32_t k = 0; k < e; ++k) {
            const uint8_t c = p[w * 16 + k];
            bad |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
            v |= (uint32_t)((c >> 1) & 3) << (2 * k);
        }
        words[m.in_off + w] = v;
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) {
        const uint32_t k = atomicAdd(n_irregular, 1u);
        meta_b[k] = ReadMeta{m.out_off, m.len, (uint32_t)m.out_off};   // byte reads index the raw buffer directly
        meta[warp].len = 0;                                            // skipped by the packed pass
    }
}

int launch_pack(const DeviceTable &dt, const uint8_t *d_bytes, ReadMeta *d_meta, uint32_t n_reads, uint32_t *d_words, ReadMeta *d_meta_b,
                uint32_t *d_n_irregular, cudaStream_t stream)
{
    (void)dt;
    CB_CUDA(cudaMemsetAsync(d_n_irregular, 0, sizeof(uint32_t), stream));
    if (n_reads) {
        k_pack_reads<<<(unsigned)(((uint64_t)n_reads * 32 + 255) / 256), 256, 0, stream>>>(d_bytes, d_meta, n_reads, d_words, d_meta_b, d_n_irregular);
        CB_CUDA(cudaGetLastError());
    }
    return COLBWT_OK;
}

// One lane per split read: verify its chunk chain, re-traverse the chunks whose speculative start did not converge.
template <typename PmlT, bool NARROW>
__global__ void __launch_bounds__(128) k_fixup(const TableView t, const BatchView bv, const uint8_t *__restrict__ code_lut_g, unsigned long long *redone)
{
    __shared__ uint8_t code_lut[256];
    __shared__ uint32_t stage_mem[128 * stage_words<PmlT>()];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) code_lut[i] = code_lut_g[i];
    __syncthreads();
    const Stage sg{stage_mem + threadIdx.x, 128, 0};
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= bv.n_chains) return;
    const ChainDesc d = bv.chains[c];
    const uint32_t n = d.packed ? fixup_chain<true, NARROW, PmlT>(sg, t, bv, d, code_lut) : fixup_chain<false, NARROW, PmlT>(sg, t, bv, d, code_lut);
    if (n) atomicAdd(redone, (unsigned long long)n);
}

// 4 CTAs of 256 lanes per SM (42-48 registers, no spills).  More resident lanes do not help: 5 or 6 CTAs give the same
// rate when the L1 is kept large (stage of 32 positions) and half of it when their shared memory squeezes the L1
// (profiles/r1/stage_sweep.log) -- the kernel sits on the DRAM random-line rate.  The shared-memory carve-out is set
// explicitly to the smallest size that holds the 4 CTAs: left to itself the driver sizes it for the occupancy the
// register count would allow and takes the difference from the L1 (42.0 instead of 49.8 Gbases/s on C2).
#ifndef COLBWT_TRAVERSE_CTAS
#define COLBWT_TRAVERSE_CTAS 4
#endif
constexpr int TRAVERSE_CTAS = COLBWT_TRAVERSE_CTAS;

static int carveout_percent(size_t smem_per_cta /* dynamic + static */, int ctas, int device)
{
    int per_sm = 0, reserved = 0;
    cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
    cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, device);
    const size_t need = (size_t)ctas * (smem_per_cta + (size_t)reserved);
    static const size_t steps_kb[] = {0, 8, 16, 32, 64, 100, 132, 164, 196, 228};   // sm_100 carve-out sizes
    size_t pick = (size_t)per_sm;
    for (size_t kb : steps_kb)
        if (kb * 1024 >= need) { pick = std::min(pick, kb * 1024); break; }
    // the hint is a percentage of the maximum and the driver rounds it UP to the next size (58 % = 132.2 KB became 164 KB
    // under ncu), so round down here: 57 % -> 132 KB, 43 % -> 100 KB
    return (int)(pick >= (size_t)per_sm ? 100 : pick * 100 / (size_t)per_sm);
}

template <bool PACKED, typename PmlT, bool NARROW, bool INTRIP>
static int launch_variant(unsigned grid, size_t smem, const DeviceTable &dt, const BatchView &bv, const uint8_t *lut, unsigned long long *cursor, cudaStream_t stream)
{
    auto kern = k_traverse<PACKED, PmlT, TRAVERSE_CTAS, NARROW, INTRIP>;
    static std::atomic<uint64_t> configured{0};   // per instantiation: devices whose function attributes are set
    const uint64_t bit = 1ull << (dt.device & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaFuncAttributes fa;
        CB_CUDA(cudaFuncGetAttributes(&fa, kern));
        CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent(smem + fa.sharedSizeBytes, TRAVERSE_CTAS, dt.device)));
        configured.fetch_or(bit, std::memory_order_release);
    }
    kern<<<grid, TRAVERSE_THREADS, smem, stream>>>(dt.view, bv, lut, cursor);
    return COLBWT_OK;
}

template <bool PACKED, typename PmlT>
static int launch_one(int sm_count, uint32_t reads, const DeviceTable &dt, const BatchView &bv, const uint8_t *lut,
                      unsigned long long *cursor, cudaStream_t stream)
{
    const uint64_t need = (!PACKED && bv.n_bytes_dev) ? (uint64_t)sm_count
                                                      : ((uint64_t)reads + (PACKED ? bv.n_tasks : bv.n_tasks_b) + TRAVERSE_THREADS - 1) / TRAVERSE_THREADS;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(need, (uint64_t)sm_count * TRAVERSE_CTAS));
    const size_t smem = (size_t)TRAVERSE_THREADS * stage_words<PmlT>() * sizeof(uint32_t);
    if (dt.view.hot != nullptr)   // narrow layout, built only when COLBWT_NARROW=1 (index.cu)
        return launch_variant<PACKED, PmlT, true, false>(grid, smem, dt, bv, lut, cursor, stream);
    // in-trip resolution of same-line neighbours pays only where every gather is a DRAM + TLB miss (colbwt_core.cuh)
    const char *env = getenv("COLBWT_INTRIP");   // read per launch: tests switch it
    const int pinned = env ? atoi(env) : -1;
    const bool intrip = pinned >= 0 ? pinned != 0 : (uint64_t)dt.view.r * sizeof(Row) >= (2ull << 30);
    if (intrip) return launch_variant<PACKED, PmlT, false, true>(grid, smem, dt, bv, lut, cursor, stream);
    return launch_variant<PACKED, PmlT, false, false>(grid, smem, dt, bv, lut, cursor, stream);
}

int launch_traverse(const DeviceTable &dt, const BatchView &bv, int pml_width, unsigned long long *d_cursors, cudaStream_t stream)
{
    // [0] cursor of the packed work list, [1] of the byte work list, [2] chunks re-traversed by k_fixup
    CB_CUDA(cudaMemsetAsync(d_cursors, 0, 3 * sizeof(unsigned long long), stream));
    const uint8_t *lut = (const uint8_t *)dt.d_code_lut;
    int rc = COLBWT_OK;
    if (bv.n_packed || bv.n_tasks) {
        if (pml_width == 2) rc = launch_one<true, uint16_t>(dt.sm_count, bv.n_packed, dt, bv, lut, d_cursors, stream);
        else if (pml_width == 1) rc = launch_one<true, uint8_t>(dt.sm_count, bv.n_packed, dt, bv, lut, d_cursors, stream);
        else rc = launch_one<true, uint32_t>(dt.sm_count, bv.n_packed, dt, bv, lut, d_cursors, stream);
        if (rc) return rc;
        CB_CUDA(cudaGetLastError());
    }
    if (bv.n_bytes || bv.n_tasks_b || bv.n_bytes_dev) {   // with a device-side count the pass is always launched (it exits at once when empty)
        if (pml_width == 2) rc = launch_one<false, uint16_t>(dt.sm_count, bv.n_bytes, dt, bv, lut, d_cursors + 1, stream);
        else if (pml_width == 1) rc = launch_one<false, uint8_t>(dt.sm_count, bv.n_bytes, dt, bv, lut, d_cursors + 1, stream);
        else rc = launch_one<false, uint32_t>(dt.sm_count, bv.n_bytes, dt, bv, lut, d_cursors + 1, stream);
        if (rc) return rc;
        CB_CUDA(cudaGetLastError());
    }
    if (bv.n_chains) {   // split reads exist: verify / repair their chunk chains (d_cursors[2] counts re-traversed chunks)
        const unsigned grid = (bv.n_chains + 127) / 128;
        const bool narrow = dt.view.hot != nullptr;
        if (pml_width == 2) {
            if (narrow) k_fixup<uint16_t, true><<<grid, 128, 0, stream>>>(dt.view, bv, lut, d_cursors + 2);
            else k_fixup<uint16_t, false><<<grid, 128, 0, stream>>>(dt.view, bv, lut, d_cursors + 2);
        } else if (pml_width == 4) {
            if (narrow) k_fixup<uint32_t, true><<<grid, 128, 0, stream>>>(dt.view, bv, lut, d_cursors + 2);
            else k_fixup<uint32_t, false><<<grid, 128, 0, stream>>>(dt.view, bv, lut, d_cursors + 2);
        } else {
            if (narrow) k_fixup<uint8_t, true><<<grid, 128, 0, stream>>>(dt.view, bv, lut, d_cursors + 2);
            else k_fixup<uint8_t, false><<<grid, 128, 0, stream>>>(dt.view, bv, lut, d_cursors + 2);
        }
        CB_CUDA(cudaGetLastError());
    }
    return COLBWT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Random-gather roofline microbenchmark (S_rand).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// Load flavours under test (the row gather of k_traverse uses the winner, see ld_row in colbwt_core.cuh).
template <int V> __device__ __forceinline__ uint4 ld_variant(const uint4 *p)
{
    uint4 v;
    if (V == 0) return __ldg(p);                 // ld.global.nc
    else if (V == 1) return __ldcg(p);           // ld.global.cg  (L2 only)
    else if (V == 2) return __ldcs(p);           // ld.global.cs  (streaming)
    else if (V == 3) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (V == 4) return __ldcv(p);           // ld.global.cv
    else if (V == 5) asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (V == 6) return __ldlu(p);           // ld.global.lu
    else if (V == 7) asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else v = *p;                                  // ld.global (default .ca)
    return v;
}

template <bool DEPENDENT, int V>
__global__ void __launch_bounds__(256, 8)
k_gather_bench(const uint4 *__restrict__ buf, uint64_t n_sectors, uint32_t loads_per_thread, uint32_t *sink)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s = mix64(tid + 0x9E3779B97F4A7C15ULL);
    uint32_t acc = 0;
    if (DEPENDENT) {
        for (uint32_t i = 0; i < loads_per_thread; ++i) {
            const uint4 v = ld_variant<V>(buf + 2 * (s % n_sectors));   // one 16-byte half of a random 32-byte sector
            acc ^= v.y;
            s = mix64(s + v.x);                                          // next address depends on the loaded value
        }
    } else {
#pragma unroll 8
        for (uint32_t i = 0; i < loads_per_thread; ++i) {
            const uint4 v = ld_variant<V>(buf + 2 * (s % n_sectors));
            acc ^= v.x ^ v.y;
            s = mix64(s + i);
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <bool DEP, int V> static void launch_gather_v(unsigned grid, size_t smem, const uint4 *buf, uint64_t n_sectors, uint32_t per_thread, uint32_t *sink)
{
    auto kern = k_gather_bench<DEP, V>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, 256, smem>>>(buf, n_sectors, per_thread, sink);
}

template <bool DEP> static void launch_gather(int variant, unsigned grid, size_t smem, const uint4 *buf, uint64_t n_sectors, uint32_t per_thread, uint32_t *sink)
{
    switch (variant) {
    case 1: launch_gather_v<DEP, 1>(grid, smem, buf, n_sectors, per_thread, sink); break;
    case 2: launch_gather_v<DEP, 2>(grid, smem, buf, n_sectors, per_thread, sink); break;
    case 3: launch_gather_v<DEP, 3>(grid, smem, buf, n_sectors, per_thread, sink); break;
    case 4: launch_gather_v<DEP, 4>(grid, smem, buf, n_sectors, per_thread, sink); break;
    case 5: launch_gather_v<DEP, 5>(grid, smem, buf, n_sectors, per_thread, sink); break;
    case 6: launch_gather_v<DEP, 6>(grid, smem, buf, n_sectors, per_thread, sink); break;
    case 7: launch_gather_v<DEP, 7>(grid, smem, buf, n_sectors, per_thread, sink); break;
    case 8: launch_gather_v<DEP, 8>(grid, smem, buf, n_sectors, per_thread, sink); break;
    default: launch_gather_v<DEP, 0>(grid, smem, buf, n_sectors, per_thread, sink); break;
    }
}

__global__ void k_fill(uint4 *buf, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t)mix64(i);
        buf[i] = make_uint4(x, x ^ 0x5bd1e995u, (uint32_t)i, 0);
    }
}

} // namespace colbwt

extern "C" int colbwt_gather_bench(int device, uint64_t bytes, uint64_t loads, int dependent, double *sectors_per_s)
{
    using namespace colbwt;
    if (!sectors_per_s || bytes < 64) {
        set_error("colbwt_gather_bench: bad argument");
        return COLBWT_ERR_ARG;
    }
    CB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (const char *e = getenv("COLBWT_L2_FETCH")) CB_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e)));
    const uint64_t n_sectors = bytes / 32;
    uint4 *buf = nullptr;
    uint32_t *sink = nullptr;
    CB_CUDA(cudaMalloc(&buf, n_sectors * 32));
    CB_CUDA(cudaMalloc(&sink, 4));
    k_fill<<<prop.multiProcessorCount * 8, 256>>>(buf, n_sectors * 2);
    // probe knobs: CTAs (of 256 lanes) resident per SM and dynamic shared memory per CTA -- shared memory comes out of the
    // same 256 KB as the L1, whose lines hold the gathers in flight
    const int gb_ctas = getenv("COLBWT_GB_CTAS") ? std::max(1, std::min(8, atoi(getenv("COLBWT_GB_CTAS")))) : 8;
    const size_t gb_smem = getenv("COLBWT_GB_SMEM") ? (size_t)atoi(getenv("COLBWT_GB_SMEM")) : 0;
    const unsigned grid = prop.multiProcessorCount * gb_ctas;
    const uint64_t threads = (uint64_t)grid * 256;
    const uint32_t per_thread = (uint32_t)std::max<uint64_t>(1, loads / threads);
    cudaEvent_t e0, e1;
    CB_CUDA(cudaEventCreate(&e0));
    CB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {   // first repetition is the warm-up
        CB_CUDA(cudaEventRecord(e0));
        if (dependent & 1) launch_gather<true>(dependent >> 8, grid, gb_smem, buf, n_sectors, per_thread, sink);
        else launch_gather<false>(dependent >> 8, grid, gb_smem, buf, n_sectors, per_thread, sink);
        CB_CUDA(cudaEventRecord(e1));
        CB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        CB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms < best) best = ms;
    }
    CB_CUDA(cudaGetLastError());
    *sectors_per_s = (double)threads * per_thread / (best * 1e-3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaFree(sink);
    return COLBWT_OK;
}
