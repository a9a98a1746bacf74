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
// Outputs are staged in registers and written as aligned 16-byte (PML) / 8-byte (CID) vectors.
#include "internal.h"

namespace colbwt {

constexpr int TRAVERSE_THREADS = 256;

template <bool PACKED, typename PmlT>
__global__ void __launch_bounds__(TRAVERSE_THREADS, 2048 / TRAVERSE_THREADS)
k_traverse(const TableView t, const BatchView bv, const uint8_t *__restrict__ code_lut_g, unsigned long long *cursor)
{
    __shared__ uint8_t code_lut[256];
    if (!PACKED) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) code_lut[i] = code_lut_g[i];
        __syncthreads();
    }
    const uint32_t lane = threadIdx.x & 31;
    const ReadMeta *meta = PACKED ? bv.meta : bv.meta_b;
    const uint32_t count = PACKED ? bv.n_packed : bv.n_bytes;
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
                if (i < count) {
                    const uint4 mv = __ldg(reinterpret_cast<const uint4 *>(meta + i));
                    ReadMeta m;
                    m.out_off = (uint64_t)mv.x | ((uint64_t)mv.y << 32);
                    m.len = mv.z;
                    m.in_off = mv.w;
                    if (m.len) lane_begin<PACKED>(L, t, bv, m);   // zero-length read: nothing to emit
                } else {
                    exhausted = true;
                }
            }
        }
        if (__all_sync(0xffffffffu, exhausted && L.state == LANE_IDLE)) break;
        // ---- one gather per active lane ------------------------------------------------------------------------
        if (L.state != LANE_IDLE) {
            const Row row = ld_row(t.rows + L.addr);
            lane_step<PACKED>(L, t, bv, row, code_lut);
        }
    }
}

int launch_traverse(const DeviceTable &dt, const BatchView &bv, int pml_width, unsigned long long *d_cursors, cudaStream_t stream)
{
    // two cursors: [0] packed reads, [1] byte reads
    CB_CUDA(cudaMemsetAsync(d_cursors, 0, 2 * sizeof(unsigned long long), stream));
    const int ctas_per_sm = 2048 / TRAVERSE_THREADS;
    const uint8_t *lut = (const uint8_t *)dt.d_code_lut;
    auto grid_for = [&](uint32_t reads) {
        const uint64_t need = ((uint64_t)reads + TRAVERSE_THREADS - 1) / TRAVERSE_THREADS;
        return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(need, (uint64_t)dt.sm_count * ctas_per_sm));
    };
    if (bv.n_packed) {
        if (pml_width == 2)
            k_traverse<true, uint16_t><<<grid_for(bv.n_packed), TRAVERSE_THREADS, 0, stream>>>(dt.view, bv, lut, d_cursors);
        else
            k_traverse<true, uint32_t><<<grid_for(bv.n_packed), TRAVERSE_THREADS, 0, stream>>>(dt.view, bv, lut, d_cursors);
        CB_CUDA(cudaGetLastError());
    }
    if (bv.n_bytes) {
        if (pml_width == 2)
            k_traverse<false, uint16_t><<<grid_for(bv.n_bytes), TRAVERSE_THREADS, 0, stream>>>(dt.view, bv, lut, d_cursors + 1);
        else
            k_traverse<false, uint32_t><<<grid_for(bv.n_bytes), TRAVERSE_THREADS, 0, stream>>>(dt.view, bv, lut, d_cursors + 1);
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

template <bool DEPENDENT>
__global__ void __launch_bounds__(256, 8)
k_gather_bench(const uint4 *__restrict__ buf, uint64_t n_sectors, uint32_t loads_per_thread, uint32_t *sink)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s = mix64(tid + 0x9E3779B97F4A7C15ULL);
    uint32_t acc = 0;
    if (DEPENDENT) {
        for (uint32_t i = 0; i < loads_per_thread; ++i) {
            const uint4 v = __ldg(buf + 2 * (s % n_sectors));      // one 16-byte half of a random 32-byte sector
            acc ^= v.y;
            s = mix64(s + v.x);                                     // next address depends on the loaded value
        }
    } else {
#pragma unroll 8
        for (uint32_t i = 0; i < loads_per_thread; ++i) {
            const uint4 v = __ldg(buf + 2 * (s % n_sectors));
            acc ^= v.x ^ v.y;
            s = mix64(s + i);
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
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
    const uint64_t n_sectors = bytes / 32;
    uint4 *buf = nullptr;
    uint32_t *sink = nullptr;
    CB_CUDA(cudaMalloc(&buf, n_sectors * 32));
    CB_CUDA(cudaMalloc(&sink, 4));
    k_fill<<<prop.multiProcessorCount * 8, 256>>>(buf, n_sectors * 2);
    const unsigned grid = prop.multiProcessorCount * 8;
    const uint64_t threads = (uint64_t)grid * 256;
    const uint32_t per_thread = (uint32_t)std::max<uint64_t>(1, loads / threads);
    cudaEvent_t e0, e1;
    CB_CUDA(cudaEventCreate(&e0));
    CB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {   // first repetition is the warm-up
        CB_CUDA(cudaEventRecord(e0));
        if (dependent) k_gather_bench<true><<<grid, 256>>>(buf, n_sectors, per_thread, sink);
        else k_gather_bench<false><<<grid, 256>>>(buf, n_sectors, per_thread, sink);
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
