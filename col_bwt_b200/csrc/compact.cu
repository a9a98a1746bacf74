// compact.cu -- compact result form of one traversed chunk, for the trip over PCIe.
//
// The dense result the reference returns (col_pml::query_pml, include/col_bwt.hpp:409-412: one length and one chain
// id per base) carries little information: a pseudo matching length is reset to 0 on a mismatch and grows by one per
// matching base (col_bwt.hpp:516-523), so the whole PML array of a read is a function of ONE match bit per base, and
// chain ids are 0 except on marked sub-runs (6.6 % of the bases on BASELINE configs[1]).  2-5 bytes per base over a
// 55 GB/s link bound the end-to-end rate well below the kernel's; the compact form is ~0.3 bytes per base:
//   match words   u32[ceil(n/32)]  bit b of word w = (PML of base 32w+b != 0)
//   cid words     u32[ceil(n/32)]  bit b of word w = (chain id of base 32w+b != 0)
//   prefix        u32[G+1]         number of non-zero chain ids before base 2048*g (G = ceil(n/2048) groups); [G] = total
//   values        u8[total]        the non-zero chain ids in base order
// Host side: expand.cpp rebuilds the dense arrays bit-exactly (PML[j] = distance to the next mismatch at or after j).
// Two kernels around one cub scan, all streaming over the chunk's dense PML/CID still in HBM (2-5 B/base read at
// HBM speed: ~2 % of the traversal time).
#include <cub/device/device_scan.cuh>

#include "internal.h"

namespace colbwt {

constexpr uint32_t GROUP_WORDS = COMPACT_GROUP_WORDS;   // 64 words = 2048 bases per prefix entry, one warp per group

// 32 consecutive elements -> bit mask of the non-zero ones.  `valid` < 32 only for the last word of the chunk.
template <typename T> __device__ __forceinline__ uint32_t nonzero_mask(const T *p, uint32_t valid)
{
    uint32_t m = 0;
    if (valid == 32) {
        constexpr int VEC = 16 / sizeof(T);                     // elements per 16-byte vector
        const uint4 *v = reinterpret_cast<const uint4 *>(p);    // p is 32-element aligned: 32 / 64 / 128 bytes
#pragma unroll
        for (int q = 0; q < 32 / VEC; ++q) {
            const uint4 x = __ldcs(v + q);                      // streaming: read once
            const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (sizeof(T) == 4) m |= (uint32_t)(w[k] != 0) << (q * VEC + k);
                else if (sizeof(T) == 2) m |= ((uint32_t)((w[k] & 0xFFFFu) != 0) | ((uint32_t)((w[k] >> 16) != 0) << 1)) << (q * VEC + 2 * k);
                else {
#pragma unroll
                    for (int b = 0; b < 4; ++b) m |= (uint32_t)(((w[k] >> (8 * b)) & 0xFFu) != 0) << (q * VEC + 4 * k + b);
                }
            }
        }
    } else {
        for (uint32_t b = 0; b < valid; ++b) m |= (uint32_t)(p[b] != 0) << b;
    }
    return m;
}

template <typename PmlT>
__global__ void __launch_bounds__(256) k_compact_bits(const PmlT *__restrict__ pml, const uint8_t *__restrict__ cid, uint64_t n_bases,
                                                       uint32_t *__restrict__ match_words, uint32_t *__restrict__ cid_words,
                                                       uint32_t *__restrict__ group_count)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t group = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_words = (n_bases + 31) >> 5;
    if (group * GROUP_WORDS >= n_words) return;
    uint32_t count = 0;
#pragma unroll
    for (uint32_t h = 0; h < GROUP_WORDS / 32; ++h) {
        const uint64_t w = group * GROUP_WORDS + h * 32 + lane;
        if (w < n_words) {
            const uint64_t b0 = w << 5;
            const uint32_t valid = (uint32_t)(n_bases - b0 < 32 ? n_bases - b0 : 32);
            const uint32_t mm = nonzero_mask(pml + b0, valid), cm = nonzero_mask(cid + b0, valid);
            match_words[w] = mm;
            cid_words[w] = cm;
            count += __popc(cm);
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) count += __shfl_xor_sync(0xffffffffu, count, d);
    if (lane == 0) group_count[group] = count;
}

// One warp per group: the non-zero chain ids of its 2048 bases go to values[prefix[group] ...] in base order.
__global__ void __launch_bounds__(256) k_compact_values(const uint8_t *__restrict__ cid, uint64_t n_bases, const uint32_t *__restrict__ cid_words,
                                                         const uint32_t *__restrict__ prefix, uint8_t *__restrict__ values)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t group = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_words = (n_bases + 31) >> 5;
    if (group * GROUP_WORDS >= n_words) return;
    uint32_t at = prefix[group];
#pragma unroll
    for (uint32_t h = 0; h < GROUP_WORDS / 32; ++h) {
        const uint64_t w = group * GROUP_WORDS + h * 32 + lane;
        uint32_t bits = w < n_words ? cid_words[w] : 0;
        const uint32_t mine = __popc(bits);
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)lane >= d) incl += o;
        }
        uint32_t o = at + incl - mine;
        while (bits) {
            const uint32_t b = __ffs(bits) - 1;
            bits &= bits - 1;
            values[o++] = cid[(w << 5) + b];
        }
        at += __shfl_sync(0xffffffffu, incl, 31);
    }
}

size_t compact_scan_temp_bytes(uint64_t max_groups)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)(max_groups + 1));
    return bytes;
}

// d_pml/d_cid: the dense results of one chunk (n_bases of them).  d_out: match words | cid words | prefix (CompactLayout);
// d_values: room for n_bases bytes.  group_count is staged in the prefix area of d_scratch.
int launch_compact(const void *d_pml, int pml_width, const uint8_t *d_cid, uint64_t n_bases, uint8_t *d_out, uint32_t *d_group_count,
                   void *d_scan_temp, size_t scan_temp_bytes, uint8_t *d_values, cudaStream_t stream)
{
    const CompactLayout lay(n_bases);
    if (n_bases == 0) return COLBWT_OK;
    uint32_t *match_words = reinterpret_cast<uint32_t *>(d_out + lay.match_off), *cid_words = reinterpret_cast<uint32_t *>(d_out + lay.cid_off),
             *prefix = reinterpret_cast<uint32_t *>(d_out + lay.prefix_off);
    const unsigned grid = (unsigned)((lay.n_groups * 32 + 255) / 256);
    if (pml_width == 1) k_compact_bits<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t *)d_pml, d_cid, n_bases, match_words, cid_words, d_group_count);
    else if (pml_width == 2) k_compact_bits<uint16_t><<<grid, 256, 0, stream>>>((const uint16_t *)d_pml, d_cid, n_bases, match_words, cid_words, d_group_count);
    else k_compact_bits<uint32_t><<<grid, 256, 0, stream>>>((const uint32_t *)d_pml, d_cid, n_bases, match_words, cid_words, d_group_count);
    CB_CUDA(cudaGetLastError());
    // exclusive sum over n_groups + 1 entries: the last one (input ignored) receives the total
    CB_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_temp, scan_temp_bytes, d_group_count, prefix, (int)(lay.n_groups + 1), stream));
    k_compact_values<<<grid, 256, 0, stream>>>(d_cid, n_bases, cid_words, prefix, d_values);
    CB_CUDA(cudaGetLastError());
    return COLBWT_OK;
}

} // namespace colbwt
