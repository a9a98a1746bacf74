// Scratch probe (not part of the product): does anything stop the L2 from fetching the whole 128-byte line for a
// random 16-byte load?  Variants: cudaLimitMaxL2FetchGranularity, texture-object fetch, cp.async, TMA bulk copy,
// L2 cache-hint policies.  Run under ncu to read dram__bytes_read.sum per kernel.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

template <int V>
__global__ void __launch_bounds__(256, 8) k_probe(const uint4 *__restrict__ buf, cudaTextureObject_t tex, uint64_t n_sectors, uint32_t per_thread, uint32_t *sink)
{
    __shared__ __align__(128) uint4 stage[256 * 2];
    __shared__ __align__(8) uint64_t bar;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s = mix64(tid + 0x9E3779B97F4A7C15ULL);
    uint32_t acc = 0;
    uint64_t pol = 0;
    if (V == 4) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (V == 5) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    if (V == 6) asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol));
    for (uint32_t i = 0; i < per_thread; ++i) {
        const uint64_t e = 2 * (s % n_sectors);
        uint4 v;
        if (V == 0) v = __ldg(buf + e);
        else if (V == 1) v = tex1Dfetch<uint4>(tex, (int)e);
        else if (V == 2) {   // cp.async 16 B
            unsigned sa = (unsigned)__cvta_generic_to_shared(&stage[threadIdx.x]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(buf + e));
            asm volatile("cp.async.wait_all;");
            v = stage[threadIdx.x];
        } else if (V == 3) {   // ld with .L2::64B? no: plain ld.global.cg
            v = __ldcg(buf + e);
        } else if (V >= 4 && V <= 6) {
            asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + e), "l"(pol));
        } else if (V == 7) {   // 32-byte load (whole sector) as two 16-byte halves
            uint4 a = __ldg(buf + e), b = __ldg(buf + e + 1);
            v = make_uint4(a.x ^ b.x, a.y ^ b.y, a.z, a.w);
        } else if (V == 8) {   // 256-bit load
            uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
            asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "l"(buf + e));
            v = make_uint4(r0 ^ r4, r1 ^ r5, r2 ^ r6, r3 ^ r7);
        } else {
            v = __ldg(buf + e);
        }
        acc ^= v.y;
        s = mix64(s + v.x);
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

__global__ void k_fill(uint4 *buf, uint64_t n) { for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) { uint32_t x = (uint32_t)mix64(i); buf[i] = make_uint4(x, x ^ 0x5bd1e995u, (uint32_t)i, 0); } }

template <int V> float run(const uint4 *buf, cudaTextureObject_t tex, uint64_t n_sectors, uint32_t per, uint32_t *sink, int grid)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(a));
        k_probe<V><<<grid, 256>>>(buf, tex, n_sectors, per, sink);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (r && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char **argv)
{
    size_t lim = 0;
    if (argc > 1) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1])); printf("set limit %s -> %s\n", argv[1], cudaGetErrorString(e)); }
    CK(cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity));
    printf("cudaLimitMaxL2FetchGranularity = %zu\n", lim);
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const uint64_t bytes = (argc > 2) ? (uint64_t)atoll(argv[2]) << 20 : (1ull << 30);
    const uint64_t n_sectors = bytes / 32;
    uint4 *buf; uint32_t *sink;
    CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&sink, 4));
    const int grid = prop.multiProcessorCount * 8;
    k_fill<<<grid, 256>>>(buf, n_sectors * 2);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = buf; rd.res.linear.desc = cudaCreateChannelDesc<uint4>(); rd.res.linear.sizeInBytes = bytes;
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    const uint32_t per = 256;
    const double loads = (double)grid * 256 * per;
    const char *names[] = {"ldg.nc", "tex1Dfetch", "cp.async.cg", "ld.cg", "hint evict_first", "hint evict_last", "hint evict_unchanged", "2x16B same sector", "ld.v8 (256-bit)"};
    float t;
    t = run<0>(buf, tex, n_sectors, per, sink, grid); printf("%-22s %.1f G loads/s\n", names[0], loads / t / 1e6);
    t = run<1>(buf, tex, n_sectors, per, sink, grid); printf("%-22s %.1f G loads/s\n", names[1], loads / t / 1e6);
    t = run<2>(buf, tex, n_sectors, per, sink, grid); printf("%-22s %.1f G loads/s\n", names[2], loads / t / 1e6);
    t = run<3>(buf, tex, n_sectors, per, sink, grid); printf("%-22s %.1f G loads/s\n", names[3], loads / t / 1e6);
    t = run<4>(buf, tex, n_sectors, per, sink, grid); printf("%-22s %.1f G loads/s\n", names[4], loads / t / 1e6);
    t = run<5>(buf, tex, n_sectors, per, sink, grid); printf("%-22s %.1f G loads/s\n", names[5], loads / t / 1e6);
    t = run<6>(buf, tex, n_sectors, per, sink, grid); printf("%-22s %.1f G loads/s\n", names[6], loads / t / 1e6);
    t = run<7>(buf, tex, n_sectors, per, sink, grid); printf("%-22s %.1f G loads/s\n", names[7], loads / t / 1e6);
    t = run<8>(buf, tex, n_sectors, per, sink, grid); printf("%-22s %.1f G loads/s\n", names[8], loads / t / 1e6);
    return 0;
}
