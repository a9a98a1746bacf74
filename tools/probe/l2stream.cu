// Scratch probe: does a per-thread output stream (16 B every 8 gathers, like the PML/CID staging of k_traverse) hurt the
// L2 hit rate of random gathers over a buffer that partly fits the L2?  And which store flavour avoids it?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

// ST: 0 none, 1 plain st.v4, 2 st.cs, 3 st.cg(wt?), 4 L1::no_allocate + L2 evict_first hint, 5 st.wt
template <int ST, int CTAS>
__global__ void __launch_bounds__(256, CTAS) k(const uint4 *__restrict__ buf, uint64_t n_rows, uint32_t per_thread, uint4 *out, uint64_t out_stride, uint32_t *sink)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s = mix64(tid + 0x9E3779B97F4A7C15ULL);
    uint4 *o = out + tid * out_stride;
    uint32_t acc = 0;
    uint64_t pol = 0;
    if (ST == 4) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (uint32_t i = 0; i < per_thread; ++i) {
        const uint4 v = __ldg(buf + (s % n_rows));
        acc ^= v.y;
        s = mix64(s + v.x);
        if (ST >= 1 && ST <= 5 && (i & 7) == 7) {
            uint4 w = make_uint4(acc, i, v.z, v.w);
            uint4 *p = o + (i >> 3);
            if (ST == 1) *p = w;
            else if (ST == 2) __stcs(p, w);
            else if (ST == 3) __stcg(p, w);
            else if (ST == 4) asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w), "l"(pol) : "memory");
            else if (ST == 5) __stwt(p, w);
        }
        if (ST == 6 && (i & 31) == 31) {
            uint4 *p = o + (i >> 5) * 4;
            uint4 w = make_uint4(acc, i, v.z, v.w);
            p[0] = w; p[1] = w; p[2] = w; p[3] = w;
        }
        if (ST == 7 && (i & 63) == 63) {
            uint4 *p = o + (i >> 6) * 8;
            uint4 w = make_uint4(acc, i, v.z, v.w);
#pragma unroll
            for (int q = 0; q < 8; ++q) p[q] = w;
        }
        if (ST == 8 && (i & 3) == 3) {
            uint2 *p = reinterpret_cast<uint2 *>(o) + (i >> 2);
            *p = make_uint2(acc, i);
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}
__global__ void k_fill(uint4 *buf, uint64_t n) { for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) { uint32_t x = (uint32_t)mix64(i); buf[i] = make_uint4(x, x ^ 0x5bd1e995u, (uint32_t)i, 0); } }

template <int ST, int CTAS> void run(const char *name, const uint4 *buf, uint64_t n_rows, uint4 *out, uint32_t *sink, int sms)
{
    const uint32_t per = 1024;
    const int grid = sms * CTAS;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(a));
        k<ST, CTAS><<<grid, 256>>>(buf, n_rows, per, out, per / 8, sink);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (r && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    printf("%-34s CTAs/SM %d: %.1f G gathers/s\n", name, CTAS, (double)grid * 256 * per / best / 1e6);
}

int main(int argc, char **argv)
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const uint64_t bytes = (argc > 1 ? (uint64_t)atoll(argv[1]) : 366) << 20;
    const uint64_t n_rows = bytes / 16;
    uint4 *buf, *out; uint32_t *sink;
    CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMalloc(&out, (size_t)prop.multiProcessorCount * 8 * 256 * 128 * 16));
    k_fill<<<prop.multiProcessorCount * 8, 256>>>(buf, n_rows);
    printf("buffer %llu MiB\n", (unsigned long long)(bytes >> 20));
    run<0, 4>("no stores", buf, n_rows, out, sink, prop.multiProcessorCount);
    run<1, 4>("st.v4 every 8 gathers", buf, n_rows, out, sink, prop.multiProcessorCount);
    run<2, 4>("st.cs", buf, n_rows, out, sink, prop.multiProcessorCount);
    run<3, 4>("st.cg", buf, n_rows, out, sink, prop.multiProcessorCount);
    run<4, 4>("st no_allocate + L2 evict_first", buf, n_rows, out, sink, prop.multiProcessorCount);
    run<5, 4>("st.wt", buf, n_rows, out, sink, prop.multiProcessorCount);
    run<8, 4>("8 B every 4 gathers", buf, n_rows, out, sink, prop.multiProcessorCount);
    run<6, 4>("64 B every 32 gathers", buf, n_rows, out, sink, prop.multiProcessorCount);
    run<7, 4>("128 B every 64 gathers", buf, n_rows, out, sink, prop.multiProcessorCount);
    run<0, 8>("no stores", buf, n_rows, out, sink, prop.multiProcessorCount);
    run<1, 8>("st.v4 every 8 gathers", buf, n_rows, out, sink, prop.multiProcessorCount);
    run<2, 8>("st.cs", buf, n_rows, out, sink, prop.multiProcessorCount);
    return 0;
}
