// rlbwt.cu -- run-length BWT -> plain BWT on the GPU.
//
// Replaces src/rlbwt_to_bwt.cpp:8-34: PREFIX.bwt.heads (1 byte per run) + PREFIX.bwt.len (5-byte LE per run) ->
// PREFIX.bwt, every head byte repeated `len` times, heads written as they are (no terminator folding), runs taken
// while BOTH files still deliver a record (rlbwt_to_bwt.cpp:25), zero-length runs contributing nothing.
//
// The run table goes to the device once (6 bytes per run); the text is produced in chunks: each thread writes one
// 16-byte vector, finds the run of its first byte with a binary search over the run starts and walks forward from
// there; chunks are copied back into two pinned buffers and written while the next one is expanded.
#include <cub/cub.cuh>

#include <algorithm>
#include <memory>

#include "internal.h"

namespace colbwt {

__global__ void k_unpack_len(const uint8_t *__restrict__ len5, uint64_t r, uint64_t *len)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r) return;
    uint64_t v = 0;
    for (int b = 0; b < 5; ++b) v |= (uint64_t)len5[i * 5 + b] << (8 * b);
    len[i] = v;
}

// start[i] = first text position of run i; start[r] = n.  Output bytes [p0, p0 + count) -> out[0, count).
__global__ void k_expand_runs(const uint8_t *__restrict__ heads, const uint64_t *__restrict__ start, uint64_t r, uint64_t p0, uint64_t count,
                              uint8_t *__restrict__ out)
{
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // 16-byte vector number
    if (v * 16 >= count) return;
    const uint64_t p = p0 + v * 16;
    // last run whose start is <= p (zero-length runs share their successor's start and are skipped by taking the last)
    uint64_t lo = 0, hi = r;   // invariant: start[lo] <= p < start[hi]
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (start[mid] <= p) lo = mid; else hi = mid;
    }
    uint64_t run = lo, next = start[run + 1];
    uint8_t c = heads[run];
    const uint64_t e = min((uint64_t)16, count - v * 16);
    if (p + 16 <= next && e == 16) {   // whole vector inside one run: the common case
        const uint32_t w = c * 0x01010101u;
        reinterpret_cast<uint4 *>(out)[v] = make_uint4(w, w, w, w);
        return;
    }
    for (uint64_t k = 0; k < e; ++k) {
        while (p + k >= next) {
            ++run;
            next = start[run + 1];
            c = heads[run];
        }
        out[v * 16 + k] = c;
    }
}

static bool slurp(const std::string &path, std::vector<uint8_t> &out)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(sz > 0 ? (size_t)sz : 0);
    const bool ok = out.empty() || fread(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

struct RlbwtScratch {
    void *d_heads = nullptr, *d_len5 = nullptr, *d_len = nullptr, *d_start = nullptr, *d_tmp = nullptr, *d_out[2] = {nullptr, nullptr};
    uint8_t *h_out[2] = {nullptr, nullptr};
    cudaStream_t stream[2] = {nullptr, nullptr};
    FILE *fp = nullptr;
    ~RlbwtScratch()
    {
        for (void *p : {d_heads, d_len5, d_len, d_start, d_tmp, d_out[0], d_out[1]}) cudaFree(p);
        for (uint8_t *p : h_out) cudaFreeHost(p);
        for (cudaStream_t s : stream) if (s) cudaStreamDestroy(s);
        if (fp) fclose(fp);
    }
};

} // namespace colbwt

extern "C" int colbwt_rlbwt_to_bwt(const char *prefix, int device, uint64_t *n_out)
{
    using namespace colbwt;
    if (!prefix) {
        set_error("colbwt_rlbwt_to_bwt: null prefix");
        return COLBWT_ERR_ARG;
    }
    const std::string p(prefix);
    std::vector<uint8_t> heads, len5;
    if (!slurp(p + ".bwt.heads", heads)) {
        set_error("cannot read %s.bwt.heads", prefix);
        return COLBWT_ERR_IO;
    }
    if (!slurp(p + ".bwt.len", len5)) {
        set_error("cannot read %s.bwt.len", prefix);
        return COLBWT_ERR_IO;
    }
    const uint64_t r = std::min<uint64_t>(heads.size(), len5.size() / 5);   // the reference stops at the shorter file
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
        cudaGetLastError();
        set_error("CUDA device %d not available; this library has no CPU path", device);
        return COLBWT_ERR_CUDA;
    }
    CB_CUDA(cudaSetDevice(device));
    RlbwtScratch sc;
    sc.fp = fopen((p + ".bwt").c_str(), "wb");
    if (!sc.fp) {
        set_error("cannot write %s.bwt", prefix);
        return COLBWT_ERR_IO;
    }
    uint64_t n = 0;
    if (r) {
        CB_CUDA(cudaMalloc(&sc.d_heads, r));
        CB_CUDA(cudaMalloc(&sc.d_len5, r * 5));
        CB_CUDA(cudaMalloc(&sc.d_len, (r + 1) * 8));
        CB_CUDA(cudaMalloc(&sc.d_start, (r + 1) * 8));
        CB_CUDA(cudaMemcpy(sc.d_heads, heads.data(), r, cudaMemcpyHostToDevice));
        CB_CUDA(cudaMemcpy(sc.d_len5, len5.data(), r * 5, cudaMemcpyHostToDevice));
        CB_CUDA(cudaMemset(sc.d_len, 0, (r + 1) * 8));
        k_unpack_len<<<(unsigned)((r + 255) / 256), 256>>>((const uint8_t *)sc.d_len5, r, (uint64_t *)sc.d_len);
        CB_CUDA(cudaGetLastError());
        size_t tmp_bytes = 0;
        CB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const uint64_t *)sc.d_len, (uint64_t *)sc.d_start, r + 1));
        CB_CUDA(cudaMalloc(&sc.d_tmp, tmp_bytes));
        CB_CUDA(cub::DeviceScan::ExclusiveSum(sc.d_tmp, tmp_bytes, (const uint64_t *)sc.d_len, (uint64_t *)sc.d_start, r + 1));
        CB_CUDA(cudaMemcpy(&n, (const uint64_t *)sc.d_start + r, 8, cudaMemcpyDeviceToHost));
    }
    const uint64_t chunk = std::min<uint64_t>(256ull << 20, (n + 15) & ~15ull);
    if (n) {
        for (int s = 0; s < 2; ++s) {
            CB_CUDA(cudaMalloc(&sc.d_out[s], chunk));
            CB_CUDA(cudaMallocHost(&sc.h_out[s], chunk));
            CB_CUDA(cudaStreamCreateWithFlags(&sc.stream[s], cudaStreamNonBlocking));
        }
    }
    uint64_t pending[2] = {0, 0};
    auto drain = [&](int s) -> int {
        if (!pending[s]) return COLBWT_OK;
        CB_CUDA(cudaStreamSynchronize(sc.stream[s]));
        if (fwrite(sc.h_out[s], 1, pending[s], sc.fp) != pending[s]) {
            set_error("short write to %s.bwt", prefix);
            return COLBWT_ERR_IO;
        }
        pending[s] = 0;
        return COLBWT_OK;
    };
    int s = 0;
    for (uint64_t p0 = 0; p0 < n; p0 += chunk, s ^= 1) {
        if (int rc = drain(s)) return rc;
        const uint64_t count = std::min(chunk, n - p0);
        const uint64_t vecs = (count + 15) / 16;
        k_expand_runs<<<(unsigned)((vecs + 255) / 256), 256, 0, sc.stream[s]>>>((const uint8_t *)sc.d_heads, (const uint64_t *)sc.d_start, r, p0, count,
                                                                                 (uint8_t *)sc.d_out[s]);
        CB_CUDA(cudaGetLastError());
        CB_CUDA(cudaMemcpyAsync(sc.h_out[s], sc.d_out[s], count, cudaMemcpyDeviceToHost, sc.stream[s]));
        pending[s] = count;
    }
    if (int rc = drain(s)) return rc;
    if (int rc = drain(s ^ 1)) return rc;
    if (fclose(sc.fp) != 0) {
        sc.fp = nullptr;
        set_error("cannot finish %s.bwt", prefix);
        return COLBWT_ERR_IO;
    }
    sc.fp = nullptr;
    if (n_out) *n_out = n;
    return COLBWT_OK;
}
