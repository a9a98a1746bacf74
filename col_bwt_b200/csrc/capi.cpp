// capi.cpp -- C-ABI entry points that do not launch kernels themselves: index loading (file / memory), stats,
// the pml_query text formatter, error reporting.  See include/colbwt_b200.h for the reference interface each
// one replaces.
#include <sys/stat.h>

#include <algorithm>
#include <cstring>
#include <memory>

#include "internal.h"

namespace colbwt {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int device_list(const int *devices, int n_devices, std::vector<int> &out)
{
    int have = 0;
    cudaError_t e = cudaGetDeviceCount(&have);
    if (e != cudaSuccess || have == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); this library has no CPU path", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return COLBWT_ERR_CUDA;
    }
    if (n_devices < 1) {
        set_error("n_devices must be >= 1");
        return COLBWT_ERR_ARG;
    }
    for (int i = 0; i < n_devices; ++i) {
        const int d = devices ? devices[i] : i;
        if (d < 0 || d >= have) {
            set_error("CUDA device %d requested but only %d present", d, have);
            return COLBWT_ERR_ARG;
        }
        out.push_back(d);
    }
    return COLBWT_OK;
}

static int build_index(const void *rows, FILE *fp, long data_start, uint64_t bwt_r, uint64_t n, uint64_t r,
                       const std::vector<int> &devs, colbwt_index **out)
{
    if (r == 0 || r >= (1ull << 32) || n == 0 || n >= (1ull << 40)) {
        set_error("inconsistent header: n=%llu r=%llu (need 0 < r < 2^32, 0 < n < 2^40)", (unsigned long long)n, (unsigned long long)r);
        return COLBWT_ERR_FORMAT;
    }
    std::unique_ptr<colbwt_index> idx(new colbwt_index);
    idx->stats.n = n;
    idx->stats.r = r;
    idx->stats.bwt_r = bwt_r;
    idx->stats.n_devices = (int)devs.size();
    idx->dev.resize(devs.size());
    for (size_t i = 0; i < devs.size(); ++i) {
        if (fp) fseek(fp, data_start, SEEK_SET);
        int rc = build_device_table(idx->dev[i], devs[i], rows, fp, n, r, i == 0 ? &idx->stats : nullptr, idx->code_lut);
        if (rc != COLBWT_OK) {
            for (auto &d : idx->dev) free_device_table(d);
            return rc;
        }
    }
    *out = idx.release();
    return COLBWT_OK;
}

static bool read_file(const std::string &path, std::vector<uint8_t> &out)
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

static void unpack_u40(const std::vector<uint8_t> &raw, std::vector<uint64_t> &out)
{
    out.resize(raw.size() / 5);
    for (size_t i = 0; i < out.size(); ++i) {
        uint64_t v = 0;
        memcpy(&v, raw.data() + 5 * i, 5);   // little endian, RW_BYTES = 5 (common.hpp:46)
        out[i] = v;
    }
}

// prefix = the reference's `<out>.fa` stem (build_col_bwt.cpp:17-33 forms the same five names)
static int load_primaries(const std::string &prefix, Primaries &pr)
{
    std::vector<uint8_t> raw;
    if (!read_file(prefix + ".bwt.heads", pr.heads) || pr.heads.empty()) {
        set_error("cannot read %s.bwt.heads", prefix.c_str());
        return COLBWT_ERR_IO;
    }
    if (!read_file(prefix + ".bwt.len", raw)) {
        set_error("cannot read %s.bwt.len", prefix.c_str());
        return COLBWT_ERR_IO;
    }
    unpack_u40(raw, pr.lens);
    if (!read_file(prefix + ".thr_pos", raw)) {
        set_error("cannot read %s.thr_pos", prefix.c_str());
        return COLBWT_ERR_IO;
    }
    unpack_u40(raw, pr.thr);
    if (pr.lens.size() != pr.heads.size() || pr.thr.size() != pr.heads.size()) {
        set_error("%s: %zu run heads but %zu lengths and %zu thresholds", prefix.c_str(), pr.heads.size(), pr.lens.size(), pr.thr.size());
        return COLBWT_ERR_FORMAT;
    }
    if (!read_file(prefix + ".col_runs", raw) || raw.size() < 8) {
        set_error("cannot read %s.col_runs", prefix.c_str());
        return COLBWT_ERR_IO;
    }
    memcpy(&pr.n_bits, raw.data(), 8);   // sdsl bit_vector: u64 length in bits, then 64-bit words (col_split.hpp:384-386)
    const uint64_t nw = (pr.n_bits + 63) / 64;
    if (raw.size() < 8 + nw * 8) {
        set_error("%s.col_runs: %zu bytes cannot hold a bit_vector of %llu bits (the sd_vector form is not supported)", prefix.c_str(),
                  raw.size(), (unsigned long long)pr.n_bits);
        return COLBWT_ERR_FORMAT;
    }
    pr.bits.resize(nw);
    memcpy(pr.bits.data(), raw.data() + 8, nw * 8);
    if (pr.n_bits & 63) pr.bits[nw - 1] &= (1ull << (pr.n_bits & 63)) - 1;
    if (!read_file(prefix + ".col_ids", pr.ids)) {
        set_error("cannot read %s.col_ids", prefix.c_str());
        return COLBWT_ERR_IO;
    }
    uint64_t total = 0;
    for (uint64_t l : pr.lens) {
        if (l == 0) {
            set_error("%s.bwt.len: a run of length 0", prefix.c_str());
            return COLBWT_ERR_FORMAT;
        }
        total += l;
    }
    if (total != pr.n_bits) {
        set_error("%s: run lengths sum to %llu but .col_runs has %llu bits", prefix.c_str(), (unsigned long long)total, (unsigned long long)pr.n_bits);
        return COLBWT_ERR_FORMAT;
    }
    for (size_t i = 1; i < pr.heads.size(); ++i) {
        const uint8_t a = pr.heads[i - 1], b = pr.heads[i];
        const uint8_t ma = (a <= 1 || a >= 128) ? 1 : a, mb = (b <= 1 || b >= 128) ? 1 : b;
        if (ma == mb) {   // read_thresholds (col_bwt.hpp:450-453) would give both runs one threshold and drift out of step
            set_error("%s.bwt.heads: runs %zu and %zu carry the same character after terminator folding", prefix.c_str(), i - 1, i);
            return COLBWT_ERR_FORMAT;
        }
    }
    return COLBWT_OK;
}

} // namespace colbwt

using namespace colbwt;

extern "C" const char *colbwt_last_error(void) { return g_err; }
extern "C" const char *colbwt_version(void) { return "col_bwt_b200 0.1 (sm_100a)"; }

extern "C" int colbwt_index_load(const char *path, const int *devices, int n_devices, colbwt_index **out)
{
    if (!path || !out) {
        set_error("colbwt_index_load: null argument");
        return COLBWT_ERR_ARG;
    }
    std::string p(path);
    struct stat sb;
    if (stat(p.c_str(), &sb) != 0 || !S_ISREG(sb.st_mode)) p += ".col_pml";   // pml_query.cpp:110-111
    FILE *fp = fopen(p.c_str(), "rb");
    if (!fp) {
        set_error("cannot open %s", p.c_str());
        return COLBWT_ERR_IO;
    }
    std::unique_ptr<FILE, int (*)(FILE *)> guard(fp, fclose);
    uint64_t hdr[4];   // bwt_r (col_bwt.hpp:377), n, r, size (LF_table.hpp:351-354)
    if (fread(hdr, 8, 4, fp) != 4) {
        set_error("%s: shorter than the 32-byte header", p.c_str());
        return COLBWT_ERR_IO;
    }
    if (hdr[2] != hdr[3]) {
        set_error("%s: header says r=%llu but stores %llu rows", p.c_str(), (unsigned long long)hdr[2], (unsigned long long)hdr[3]);
        return COLBWT_ERR_FORMAT;
    }
    if (fstat(fileno(fp), &sb) == 0 && (uint64_t)sb.st_size < 32 + hdr[3] * 18) {
        set_error("%s: %llu bytes, but %llu rows need %llu", p.c_str(), (unsigned long long)sb.st_size, (unsigned long long)hdr[3],
                  (unsigned long long)(32 + hdr[3] * 18));
        return COLBWT_ERR_IO;
    }
    std::vector<int> devs;
    if (int rc = device_list(devices, n_devices, devs)) return rc;
    return build_index(nullptr, fp, 32, hdr[0], hdr[1], hdr[2], devs, out);
}

extern "C" int colbwt_index_from_rows(const void *rows, uint64_t bwt_r, uint64_t n, uint64_t r, const int *devices,
                                      int n_devices, colbwt_index **out)
{
    if (!rows || !out) {
        set_error("colbwt_index_from_rows: null argument");
        return COLBWT_ERR_ARG;
    }
    std::vector<int> devs;
    if (int rc = device_list(devices, n_devices, devs)) return rc;
    return build_index(rows, nullptr, 0, bwt_r, n, r, devs, out);
}

extern "C" int colbwt_index_from_primaries(const char *prefix, const int *devices, int n_devices, colbwt_index **out)
{
    if (!prefix || !out) {
        set_error("colbwt_index_from_primaries: null argument");
        return COLBWT_ERR_ARG;
    }
    Primaries pr;
    if (int rc = load_primaries(prefix, pr)) return rc;
    std::vector<int> devs;
    if (int rc = device_list(devices, n_devices, devs)) return rc;
    std::unique_ptr<colbwt_index> idx(new colbwt_index);
    idx->dev.resize(devs.size());
    idx->stats.bwt_r = pr.heads.size();
    idx->stats.n_devices = (int)devs.size();
    for (size_t i = 0; i < devs.size(); ++i) {
        uint64_t n = 0, r = 0;
        int rc = build_device_table_from_primaries(idx->dev[i], devs[i], pr, i == 0 ? &idx->stats : nullptr, idx->code_lut, &n, &r);
        if (rc != COLBWT_OK) {
            for (auto &d : idx->dev) free_device_table(d);
            return rc;
        }
        idx->stats.n = n;
        idx->stats.r = r;
    }
    *out = idx.release();
    return COLBWT_OK;
}

extern "C" int colbwt_index_save(const colbwt_index *idx, const char *path)
{
    if (!idx || !path || idx->dev.empty()) {
        set_error("colbwt_index_save: null argument");
        return COLBWT_ERR_ARG;
    }
    FILE *f = fopen(path, "wb");
    if (!f) {
        set_error("cannot write %s", path);
        return COLBWT_ERR_IO;
    }
    std::unique_ptr<FILE, int (*)(FILE *)> guard(f, fclose);
    const uint64_t hdr[4] = {idx->stats.bwt_r, idx->stats.n, idx->stats.r, idx->stats.r};   // col_bwt.hpp:364, LF_table.hpp:329-337
    if (fwrite(hdr, 8, 4, f) != 4) {
        set_error("short write to %s", path);
        return COLBWT_ERR_IO;
    }
    const uint64_t chunk = 4u << 20;
    std::vector<uint8_t> buf(chunk * 18);
    for (uint64_t first = 0; first < idx->stats.r; first += chunk) {
        const uint64_t count = std::min<uint64_t>(chunk, idx->stats.r - first);
        if (int rc = export_rows(idx->dev[0], first, count, buf.data())) return rc;
        if (fwrite(buf.data(), 18, count, f) != count) {
            set_error("short write to %s", path);
            return COLBWT_ERR_IO;
        }
    }
    return COLBWT_OK;
}

extern "C" int colbwt_index_stats(const colbwt_index *idx, colbwt_stats *out)
{
    if (!idx || !out) {
        set_error("colbwt_index_stats: null argument");
        return COLBWT_ERR_ARG;
    }
    *out = idx->stats;
    return COLBWT_OK;
}

extern "C" void colbwt_index_free(colbwt_index *idx)
{
    if (!idx) return;
    destroy_pipeline(idx->pipeline);
    for (auto &d : idx->dev) free_device_table(d);
    delete idx;
}

// pml_query.cpp:79-85: fs << '>' << id << " \n"; std::copy(..., ostream_iterator<size_t>(fs, " ")); fs << "\n";
extern "C" size_t colbwt_format_stats(char *buf, size_t cap, const char *id, size_t id_len, const void *values, int width, uint64_t m)
{
    auto value_at = [&](uint64_t i) -> uint32_t {
        return width == 1 ? ((const uint8_t *)values)[i] : width == 2 ? ((const uint16_t *)values)[i] : ((const uint32_t *)values)[i];
    };
    // a buffer that holds the worst case (3 / 5 / 10 digits + blank per value) needs no counting pass
    const size_t worst = 1 + id_len + 2 + 1 + (size_t)m * (width == 1 ? 4 : width == 2 ? 6 : 11);
    if (!buf || cap < worst) {
        size_t need = 1 + id_len + 2 + 1;
        for (uint64_t i = 0; i < m; ++i) {
            uint32_t v = value_at(i);
            need += (v < 10 ? 1 : v < 100 ? 2 : v < 1000 ? 3 : v < 10000 ? 4 : v < 100000 ? 5 : v < 1000000 ? 6 : v < 10000000 ? 7 : v < 100000000 ? 8 : v < 1000000000 ? 9 : 10) + 1;
        }
        if (!buf || cap < need) return need;
    }
    char *p = buf;
    *p++ = '>';
    memcpy(p, id, id_len);
    p += id_len;
    *p++ = ' ';
    *p++ = '\n';
    for (uint64_t i = 0; i < m; ++i) {
        uint32_t v = value_at(i);
        char tmp[12];
        int k = 0;
        do {
            tmp[k++] = (char)('0' + v % 10);
            v /= 10;
        } while (v);
        while (k) *p++ = tmp[--k];
        *p++ = ' ';
    }
    *p++ = '\n';
    return (size_t)(p - buf);
}
