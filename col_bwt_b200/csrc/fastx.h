// fastx.h -- FASTA/FASTQ(.gz) batch reader for the pml_query-compatible CLI.
//
// Replaces PatternProcessor (include/common/io.hpp:6-35), which wraps klib's kseq: a record starts at '>' or '@';
// the id is the header up to the first whitespace (io.hpp:24-26); the sequence is every following line, joined,
// bytes verbatim (no case folding, io.hpp:17-22), until a line starting with '>', '@' or '+'; after '+' as many
// quality bytes as sequence bytes are skipped.  Instead of one record per call, records are appended to one
// concatenated buffer + offsets, which is the layout colbwt_query takes.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace colbwt {

class FastxReader {
public:
    explicit FastxReader(const std::string &path) : fp_(gzopen(path.c_str(), "r")), buf_(1 << 20) {
        if (fp_) gzbuffer(fp_, 1 << 20);
    }
    ~FastxReader() { if (fp_) gzclose(fp_); }
    bool ok() const { return fp_ != nullptr; }

    // Appends records until `max_bases` sequence bytes or `max_reads` records have been added (at least one record
    // if any is left).  seqs/off/ids are cleared first; off gets n+1 entries.  Returns the number of records.
    size_t next_batch(std::vector<uint8_t> &seqs, std::vector<uint64_t> &off, std::vector<std::string> &ids,
                      uint64_t max_bases, uint64_t max_reads)
    {
        seqs.clear();
        off.assign(1, 0);
        ids.clear();
        while (ids.size() < max_reads && seqs.size() < max_bases) {
            if (!read_record(seqs, ids)) break;
            off.push_back(seqs.size());
        }
        return ids.size();
    }

private:
    int getc_()
    {
        if (beg_ >= end_) {
            if (eof_) return -1;
            end_ = gzread(fp_, buf_.data(), (unsigned)buf_.size());
            beg_ = 0;
            if (end_ <= 0) { eof_ = true; end_ = 0; return -1; }
        }
        return (unsigned char)buf_[beg_++];
    }
    static bool is_space(int c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

    // Appends the rest of the current line (without the newline) to `out` using memchr over the buffer; returns the
    // terminating character ('\n', or -1 at end of file).
    int rest_of_line(std::vector<uint8_t> &out)
    {
        for (;;) {
            if (beg_ >= end_) {
                if (eof_) return -1;
                end_ = gzread(fp_, buf_.data(), (unsigned)buf_.size());
                beg_ = 0;
                if (end_ <= 0) { eof_ = true; end_ = 0; return -1; }
            }
            const char *p = buf_.data() + beg_;
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end_ - beg_));
            const size_t k = nl ? (size_t)(nl - p) : (size_t)(end_ - beg_);
            out.insert(out.end(), (const uint8_t *)p, (const uint8_t *)p + k);
            beg_ += (int)k + (nl ? 1 : 0);
            if (nl) return '\n';
        }
    }

    bool read_record(std::vector<uint8_t> &seqs, std::vector<std::string> &ids)
    {
        int c;
        if (last_ == 0) {   // find the next header
            while ((c = getc_()) != -1 && c != '>' && c != '@') {}
            if (c == -1) return false;
            last_ = c;
        }
        std::string id;
        while ((c = getc_()) != -1 && !is_space(c)) id.push_back((char)c);
        if (c == -1) return false;   // kseq: a header cut off by EOF is not a record
        if (c != '\n') { scratch_.clear(); rest_of_line(scratch_); }   // comment
        const size_t start = seqs.size();
        while ((c = getc_()) != -1 && c != '>' && c != '+' && c != '@') {
            if (c == '\n') continue;
            const size_t line = seqs.size();
            seqs.push_back((uint8_t)c);
            rest_of_line(seqs);
            if (seqs.size() - line > 1 && seqs.back() == '\r') seqs.pop_back();
        }
        if (c == '>' || c == '@') last_ = c;
        else last_ = 0;
        if (c == '+') {   // FASTQ: skip the rest of the '+' line, then as many quality bytes as bases
            scratch_.clear();
            rest_of_line(scratch_);
            size_t q = 0, want = seqs.size() - start;
            while (q < want) {
                scratch_.clear();
                const int e = rest_of_line(scratch_);
                size_t n = scratch_.size();
                if (n && scratch_.back() == '\r') --n;
                q += n;
                if (e == -1) break;
            }
        }
        ids.push_back(std::move(id));
        return true;
    }

    std::vector<uint8_t> scratch_;
    gzFile fp_;
    std::vector<char> buf_;
    int beg_ = 0, end_ = 0, last_ = 0;
    bool eof_ = false;
};

} // namespace colbwt
