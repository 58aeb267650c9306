// fastx.hpp -- FASTA/FASTQ reader with the record semantics of the reference's kseq.h (kseq_read, kseq.h:170-208),
// over zlib's gzread (plain or gzipped input):
//   * a header starts at '>' or '@'; name = up to the first white-space character, comment = rest of that line;
//   * the sequence runs until the next '>', '+' or '@' byte (anywhere, as in kseq) and keeps only isgraph() bytes;
//   * after '+' the rest of the line is skipped and as many quality characters (33..127) as sequence bytes are read;
//   * return value = sequence length, -1 at end of file, -2 for a truncated quality string.
// Implementation differs (block-wise scanning with a byte-class table instead of one ks_getc call per byte).
#pragma once
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "seqbuf.hpp"

namespace mashhost {

class FastxReader {
public:
    std::string name, comment;
    SeqBuffer seq;          // sequence of the current record (pool storage, seqbuf.hpp); move it out to keep it

    explicit FastxReader(gzFile f) : f_(f), buf_(1 << 18)
    {
        for (int c = 0; c < 256; c++) cls_[c] = (c == '>' || c == '+' || c == '@') ? 2 : (isgraph(c) ? 0 : 1);
    }
    ~FastxReader() {}

    // Size of the (uncompressed or compressed) input, when known: the first record's sequence buffer is reserved up front and
    // later ones at twice the previous record's length, so that a chromosome-sized record is not copied log2(n) times while its
    // string grows (half of the reader's time on 70-column FASTA before this).
    void setSizeHint(size_t bytes) { remaining_ = bytes; }

    static gzFile openPath(const std::string &path)
    {
        if (path == "-") return gzdopen(fileno(stdin), "r");
        return gzopen(path.c_str(), "r");
    }

    int read()
    {
        int c;
        if (lastChar_ == 0) {   // jump to the next header line
            while ((c = getc()) != -1 && c != '>' && c != '@') {}
            if (c == -1) return -1;
            lastChar_ = c;
        }
        comment.clear(); seq.clear();
        {
            const size_t want = lastLen_ ? std::min(remaining_, 2 * lastLen_ + 64) : remaining_;
            if (want > seq.capacity()) seq.reserve(want);
        }
        size_t qual = 0;
        if (!getUntilSpace(name, c)) return -1;
        if (c != '\n') getUntilNewline(comment);
        // sequence: append graph bytes until a terminator
        for (;;) {
            if (begin_ >= end_ && !fill()) { c = -1; break; }
            const unsigned char *p = buf_.data() + begin_, *e = buf_.data() + end_;
            const unsigned char *run = p;
            int stop = 0;
            while (p < e) {
                p = skipGraph(p, e);         // next byte that is not a plain sequence byte (class != 0), 32 bytes at a time
                if (p >= e) break;
                int k = cls_[*p];
                if (k == 0) { p++; continue; }
                if (p > run) seq.append(reinterpret_cast<const char *>(run), p - run);
                if (k == 2) { stop = *p; p++; break; }
                p++;
                run = p;
            }
            if (!stop && p > run) seq.append(reinterpret_cast<const char *>(run), p - run);
            begin_ = p - buf_.data();
            if (stop) { c = stop; break; }
        }
        if (c == '>' || c == '@') lastChar_ = c;
        lastLen_ = seq.size();
        if (c != '+') return (int)seq.size();   // FASTA
        // rest of the '+' line (block-wise memchr instead of one getc per byte)
        for (c = -1;;) {
            if (begin_ >= end_ && !fill()) break;
            const unsigned char *nlp = static_cast<const unsigned char *>(memchr(buf_.data() + begin_, '\n', end_ - begin_));
            if (nlp) { begin_ = (size_t)(nlp - buf_.data()) + 1; c = '\n'; break; }
            begin_ = end_;
        }
        if (c == -1) return -2;
        // quality: kseq reads one byte past the last quality character it needs (`while ((c = ks_getc(ks)) != -1 && qual.l < seq.l)`
        // consumes a byte, then tests the length), so one further byte is taken after the count is complete
        for (;;) {
            if (begin_ >= end_ && !fill()) break;
            if (qual >= seq.size()) { begin_++; break; }                 // the byte kseq consumes before it notices it is done
            const unsigned char *q = buf_.data() + begin_, *qe = buf_.data() + end_;
            size_t need = seq.size() - qual;
            while (q < qe && need) { need -= (size_t)(*q >= 33 && *q <= 127); q++; }
            qual = seq.size() - need;
            begin_ = (size_t)(q - buf_.data());
        }
        lastChar_ = 0;
        if (seq.size() != qual) return -2;
        return (int)seq.size();
    }

private:
    gzFile f_;
    std::vector<unsigned char> buf_;
    size_t begin_ = 0, end_ = 0;
    bool eof_ = false;
    int lastChar_ = 0;
    size_t remaining_ = 0, lastLen_ = 0;     // input bytes not yet read (hint), length of the previous record
    unsigned char cls_[256];

    // First position in [p, e) whose byte is not a plain sequence byte: outside isgraph() (33..126) or one of '>' '+' '@'.
#if defined(__x86_64__)
    __attribute__((target("avx2"))) static const unsigned char *skipGraphAvx2(const unsigned char *p, const unsigned char *e)
    {
        const __m256i lo = _mm256_set1_epi8(33), flip = _mm256_set1_epi8((char)0x80), lim = _mm256_set1_epi8((char)(93 ^ 0x80));
        const __m256i gt = _mm256_set1_epi8('>'), pl = _mm256_set1_epi8('+'), at = _mm256_set1_epi8('@');
        while (p + 32 <= e) {
            const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p));
            const __m256i a = _mm256_xor_si256(_mm256_sub_epi8(x, lo), flip);          // (x - 33) as a biased signed value
            const __m256i nongraph = _mm256_cmpgt_epi8(a, lim);                          // (x - 33) > 93 unsigned
            const __m256i term = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(x, gt), _mm256_cmpeq_epi8(x, pl)), _mm256_cmpeq_epi8(x, at));
            const unsigned m = (unsigned)_mm256_movemask_epi8(_mm256_or_si256(nongraph, term));
            if (m) return p + __builtin_ctz(m);
            p += 32;
        }
        return p;
    }
#endif
    const unsigned char *skipGraph(const unsigned char *p, const unsigned char *e) const
    {
#if defined(__x86_64__)
        static const bool avx2 = __builtin_cpu_supports("avx2");
        if (avx2) p = skipGraphAvx2(p, e);
#endif
        while (p < e && cls_[*p] == 0) p++;
        return p;
    }

    bool fill()
    {
        if (eof_) return false;
        int n = gzread(f_, buf_.data(), (unsigned)buf_.size());
        if (n > 0) remaining_ -= std::min(remaining_, (size_t)n);
        begin_ = 0;
        end_ = n > 0 ? (size_t)n : 0;
        if (n < (int)buf_.size()) eof_ = true;
        return end_ > 0;
    }
    int getc()
    {
        if (begin_ >= end_ && !fill()) return -1;
        return buf_[begin_++];
    }
    // ks_getuntil(ks, KS_SEP_SPACE, ...): false when already at end of file
    bool getUntilSpace(std::string &out, int &delim)
    {
        out.clear();
        delim = 0;
        if (begin_ >= end_ && eof_) return false;
        for (;;) {
            if (begin_ >= end_ && !fill()) break;
            size_t i = begin_;
            while (i < end_ && !isspace(buf_[i])) i++;
            out.append(reinterpret_cast<const char *>(buf_.data() + begin_), i - begin_);
            begin_ = i + 1;
            if (i < end_) { delim = buf_[i]; break; }
            begin_ = end_;
        }
        return true;
    }
    void getUntilNewline(std::string &out)
    {
        out.clear();
        for (;;) {
            if (begin_ >= end_ && !fill()) break;
            const unsigned char *p = static_cast<const unsigned char *>(memchr(buf_.data() + begin_, '\n', end_ - begin_));
            size_t i = p ? (size_t)(p - buf_.data()) : end_;
            out.append(reinterpret_cast<const char *>(buf_.data() + begin_), i - begin_);
            if (p) { begin_ = i + 1; break; }
            begin_ = end_;
        }
    }
};

}  // namespace mashhost
