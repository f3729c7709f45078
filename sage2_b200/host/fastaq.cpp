// fastaq.cpp -- see fastaq.h.
#include "fastaq.h"

#include <string.h>
#include <zlib.h>
#include <stdexcept>

namespace sg_host {

FastAQStream::FastAQStream(const std::string &path) : buf_(1 << 20)
{
    gz_ = gzopen(path.c_str(), "r");
    if (!gz_) throw std::runtime_error("cannot open " + path);
    gzbuffer((gzFile)gz_, 1 << 20);
}

FastAQStream::FastAQStream(void *gz_handle, const uint8_t *pending, size_t n) : gz_(gz_handle), buf_(n > ((size_t)1 << 20) ? n : (size_t)1 << 20)
{
    if (n) memcpy(buf_.data(), pending, n);
    pos_ = 0; end_ = n;
}

FastAQStream::~FastAQStream()
{
    if (gz_) gzclose((gzFile)gz_);
}

int FastAQStream::get()
{
    if (pos_ == end_) {
        if (eof_) return -1;
        const int n = gzread((gzFile)gz_, buf_.data(), (unsigned)buf_.size());
        if (n <= 0) { eof_ = true; return -1; }
        pos_ = 0; end_ = (size_t)n;
    }
    return buf_[pos_++];
}

void FastAQStream::skip_line()
{
    int c;
    while ((c = get()) != -1 && c != '\n') {}
}

bool FastAQStream::next(std::vector<uint8_t> &out, uint64_t &seq_len)
{
    int c;
    if (marker_ == 0) {                 // find the first header
        while ((c = get()) != -1 && c != '>' && c != '@') {}
        if (c == -1) return false;
    }
    marker_ = 0;
    skip_line();                        // name and comment are not needed for steps 1-3
    const size_t start = out.size();
    // sequence lines
    for (;;) {
        c = get();
        if (c == -1 || c == '>' || c == '@' || c == '+') break;
        if (c == '\n') continue;
        // rest of this line
        do {
            if (c != '\r') out.push_back((uint8_t)c);
            c = get();
        } while (c != -1 && c != '\n');
        if (c == -1) break;
    }
    seq_len = out.size() - start;
    if (c == '>' || c == '@') { marker_ = c; return true; }
    if (c != '+') return true;           // FASTA record at end of file
    skip_line();                        // the '+' line
    uint64_t q = 0;
    while (q < seq_len) {
        c = get();
        if (c == -1) break;
        if (c == '\n' || c == '\r') continue;
        ++q;
    }
    if (q != seq_len) { out.resize(start); return false; }     // truncated quality: the stream ends here
    // the quality string must end at a line end; a longer one is an error the reference stops at
    c = get();
    while (c == '\r') c = get();
    if (c != -1 && c != '\n') { out.resize(start); return false; }
    return true;
}

MatePairStream::MatePairStream(const std::string &file1, const std::string &file2)
{
    a_ = new FastAQStream(file1);
    if (!file2.empty()) {
        try { b_ = new FastAQStream(file2); } catch (...) { delete a_; throw; }
    }
}

MatePairStream::~MatePairStream()
{
    delete a_;
    delete b_;
}

bool MatePairStream::next(std::vector<uint8_t> &out, uint64_t &seq_len)
{
    FastAQStream *s = (n_ % 2 == 1 && b_) ? b_ : a_;
    if (!s->next(out, seq_len)) return false;
    ++n_;
    return true;
}

}  // namespace sg_host
