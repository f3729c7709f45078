// fastaq.h -- streaming FASTA / FASTQ reader (plain or gzip) for the sage2gpu host program.
// Same record grammar the reference reads through kseq (inputReader/fastAQReader.cpp:16-45): a header
// line starting with '>' or '@' (name = up to the first white space), sequence lines concatenated up
// to the next line that starts with '>', '@' or '+', and for FASTQ a '+' line followed by as many
// quality characters as there are bases (possibly over several lines).  A record whose quality
// string is shorter or longer than its sequence ends the stream, as the reference's reader does.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

namespace sg_host {

class FastAQStream {
public:
    explicit FastAQStream(const std::string &path);     // throws std::runtime_error if the file cannot be opened
    // continue on an already open zlib handle (ownership is taken) after `n` bytes that were read from it before
    FastAQStream(void *gz_handle, const uint8_t *pending, size_t n);
    ~FastAQStream();
    FastAQStream(const FastAQStream &) = delete;
    FastAQStream &operator=(const FastAQStream &) = delete;
    // next record's sequence appended to `out` (no terminator); returns false at the end of the stream
    bool next(std::vector<uint8_t> &out, uint64_t &seq_len);

private:
    int get();                       // next byte or -1
    void skip_line();
    void *gz_ = nullptr;
    std::vector<uint8_t> buf_;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false;
    int marker_ = 0;                 // header character already consumed ('>' / '@'), 0 = none yet
};

// One or two mate files read alternately: record 2r from the first file, 2r+1 from the second one
// (or from the first one again when the pairs are interleaved), inputReader/inputReader.cpp:26-49.
class MatePairStream {
public:
    MatePairStream(const std::string &file1, const std::string &file2);
    ~MatePairStream();
    bool next(std::vector<uint8_t> &out, uint64_t &seq_len);

private:
    FastAQStream *a_ = nullptr, *b_ = nullptr;
    uint64_t n_ = 0;
};

}  // namespace sg_host
