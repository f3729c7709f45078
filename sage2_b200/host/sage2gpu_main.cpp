// sage2gpu -- host program: SAGE2's command line and step structure (main.cpp:11-132, 384-521) with
// steps 1-3 (organise reads, prefix/suffix hash table, economy overlap graph) running on a B200 through
// the C ABI of libsage2gpu (include/sage2gpu.h).  It writes the reference's own intermediate files
// (<prefix>.reads, <prefix>.graph3, byte for byte what `SAGE2 -s -M 3` writes), so the reference's
// steps 4-7 continue from them unchanged: `SAGE2 -m 4 -i <prefix> ...`, which this program runs
// itself when it is told where the reference binary is (--reference PATH or $SAGE2_REFERENCE_BIN).
//
// Flags are the reference's: -f/--fileInput, -l/--listInput, -k/--minOverlap, -o/--outputDir,
// -p/--prefix, -i/--inputPrefix, -m/--minStep, -M/--maxStep, -s/--saveAll, -d/--debug, -h/--help.
// Added: --device N (CUDA ordinal), --reference PATH, --devices LIST (several GPUs in ONE process: one context per listed
// device, every stage partitioned -- reads organised by key range, table built by key-hash shard, search by id slice --
// and the all-gathers between the stages done as peer copies; no NCCL, no Python; same files out).
//
// Differences, all stated in the log: steps 1-3 are one device pass, so a restart at -m 2 / -m 3
// recomputes from the input files instead of loading <prefix>.reads / <prefix>.hashTable (results are
// identical); <prefix>.hashTable is not written (the reference's slot order is an artefact of its own
// hash function and is read by nothing but its -m 3 restart).  There is no CPU fallback: without a
// CUDA device the program stops with an error, like the reference's printError (utils.cpp:36-40).
#include <fcntl.h>
#include <getopt.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <functional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sage2gpu.h"
#include <zlib.h>

#include "fastaq.h"

namespace {

struct Options {
    int minStep = 1, maxStep = 7, minOverlap = 0, device = 0;
    bool saveAll = false, debugging = false, parseOnly = false;
    std::vector<int> devices;       // --devices: one context per listed GPU, every stage partitioned, peer copies between them
    std::string fileInput, listInput, outputDir, prefixName = "untitled", inputPrefix, referenceBin;
};

std::ofstream logStream;

[[noreturn]] void printError(const std::string &kind, const std::string &msg)     // utils.cpp:36-40
{
    logStream << kind << " : " << msg << "!\n";
    logStream.flush();
    std::cerr << kind << " : " << msg << "!\n";
    exit(EXIT_FAILURE);
}

std::string trim(const std::string &s)
{
    const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}

void printUsage()
{
    std::cout << "USAGE:\n\tsage2gpu [options] -f <inputFile> -k <minOverlap>\n\tsage2gpu [options] -l <inputList> -k <minOverlap>\n\n";
}

void printListOfArgs()
{
    std::cout << "OPTIONS (those of SAGE2):\n"
                 "\t-f|--fileInput <file>   interleaved FASTA/FASTQ file (plain or gzip)\n"
                 "\t-l|--listInput <file>   list of input files (f1=/f2= pairs or f= interleaved)\n"
                 "\t-k|--minOverlap <int>   minimum overlap length\n"
                 "\t-o|--outputDir <dir>    output directory\n"
                 "\t-p|--prefix <name>      prefix of the output files [untitled]\n"
                 "\t-i|--inputPrefix <name> prefix of the files a restart reads [= prefix]\n"
                 "\t-m|--minStep <1..7>     first step to run [1]\n"
                 "\t-M|--maxStep <1..7>     last step to run [7]\n"
                 "\t-s|--saveAll            save the intermediate files of every step\n"
                 "\t-d|--debug\n"
                 "\t-h|--help\n"
                 "ADDED:\n"
                 "\t--device <int>          CUDA device ordinal [0]\n"
                 "\t--reference <path>      SAGE2 binary that runs steps 4-7 from the files written here\n\n";
}

// "0,1,2" or "0-3" (or a mix); a device may be listed twice (two contexts on one GPU)
std::vector<int> parseDevices(const char *arg)
{
    std::vector<int> v;
    std::stringstream ss(arg);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        const size_t dash = tok.find('-', 1);
        if (dash != std::string::npos) { for (int d = atoi(tok.substr(0, dash).c_str()); d <= atoi(tok.substr(dash + 1).c_str()); ++d) v.push_back(d); }
        else if (!tok.empty()) v.push_back(atoi(tok.c_str()));
    }
    return v;
}

void parseArgs(int argc, char **argv, Options &o)
{
    static struct option long_options[] = {
        { "help", no_argument, 0, 'h' },          { "fileInput", required_argument, 0, 'f' },
        { "minOverlap", required_argument, 0, 'k' }, { "listInput", required_argument, 0, 'l' },
        { "outputDir", required_argument, 0, 'o' }, { "prefix", required_argument, 0, 'p' },
        { "inputPrefix", required_argument, 0, 'i' }, { "minStep", required_argument, 0, 'm' },
        { "maxStep", required_argument, 0, 'M' },  { "saveAll", no_argument, 0, 's' },
        { "debug", no_argument, 0, 'd' },          { "device", required_argument, 0, 1001 },
        { "reference", required_argument, 0, 1002 }, { "parse-only", no_argument, 0, 1003 },
        { "devices", required_argument, 0, 1004 }, { 0, 0, 0, 0 } };
    bool fFlag = false, lFlag = false;
    int c, idx = 0;
    while ((c = getopt_long(argc, argv, "hf:l:k:o:p:i:m:M:sd", long_options, &idx)) != -1) {
        switch (c) {
            case 'h': std::cout << "\n"; printUsage(); printListOfArgs(); exit(0);
            case 'f': o.fileInput = optarg; fFlag = true; break;
            case 'l': o.listInput = optarg; lFlag = true; break;
            case 'k': o.minOverlap = atoi(optarg); break;
            case 'o': {
                o.outputDir = optarg;
                while (!o.outputDir.empty() && o.outputDir.back() == '/') o.outputDir.pop_back();
                if (!o.outputDir.empty() || (optarg[0] == '/')) o.outputDir += "/";
                break;
            }
            case 'p': o.prefixName = optarg; break;
            case 'i': o.inputPrefix = optarg; break;
            case 'm': o.minStep = atoi(optarg); if (o.minStep < 1) o.minStep = 1; break;
            case 'M': o.maxStep = atoi(optarg); if (o.maxStep > 7) o.maxStep = 7; break;
            case 's': o.saveAll = true; break;
            case 'd': o.debugging = true; break;
            case 1001: o.device = atoi(optarg); break;
            case 1002: o.referenceBin = optarg; break;
            case 1003: o.parseOnly = true; break;
            case 1004: o.devices = parseDevices(optarg); break;
            case '?': std::cout << "\n"; exit(0);
            default: std::cout << "[ERROR] Wrong command line arguments!\n\n"; exit(0);
        }
    }
    if (optind < argc) {
        std::cout << "[WARNING] There are some non-option arguments: ";
        while (optind < argc) std::cout << argv[optind++] << " ";
        std::cout << "\n";
    }
    if (fFlag && lFlag) {
        std::cout << "[ERROR] Options -f|--fileInput and -l|--listInput are mutually exclusive!\n\n";
        exit(0);
    }
    if (o.referenceBin.empty() && getenv("SAGE2_REFERENCE_BIN")) o.referenceBin = getenv("SAGE2_REFERENCE_BIN");
}

bool checkRequired(Options &o)      // main.cpp:498-521
{
    bool ok = true;
    if (o.fileInput.empty() && o.listInput.empty()) {
        std::cout << "[ERROR] One of the options -f|--fileInput or -l|--listInput is required.\n";
        ok = false;
    }
    if (o.minOverlap == 0) { std::cout << "[ERROR] Option -k|--minOverlap is required.\n"; ok = false; }
    if (o.maxStep < o.minStep) { std::cout << "[ERROR] maxStep should not be smaller than minStep!\n\n"; ok = false; }
    if (o.inputPrefix.empty()) o.inputPrefix = o.prefixName;
    return ok;
}

void banner(const char *step, const char *what)
{
    const char *bar = "***********************************************************************************************************\n";
    logStream << bar << "                                               " << step << "\n" << what << "\n" << bar;
    logStream.flush();
}

// ---- step 1 input: parse on the host, upload chunk by chunk (pinned, asynchronous) -------------------
struct Chunk {
    uint8_t *bases = nullptr;
    int64_t *offsets = nullptr;
    size_t cap_bases = 0, cap_reads = 0, n_reads = 0, n_bases = 0;
    bool pinned = false;
};

class Uploader {
public:
    Uploader(sage2gpu_ctx *ctx, int k) : ctx_(ctx)
    {
        if (!ctx_) return;          // --parse-only
        for (Chunk &c : chunk_) {
            c.cap_bases = (size_t)64 << 20;
            c.cap_reads = (size_t)1 << 19;
            alloc(c);
        }
        if (sage2gpu_load_begin(ctx_, k) != 0) printError("CUDA", sage2gpu_last_error(ctx_));
    }
    ~Uploader()
    {
        for (Chunk &c : chunk_) release(c);
        sage2gpu_host_free(text_);
    }
    void add(const uint8_t *seq, size_t len)
    {
        Chunk *c = &chunk_[cur_];
        if (c->n_reads == c->cap_reads || c->n_bases + len > c->cap_bases) {
            flush();
            c = &chunk_[cur_];
            if (len > c->cap_bases) { release(*c); c->cap_bases = len; alloc(*c); }
        }
        memcpy(c->bases + c->n_bases, seq, len);
        c->offsets[c->n_reads] = (int64_t)c->n_bases;
        c->n_bases += len;
        c->n_reads++;
        c->offsets[c->n_reads] = (int64_t)c->n_bases;
    }
    void finish()
    {
        flush();
        if (sage2gpu_load_finish(ctx_) != 0) printError("CUDA", sage2gpu_last_error(ctx_));
    }
    // several GPUs: filter + pack only; organizeReads follows partitioned over the contexts
    void finishPacked(uint64_t &n_reads, int &max_len, uint64_t &good, uint64_t &bp)
    {
        flush();
        if (sage2gpu_load_finish_packed(ctx_, &n_reads, &max_len, &good, &bp) != 0) printError("CUDA", sage2gpu_last_error(ctx_));
    }
    // everything added so far is on the device after this
    void sync()
    {
        if (!ctx_) return;
        flush();
        if (sage2gpu_load_append(ctx_, nullptr, nullptr, 0) != 0) printError("CUDA", sage2gpu_last_error(ctx_));
    }
    sage2gpu_ctx *ctx() const { return ctx_; }
    uint64_t count()
    {
        uint64_t n = 0;
        if (sage2gpu_load_count(ctx_, &n) != 0) printError("CUDA", sage2gpu_last_error(ctx_));
        return n;
    }
    // pinned buffer for raw file text (device-side record splitting)
    uint8_t *text(size_t &cap)
    {
        if (!text_) { text_ = (uint8_t *)sage2gpu_host_alloc(kTextCap); if (!text_) printError("MEM_ALLOC", "pinned text buffer"); }
        cap = kTextCap;
        return text_;
    }

private:
    static constexpr size_t kTextCap = (size_t)32 << 20;
    uint8_t *text_ = nullptr;
    void alloc(Chunk &c)
    {
        c.bases = (uint8_t *)sage2gpu_host_alloc(c.cap_bases);
        c.offsets = (int64_t *)sage2gpu_host_alloc((c.cap_reads + 1) * sizeof(int64_t));
        c.pinned = c.bases && c.offsets;
        if (!c.pinned) printError("MEM_ALLOC", "pinned upload buffers");
        c.n_reads = c.n_bases = 0;
        c.offsets[0] = 0;
    }
    void release(Chunk &c)
    {
        sage2gpu_host_free(c.bases);
        sage2gpu_host_free(c.offsets);
        c.bases = nullptr; c.offsets = nullptr;
    }
    void flush()
    {
        Chunk &c = chunk_[cur_];
        if (c.n_reads) {
            // returns once the PREVIOUS chunk's copy is complete; this chunk's copy runs while the parser fills the other buffer
            if (sage2gpu_load_append(ctx_, c.bases, c.offsets, (int64_t)c.n_reads) != 0) printError("CUDA", sage2gpu_last_error(ctx_));
        }
        cur_ ^= 1;
        Chunk &n = chunk_[cur_];
        n.n_reads = n.n_bases = 0;
        n.offsets[0] = 0;
    }
    sage2gpu_ctx *ctx_;
    Chunk chunk_[2];
    int cur_ = 0;
};

// --parse-only: the input side alone (no device): record count, base count and an FNV-1a checksum of the
// sequences in the order they would be uploaded.  Used by the CPU tests of the parser / list grammar.
struct ParseSink {
    uint64_t reads = 0, bases = 0, fnv = 1469598103934665603ull;
    void add(const uint8_t *s, size_t n)
    {
        for (size_t i = 0; i < n; ++i) { fnv ^= s[i]; fnv *= 1099511628211ull; }
        fnv ^= 0xFF; fnv *= 1099511628211ull;
        reads++; bases += n;
    }
};
ParseSink *g_parse_sink = nullptr;

// sequential parser (any layout the reference accepts) on `in`, at most max_records records
uint64_t parseSequential(Uploader &up, sg_host::FastAQStream &in, uint64_t max_records)
{
    std::vector<uint8_t> seq;
    uint64_t len = 0, n = 0;
    while (n < max_records) {
        seq.clear();
        if (!in.next(seq, len)) break;
        if (g_parse_sink) g_parse_sink->add(seq.data(), seq.size());
        else up.add(seq.data(), seq.size());
        ++n;
    }
    return n;
}

// One file: raw text goes to the device in 32 MB pieces and the records are split there
// (sage2gpu_load_append_text); text that is not in the regular layout is parsed sequentially from that point on.
uint64_t readFile(Uploader &up, const std::string &path, uint64_t max_records, bool &used_device_parser)
{
    used_device_parser = false;
    if (!up.ctx()) {                       // --parse-only
        sg_host::FastAQStream in(path);
        return parseSequential(up, in, max_records);
    }
    // plain files are read with read(2) straight into the pinned buffer; gzip goes through zlib
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("cannot open " + path);
    unsigned char magic[2] = { 0, 0 };
    const bool is_gz = pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    gzFile gz = nullptr;
    if (is_gz) {
        gz = gzdopen(fd, "r");
        if (!gz) { close(fd); throw std::runtime_error("cannot open " + path); }
        gzbuffer(gz, 1 << 20);
    }
    size_t cap = 0, have = 0;
    uint8_t *buf = up.text(cap);
    int marker = 0;
    bool eof = false;
    uint64_t total = 0;
    up.sync();                             // keep the upload order: everything added before is on the device
    for (;;) {
        while (have < cap && !eof) {
            const long n = is_gz ? (long)gzread(gz, buf + have, (unsigned)(cap - have)) : (long)read(fd, buf + have, cap - have);
            if (n <= 0) eof = true; else have += (size_t)n;
        }
        if (have == 0) break;
        uint64_t consumed = 0, nrec = 0;
        const int rc = sage2gpu_load_append_text(up.ctx(), buf, have, eof ? 1 : 0, &marker, max_records - total, &consumed, &nrec);
        if (rc == SAGE2GPU_ERR_FORMAT || (rc == 0 && consumed == 0 && !eof)) {
            if (!gz) gz = gzdopen(fd, "r");                   // transparent mode, continues at the descriptor's offset
            if (!gz) { close(fd); throw std::runtime_error("cannot read " + path); }
            sg_host::FastAQStream in(gz, buf, have);          // takes the handle; continues behind what the device consumed
            total += parseSequential(up, in, max_records - total);
            up.sync();
            return total;
        }
        if (rc != 0) printError("CUDA", sage2gpu_last_error(up.ctx()));
        used_device_parser = true;
        total += nrec;
        have -= (size_t)consumed;
        if (have) memmove(buf, buf + consumed, have);
        if (total >= max_records || (eof && have == 0)) break;
    }
    if (gz) gzclose(gz); else close(fd);
    return total;
}

uint64_t readDataset(Uploader &up, const std::string &f1, const std::string &f2)      // readLoader.cpp:133-174
{
    logStream << "In function readDatasetInBytes().\n";
    logStream << "Reading from file: " << f1 << " " << (f2.empty() ? "" : "& " + f2) << "\n";
    logStream.flush();
    uint64_t n = 0;
    try {
        bool dev1 = false, dev2 = false;
        if (f2.empty()) {
            n = readFile(up, f1, UINT64_MAX, dev1);
        } else if (!up.ctx()) {
            sg_host::MatePairStream in(f1, f2);     // --parse-only keeps the reference's alternating order
            std::vector<uint8_t> seq;
            uint64_t len = 0;
            for (;;) { seq.clear(); if (!in.next(seq, len)) break; g_parse_sink->add(seq.data(), seq.size()); ++n; }
        } else {
            // The reference alternates mate 1 / mate 2 and stops when a file ends (inputReader.cpp:26-49): with n1 and n2
            // records it keeps min(n1, n2 + 1) of file 1 and min(n2, n1) of file 2.  Read ids are ranks in the sorted set,
            // so the upload order is free: file 1, then file 2 capped at n1, then the surplus of file 1 is dropped.
            const uint64_t start1 = up.count();
            const uint64_t n1 = readFile(up, f1, UINT64_MAX, dev1);
            const uint64_t n2 = readFile(up, f2, n1, dev2);
            uint64_t keep1 = n1;
            if (n1 > n2 + 1) {
                keep1 = n2 + 1;
                if (sage2gpu_load_remove(up.ctx(), start1 + keep1, n1 - keep1) != 0) printError("CUDA", sage2gpu_last_error(up.ctx()));
            }
            n = keep1 + n2;
        }
        logStream << "\tRecords split on the device: " << ((dev1 || dev2) ? "yes" : "no (sequential parser)") << "\n";
    } catch (const std::runtime_error &e) {
        printError("OPEN_FILE", e.what());
    }
    logStream << "\t" << std::setw(21) << "Total reads in file: " << n << "\n";
    logStream.flush();
    return n;
}

uint64_t loadFromList(Uploader &up, const std::string &listPath)      // readLoader.cpp:73-131
{
    std::ifstream fin(listPath.c_str());
    if (!fin.is_open()) printError("OPEN_FILE", listPath);
    std::string line, val1;
    uint32_t mateFile = 0;
    uint64_t total = 0;
    while (std::getline(fin, line)) {
        if (line.empty() || line[0] == '#') continue;
        const size_t pos = line.find('=');
        if (pos == std::string::npos) { std::cout << "[ERROR] List of input files in a wrong format!\n\n"; exit(0); }
        const std::string var = trim(line.substr(0, pos)), val = trim(line.substr(pos + 1));
        if (mateFile % 2 == 0 && var == "f1") val1 = val;
        else if (mateFile % 2 == 0 && var == "f") { total += readDataset(up, val, ""); mateFile++; }
        else if (mateFile % 2 == 1 && var == "f2") total += readDataset(up, val1, val);
        else { std::cout << "[ERROR] List of input files in a wrong format!\n\n"; exit(0); }
        mateFile++;
    }
    logStream << std::setw(20) << "Number of datasets: " << mateFile / 2 << "\n";
    return total;
}

double seconds_since(const std::chrono::steady_clock::time_point &t0)
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int runReference(const Options &o, int fromStep)
{
    std::vector<std::string> a = { o.referenceBin };
    if (!o.listInput.empty()) { a.push_back("-l"); a.push_back(o.listInput); } else { a.push_back("-f"); a.push_back(o.fileInput); }
    a.push_back("-k"); a.push_back(std::to_string(o.minOverlap));
    if (!o.outputDir.empty()) { a.push_back("-o"); a.push_back(o.outputDir); }
    a.push_back("-p"); a.push_back(o.prefixName);
    a.push_back("-i"); a.push_back(fromStep == 4 && o.minStep <= 3 ? o.prefixName : o.inputPrefix);
    a.push_back("-m"); a.push_back(std::to_string(fromStep));
    a.push_back("-M"); a.push_back(std::to_string(o.maxStep));
    if (o.saveAll) a.push_back("-s");
    std::vector<char *> argv;
    for (auto &s : a) argv.push_back(const_cast<char *>(s.c_str()));
    argv.push_back(nullptr);
    const pid_t pid = fork();
    if (pid < 0) return -1;
    if (pid == 0) { execv(argv[0], argv.data()); _exit(127); }
    int status = 0;
    waitpid(pid, &status, 0);
    return WIFEXITED(status) ? WEXITSTATUS(status) : -1;
}

}  // namespace

// ---- several GPUs in one process ---------------------------------------------------------------------------------
// One context per device, one host thread per context and stage (the library calls block), peer copies between the
// stages.  The steps are those of sage2_b200/multi.py::partitioned_slice_steps with context 0 holding the whole input as
// its "slice" (it received the streamed upload) and the others an empty one.
struct MultiGpu {
    std::vector<sage2gpu_ctx *> ctx;
    int G() const { return (int)ctx.size(); }
    void each(const std::function<void(int)> &fn)
    {
        std::vector<std::thread> th;
        for (int r = 0; r < G(); ++r) th.emplace_back(fn, r);
        for (auto &t : th) t.join();
    }
    void check(int r, int rc) { if (rc != 0) printError("CUDA", std::string("device context ") + std::to_string(r) + ": " + sage2gpu_last_error(ctx[r])); }
    // all-gather of blocks: block q (bytes[q] bytes at byte offset off[q]) of every context's array comes from context q
    void gather(const std::vector<void *> &ptr, const std::vector<uint64_t> &off, const std::vector<uint64_t> &bytes)
    {
        each([&](int r) {
            for (int q = 0; q < G(); ++q)
                if (q != r && bytes[q]) check(r, sage2gpu_peer_copy(ctx[r], (char *)ptr[r] + off[q], ctx[q], (const char *)ptr[q] + off[q], bytes[q]));
        });
    }
};

// steps 1 (after the upload into context 0) to 3; the result is complete in every context
void runPartitioned(MultiGpu &m, Uploader &up, int k)
{
    const int G = m.G();
    uint64_t N = 0, good = 0, bp = 0;
    int max_len = 0;
    up.finishPacked(N, max_len, good, bp);
    std::vector<uint64_t> counts(G, 0), zero(G, 0);
    counts[0] = N;
    m.each([&](int r) {
        if (r > 0) m.check(r, sage2gpu_pack_slice(m.ctx[r], nullptr, nullptr, 0, k, 1, max_len > 0 ? max_len : 1, nullptr, nullptr, nullptr));
    });
    sage2gpu_counters c0;
    sage2gpu_get_counters(m.ctx[0], &c0);
    const uint64_t SW = c0.record_words;
    std::vector<void *> p(G, nullptr), p2(G, nullptr), p3(G, nullptr);
    m.each([&](int r) { m.check(r, sage2gpu_raw_gather_layout(m.ctx[r], r, G, counts.data(), &p[r], nullptr, nullptr)); });
    {
        std::vector<uint64_t> off(G, 0), bytes(G, 0);
        bytes[0] = N * SW * 8;
        m.gather(p, off, bytes);
    }
    std::vector<uint64_t> uloc(G, 0);
    m.each([&](int r) {
        m.check(r, sage2gpu_raw_gather_finish(m.ctx[r], N, good, bp));
        m.check(r, sage2gpu_organize_partition(m.ctx[r], r, G, &uloc[r]));
    });
    std::vector<uint64_t> stride(G, 0);
    m.each([&](int r) { m.check(r, sage2gpu_reads_gather_layout(m.ctx[r], uloc.data(), &p[r], &p2[r], &p3[r], nullptr, nullptr, &stride[r])); });
    {
        std::vector<uint64_t> off(G, 0), bytes(G, 0), off2(G, 0), bytes2(G, 0);
        uint64_t acc = 0;
        for (int q = 0; q < G; ++q) { off[q] = acc * stride[0] * 8; bytes[q] = uloc[q] * stride[0] * 8; off2[q] = acc * 2; bytes2[q] = uloc[q] * 2; acc += uloc[q]; }
        m.gather(p, off, bytes);
        m.gather(p2, off2, bytes2);
        m.gather(p3, off2, bytes2);
    }
    std::vector<uint64_t> ec(G, 0), dk(G, 0), ov(G, 0), sps(G, 0);
    m.each([&](int r) {
        m.check(r, sage2gpu_reads_gather_finish(m.ctx[r]));
        m.check(r, sage2gpu_build_hash_table_part(m.ctx[r], r, G));
        m.check(r, sage2gpu_table_shard_info(m.ctx[r], nullptr, &ec[r], &dk[r], &ov[r]));
    });
    m.each([&](int r) { m.check(r, sage2gpu_table_gather_layout(m.ctx[r], ec.data(), &p[r], &p2[r], &sps[r], nullptr)); });
    {
        std::vector<uint64_t> off(G, 0), bytes(G, 0), off2(G, 0), bytes2(G, 0);
        uint64_t acc = 0;
        for (int q = 0; q < G; ++q) { off[q] = (uint64_t)q * sps[0] * 8; bytes[q] = sps[0] * 8; off2[q] = acc * 4; bytes2[q] = ec[q] * 4; acc += ec[q]; }
        m.gather(p, off, bytes);
        m.gather(p2, off2, bytes2);
    }
    m.each([&](int r) {
        m.check(r, sage2gpu_table_gather_finish(m.ctx[r], ec.data(), dk.data(), ov.data()));
        m.check(r, sage2gpu_phase_a_partition(m.ctx[r], r, G));
    });
    m.each([&](int r) {
        for (int q = 0; q < G; ++q) if (q != r) m.check(r, sage2gpu_phase_a_import(m.ctx[r], m.ctx[q], q));
    });
    m.each([&](int r) { m.check(r, sage2gpu_finish_graph(m.ctx[r])); });
}

// --devices with more than one entry: steps 1-3 partitioned over the listed GPUs, the same files and log lines out
int runOnSeveralDevices(const Options &o, const std::string &base)
{
    MultiGpu m;
    for (int d : o.devices) {
        sage2gpu_ctx *c = nullptr;
        if (sage2gpu_create(&c, d) != 0) printError("CUDA", "no usable CUDA device " + std::to_string(d) + " (libsage2gpu has no CPU fallback)");
        m.ctx.push_back(c);
    }
    logStream << "Devices: " << m.G() << " contexts (one per listed GPU), every stage partitioned, peer copies between the stages.\n";
    banner("STEP 1", "                                          organizing reads");
    auto t0 = std::chrono::steady_clock::now();
    {
        Uploader up(m.ctx[0], o.minOverlap);
        if (!o.listInput.empty()) loadFromList(up, o.listInput);
        else readDataset(up, o.fileInput, "");
        logStream << "Parsed and uploaded in " << seconds_since(t0) << " sec.\n";
        banner("STEPS 2-3", "                             building hash table and overlap graph");
        auto t1 = std::chrono::steady_clock::now();
        runPartitioned(m, up, o.minOverlap);
        logStream << "Functions organizeReads() .. sortEconomyGraph() on " << m.G() << " devices in " << seconds_since(t1) << " sec.\n";
    }
    sage2gpu_counters cnt;
    sage2gpu_get_counters(m.ctx[0], &cnt);
    logStream << std::setw(20) << "Total reads: " << cnt.total_reads << "\n";
    logStream << std::setw(20) << "Good reads: " << cnt.good_reads << "\n";
    logStream << std::setw(20) << "Bad reads: " << cnt.total_reads - cnt.good_reads << "\n";
    logStream << "\t" << std::setw(21) << "Average read length: " << cnt.avg_len << "\n";
    logStream << "\tNumber of unique reads: " << cnt.unique_reads << "\n";
    logStream << "\tSize of hash table: " << cnt.table_capacity << " slots, " << cnt.distinct_keys << " distinct keys\n";
    logStream << "\tNumber of hashes over threshold: " << cnt.keys_over_threshold << "\n";
    logStream << "\tTotal reads contained by extension: " << cnt.contained_ext << "\n";
    logStream << "\tTotal reads contained by size: " << cnt.contained_size << "\n";
    logStream << "\tTotal reads left to explore: " << cnt.left_to_explore << "\n";
    logStream << "\tTotal edges inserted: " << cnt.edges_inserted_c << "\n";
    logStream << "\tTotal transitive edges removed: " << cnt.transitive_removed << "\n";
    logStream << "\tEdges in the overlap graph: " << cnt.n_edges << "\n";
    if (sage2gpu_write_reads(m.ctx[0], (base + ".reads").c_str()) != 0) printError("OPEN_FILE", sage2gpu_last_error(m.ctx[0]));
    // the graph from the LAST context: every context must hold the complete result
    if (sage2gpu_write_graph3(m.ctx[m.G() - 1], (base + ".graph3").c_str()) != 0) printError("OPEN_FILE", sage2gpu_last_error(m.ctx[m.G() - 1]));
    logStream.flush();
    for (sage2gpu_ctx *c : m.ctx) sage2gpu_destroy(c);
    if (o.maxStep <= 3) return 0;
    if (o.referenceBin.empty()) {
        std::cout << "sage2gpu: steps 1-3 done (" << base << ".reads, " << base << ".graph3); continue with SAGE2 -m 4 -i " << o.prefixName << "\n";
        return 0;
    }
    logStream << "Running " << o.referenceBin << " -m 4 for steps 4-" << o.maxStep << ".\n";
    logStream.close();
    return runReference(o, 4);
}

int main(int argc, char **argv)
{
    Options o;
    parseArgs(argc, argv, o);
    if (!checkRequired(o)) {
        std::cout << "\n";
        printUsage();
        std::cout << "(For more information run ./sage2gpu -h)\n\n";
        exit(0);
    }
    if (!o.outputDir.empty()) {
        const std::string cmd = "mkdir -p '" + o.outputDir + "'";
        if (system(cmd.c_str()) != 0) std::cerr << "[WARNING] cannot create " << o.outputDir << "\n";
    }
    if (o.parseOnly) {
        ParseSink sink;
        g_parse_sink = &sink;
        logStream.open("/dev/null");
        Uploader none(nullptr, o.minOverlap);
        if (!o.listInput.empty()) loadFromList(none, o.listInput);
        else readDataset(none, o.fileInput, "");
        printf("{\"reads\": %llu, \"bases\": %llu, \"fnv1a\": \"%016llx\"}\n", (unsigned long long)sink.reads,
               (unsigned long long)sink.bases, (unsigned long long)sink.fnv);
        return 0;
    }
    if (o.minStep >= 4) {       // nothing of ours to do: the reference's own steps
        if (o.referenceBin.empty()) { std::cerr << "[ERROR] steps 4-7 are the reference's: give --reference <SAGE2 binary>\n"; return 1; }
        return runReference(o, o.minStep);
    }

    const std::string base = o.outputDir + o.prefixName;
    logStream.open((base + (o.maxStep > 3 ? ".gpu.log" : ".log")).c_str());     // SAGE2 -m 4 rewrites <prefix>.log
    logStream << std::fixed << std::setprecision(2);
    logStream << "***********************************************************************************************************\n";
    logStream << "\tEXECUTING PROGRAM: sage2gpu (SAGE2 steps 1-3 on libsage2gpu)\n";
    if (!o.listInput.empty()) logStream << "\t  INPUT LIST PATH: " << o.listInput << "\n";
    else logStream << "\t  INPUT FILE PATH: " << o.fileInput << "\n";
    logStream << "\t OUTPUT DIRECTORY: " << o.outputDir << "\n";
    logStream << "\t    OUTPUT PREFIX: " << o.prefixName << "\n";
    logStream << "\t  MINIMUM OVERLAP: " << o.minOverlap << "\n";
    logStream << "\t       START STEP: " << o.minStep << "\n";
    logStream << "\t         END STEP: " << o.maxStep << "\n";
    logStream << "\t   SAVE ALL FILES: " << (o.saveAll ? "TRUE" : "FALSE") << "\n";
    logStream << "***********************************************************************************************************\n\n";
    if (o.minStep > 1)
        logStream << "NOTE: steps 1-3 are one device pass; -m " << o.minStep << " recomputes them from the input files (same result).\n";
    logStream.flush();

    if (o.devices.size() > 1) return runOnSeveralDevices(o, base);
    if (o.devices.size() == 1) o.device = o.devices[0];
    sage2gpu_ctx *ctx = nullptr;
    if (sage2gpu_create(&ctx, o.device) != 0)
        printError("CUDA", "no usable CUDA device " + std::to_string(o.device) + " (libsage2gpu has no CPU fallback)");
    sage2gpu_counters cnt;

    // ---- step 1 -------------------------------------------------------------------------------------
    banner("STEP 1", "                                          organizing reads");
    auto t0 = std::chrono::steady_clock::now();
    {
        Uploader up(ctx, o.minOverlap);
        if (!o.listInput.empty()) loadFromList(up, o.listInput);
        else readDataset(up, o.fileInput, "");
        logStream << "Parsed and uploaded in " << seconds_since(t0) << " sec.\n";
        auto t1 = std::chrono::steady_clock::now();
        up.finish();
        logStream << "Function organizeReads() in " << seconds_since(t1) << " sec.\n";
    }
    sage2gpu_get_counters(ctx, &cnt);
    logStream << std::setw(20) << "Total reads: " << cnt.total_reads << "\n";
    logStream << std::setw(20) << "Good reads: " << cnt.good_reads << "\n";
    logStream << std::setw(20) << "Bad reads: " << cnt.total_reads - cnt.good_reads << "\n";
    logStream << "\t" << std::setw(21) << "Average read length: " << cnt.avg_len << "\n";
    logStream << "\tNumber of unique reads: " << cnt.unique_reads << "\n";
    logStream.flush();
    const bool saveReads = o.maxStep <= 3 || o.saveAll || o.maxStep > 3;      // steps 4-7 continue from the files
    if (saveReads) {
        auto t = std::chrono::steady_clock::now();
        if (sage2gpu_write_reads(ctx, (base + ".reads").c_str()) != 0) printError("OPEN_FILE", sage2gpu_last_error(ctx));
        logStream << "Function saveReadsInFile in " << seconds_since(t) << " sec.\n";
    }
    if (o.maxStep == 1) { sage2gpu_destroy(ctx); return 0; }

    // ---- step 2 -------------------------------------------------------------------------------------
    banner("STEP 2", "                                         building hash table");
    t0 = std::chrono::steady_clock::now();
    if (sage2gpu_build_hash_table(ctx) != 0) printError("CUDA", sage2gpu_last_error(ctx));
    sage2gpu_get_counters(ctx, &cnt);
    logStream << "\tHash string length: " << cnt.hash_len << "\n";
    logStream << "\tSize of hash table: " << cnt.table_capacity << " slots, " << cnt.distinct_keys << " distinct keys\n";
    logStream << "\tNumber of hashes over threshold: " << cnt.keys_over_threshold << "\n";
    logStream << "Function hashPrefixesAndSuffix() in " << seconds_since(t0) << " sec.\n";
    if (o.maxStep == 2 || o.saveAll)
        logStream << "NOTE: " << o.prefixName << ".hashTable is not written (see sage2gpu_main.cpp header).\n";
    logStream.flush();
    if (o.maxStep == 2) { sage2gpu_destroy(ctx); return 0; }

    // ---- step 3 -------------------------------------------------------------------------------------
    banner("STEP 3", "                                      building overlap graph");
    t0 = std::chrono::steady_clock::now();
    if (sage2gpu_build_overlap_graph(ctx) != 0) printError("CUDA", sage2gpu_last_error(ctx));
    sage2gpu_get_counters(ctx, &cnt);
    logStream << "\tTotal reads contained by extension: " << cnt.contained_ext << "\n";
    logStream << "\tTotal reads contained by size: " << cnt.contained_size << "\n";
    logStream << "\tTotal reads left to explore: " << cnt.left_to_explore << "\n";
    logStream << "\tTotal edges inserted: " << cnt.edges_inserted_c << "\n";
    logStream << "\tTotal transitive edges removed: " << cnt.transitive_removed << "\n";
    logStream << "\tEdges in the overlap graph: " << cnt.n_edges << "\n";
    logStream << "Functions buildInitialOverlapGraph() + buildOverlapGraphEconomy() + sortEconomyGraph() in " << seconds_since(t0) << " sec.\n";
    sage2gpu_timers tm;
    sage2gpu_get_timers(ctx, &tm);
    logStream << "\tDevice ms: ingest " << tm.ingest << ", sort " << tm.sort_reads << ", table " << tm.build_table << ", phase A "
              << tm.phase_a << ", phase B " << tm.phase_b << ", phase C " << tm.phase_c_dev << " + host " << tm.phase_c_host
              << ", edges " << tm.sort_edges << "\n";
    {
        auto t = std::chrono::steady_clock::now();
        if (sage2gpu_write_graph3(ctx, (base + ".graph3").c_str()) != 0) printError("OPEN_FILE", sage2gpu_last_error(ctx));
        logStream << "Function saveOverlapGraphInFile in " << seconds_since(t) << " sec.\n";
    }
    logStream.flush();
    sage2gpu_destroy(ctx);
    if (o.maxStep <= 3) return 0;

    // ---- steps 4-7: the reference's own, from the files above --------------------------------------------
    if (o.referenceBin.empty()) {
        logStream << "Steps 4-" << o.maxStep << " are the reference's: run `SAGE2 -m 4 -i " << o.prefixName << " ...` (or give --reference).\n";
        std::cout << "sage2gpu: steps 1-3 done (" << base << ".reads, " << base << ".graph3); continue with SAGE2 -m 4 -i " << o.prefixName << "\n";
        return 0;
    }
    logStream << "Running " << o.referenceBin << " -m 4 for steps 4-" << o.maxStep << ".\n";
    logStream.close();
    return runReference(o, 4);
}
