// sage2gpuShim.cpp -- the reference-side binding of libsage2gpu: SAGE2's OWN main.cpp, unmodified, runs its steps 1-3
// on the GPU and hands steps 4-7 the very objects they expect.
//
// The reference has no plugin interface; its seam is the object protocol main() drives (main.cpp:44-49, 76-77,
// 108-117).  This translation unit is what a SAGE2 maintainer adds to the tree: it REDEFINES the six member functions
// through which main() enters the hot path and leaves every other member, every header and main.cpp as they are:
//
//   ReadLoader::insertReadIntoList      readLoader.cpp:179-213   a good read is appended to a pinned chunk and streamed to
//                                                                the device instead of being packed on the host
//   ReadLoader::organizeReads           readLoader.cpp:215-260   filter / pack / sort / dedupe / revcomp on the device;
//                                                                readsList[1..U] is then filled in the reference's layout
//                                                                (one malloc per strand: ~ReadLoader frees them one by one)
//   HashTable::hashPrefixesAndSuffix    hashTable.cpp:70-128     the prefix/suffix table is built in HBM
//   EconomyGraph::buildInitialOverlapGraph  economyGraph.cpp:37-489   phases A, B, C and the canonical sort on the device;
//                                                                economyGraphList[1..U] is filled with individually
//                                                                malloc'ed, compareIdBased-sorted lists, which the
//                                                                UNMODIFIED OverlapGraph::convertGraph consumes and frees
//   EconomyGraph::buildOverlapGraphEconomy, sortEconomyGraph  economyGraph.cpp:495-574, 896-913   nothing left to do
//
// Build (oracle/Makefile, target ref_gpu): the reference's readLoader.o / hashTable.o / economyGraph.o are compiled from
// the sources where they lie and the six symbols above are made weak in them (objcopy --weaken-symbol), so this file's
// definitions win at link time; in a real tree the maintainer deletes the six bodies instead.  Link line: -lsage2gpu.
// Errors follow the reference's convention: printError -> log line + exit (utils.cpp:36-40).  No CPU fallback.
#include <string.h>

#include <algorithm>
#include <vector>

#include "economyGraph/economyGraph.h"
#include "economyGraph/hashTable.h"
#include "inputReader/readLoader.h"
#include "sage2gpu.h"
#include "utils.h"

extern ofstream logStream;
extern uint64_t averageReadLength;

namespace {

// one context for the whole run (steps 1-3 of one main())
sage2gpu_ctx *g_ctx = NULL;

// two pinned chunks: the parser fills one while the previous one travels (sage2gpu_load_append)
const uint64_t kChunkBases = 64ull << 20, kChunkReads = 1ull << 20;
struct Chunk { uint8_t *bases; int64_t *offsets; uint64_t n_reads, n_bases; };
Chunk g_chunk[2];
int g_cur = 0;
bool g_loading = false;

void fail(const char *what)
{
    printError(OTHERS, string("libsage2gpu: ") + what + ": " + (g_ctx ? sage2gpu_last_error(g_ctx) : "no CUDA device"));
}
void check(int rc, const char *what) { if (rc != SAGE2GPU_OK) fail(what); }

void begin_loading(int minOverlap)
{
    if (!g_ctx && sage2gpu_create(&g_ctx, 0) != SAGE2GPU_OK) fail("sage2gpu_create");
    for (int i = 0; i < 2; ++i) {
        g_chunk[i].bases = (uint8_t *)sage2gpu_host_alloc(kChunkBases);
        g_chunk[i].offsets = (int64_t *)sage2gpu_host_alloc((kChunkReads + 1) * sizeof(int64_t));
        if (!g_chunk[i].bases || !g_chunk[i].offsets) printError(MEM_ALLOC, "pinned upload chunk");
        g_chunk[i].n_reads = g_chunk[i].n_bases = 0;
        g_chunk[i].offsets[0] = 0;
    }
    check(sage2gpu_load_begin(g_ctx, minOverlap), "sage2gpu_load_begin");
    g_loading = true;
}

void flush_chunk()
{
    Chunk &c = g_chunk[g_cur];
    if (c.n_reads) check(sage2gpu_load_append(g_ctx, c.bases, c.offsets, (int64_t)c.n_reads), "sage2gpu_load_append");
    g_cur ^= 1;                 // the other chunk's copy was waited for by the call above
    g_chunk[g_cur].n_reads = g_chunk[g_cur].n_bases = 0;
}

uint8_t *dup_bytes(const uint8_t *src, size_t n)
{
    uint8_t *p = (uint8_t *)malloc(n ? n : 1);
    if (!p) printError(MEM_ALLOC, "read");
    memcpy(p, src, n);
    return p;
}

}  // namespace

// ---- step 1 ---------------------------------------------------------------------------------------------------------
// Called by the reference's own parse loop for every read that passed `size > minOverlap` and isGoodRead
// (readLoader.cpp:146-160).  Orientation, packing and the list itself are the device's job now; the counters main() and
// the log use are kept exactly (numberOfReads, totalBP -> averageReadLength, sizeOfList).
void ReadLoader::insertReadIntoList(string &read1)
{
    if (!g_loading) begin_loading(minOverlap);
    if (sizeOfList == 0) sizeOfList = 1000000;
    if (numberOfReads >= sizeOfList - 10) sizeOfList += 1000000;
    if (read1.size() > kChunkBases) printError(OTHERS, "read longer than the upload chunk");
    Chunk *c = &g_chunk[g_cur];
    if (c->n_reads == kChunkReads || c->n_bases + read1.size() > kChunkBases) { flush_chunk(); c = &g_chunk[g_cur]; }
    memcpy(c->bases + c->n_bases, read1.data(), read1.size());
    c->n_bases += read1.size();
    c->offsets[++c->n_reads] = (int64_t)c->n_bases;
    numberOfReads++;
    totalBP += read1.size();
}

void ReadLoader::organizeReads()
{
    logStream << "\nIn function organizeReads().\n";
    logStream.flush();
    time_t seconds_s = time(NULL);
    if (!g_loading) begin_loading(minOverlap);        // an input without a single good read
    flush_chunk();
    check(sage2gpu_load_finish(g_ctx), "sage2gpu_load_finish");
    g_loading = false;
    for (int i = 0; i < 2; ++i) { sage2gpu_host_free(g_chunk[i].bases); sage2gpu_host_free(g_chunk[i].offsets); }

    sage2gpu_counters n;
    check(sage2gpu_get_counters(g_ctx, &n), "sage2gpu_get_counters");
    if (n.good_reads != numberOfReads || n.total_bp != totalBP) printError(OTHERS, "libsage2gpu: read filter disagrees with isGoodRead");
    logStream << "\tQuicksort reads finished in " << time(NULL) - seconds_s << " sec.\n";
    seconds_s = time(NULL);

    const uint64_t U = n.unique_reads;
    uint64_t nb = 0;
    check(sage2gpu_reads_bytes(g_ctx, &nb), "sage2gpu_reads_bytes");
    std::vector<uint16_t> len(U), freq(U);
    std::vector<uint64_t> off(U + 1);
    std::vector<uint8_t> fwd(nb), rc(nb);
    check(sage2gpu_get_reads(g_ctx, len.data(), freq.data(), off.data(), fwd.data(), rc.data()), "sage2gpu_get_reads");
    free(readsList);
    if ((readsList = (Read *)malloc((U + 1) * sizeof(Read))) == NULL) printError(MEM_ALLOC, "readsList");
    logStream << "\tRemoving duplicate reads finished in " << time(NULL) - seconds_s << " sec.\n";
    seconds_s = time(NULL);
    uint64_t i;
    #pragma omp parallel for
    for (i = 1; i <= U; i++) {
        const size_t bytes = (size_t)(off[i] - off[i - 1]);
        readsList[i].frequency = freq[i - 1];
        readsList[i].length = len[i - 1];
        readsList[i].readInt = dup_bytes(fwd.data() + off[i - 1], bytes);
        readsList[i].readReverseInt = dup_bytes(rc.data() + off[i - 1], bytes);
    }
    numberOfUniqueReads = U;
    logStream << "\tNumber of unique reads: " << numberOfUniqueReads << "\n";
    logStream << "\tComputing reverse complements finished in " << time(NULL) - seconds_s << " sec.\n";
    logStream.flush();
}

// ---- step 2 ---------------------------------------------------------------------------------------------------------
void HashTable::hashPrefixesAndSuffix()
{
    time_t second_s = time(NULL);
    logStream << "In function hashPrefixesAndSuffix().\n";
    logStream.flush();
    hashThreshold = 100;
    longHash = loaderObj->numberOfUniqueReads + 100;
    hashStringLength = minOverlap > 64 ? 64 : minOverlap;
    logStream << "\t         Hash string length: " << hashStringLength << "\n";
    if (!g_ctx) printError(OTHERS, "libsage2gpu: the reads were not organised by this process (-m 2/3 restarts need the reference's own classes)");
    check(sage2gpu_build_hash_table(g_ctx), "sage2gpu_build_hash_table");
    sage2gpu_counters n;
    check(sage2gpu_get_counters(g_ctx, &n), "sage2gpu_get_counters");
    // sizeOfHashTable stays 0 and hashTableList NULL: the table lives in HBM, ~HashTable has nothing to free
    logStream << "\t            Hash table size: " << n.table_capacity << "\n";
    logStream << "\t Number of hash elements over threshold: " << n.keys_over_threshold << "\n";
    logStream << "\t                        Total hash miss: " << hashMiss << "\n";
    logStream << "Function hashPrefixesAndSuffix() in " << time(NULL) - second_s << " sec.\n";
    logStream.flush();
}

// ---- step 3 ---------------------------------------------------------------------------------------------------------
static sage2gpu_counters g_graph_counters;

void EconomyGraph::buildInitialOverlapGraph()
{
    logStream << "In function buildInitialOverlapGraph().\n";
    logStream.flush();
    time_t seconds_s = time(NULL);
    if (!g_ctx) printError(OTHERS, "libsage2gpu: no table in this process");
    check(sage2gpu_build_overlap_graph(g_ctx), "sage2gpu_build_overlap_graph");
    check(sage2gpu_get_counters(g_ctx, &g_graph_counters), "sage2gpu_get_counters");
    const sage2gpu_counters &n = g_graph_counters;
    const uint64_t U = loaderObj->numberOfUniqueReads;

    uint64_t E = 0;
    check(sage2gpu_get_edges(g_ctx, NULL, 0, &E), "sage2gpu_get_edges");
    std::vector<sage2gpu_edge> e(E);
    if (E) check(sage2gpu_get_edges(g_ctx, e.data(), E, &E), "sage2gpu_get_edges");
    sage2gpu_destroy(g_ctx);          // the device's part is over (main.cpp:112 deletes the table at this point too)
    g_ctx = NULL;

    // economyGraphList[1..U]: what insertEdgeEconomy (economyGraph.cpp:813-849) leaves after sortEconomyGraph: per node
    // one malloc'ed array, [0].readId = count, entries in compareIdBased order.  Each undirected edge appears in both
    // end points' lists; convertGraph (overlapGraph.cpp:84-115) reads the `to > from` half and frees every list.
    if ((economyGraphList = (EconomyEdge **)malloc((U + 1) * sizeof(EconomyEdge *))) == NULL) printError(MEM_ALLOC, "graphEconomy");
    std::vector<uint32_t> deg(U + 2, 0);
    for (uint64_t x = 0; x < E; ++x) { deg[e[x].from]++; deg[e[x].to]++; }
    uint64_t i;
    #pragma omp parallel for
    for (i = 1; i <= U; i++) {
        economyGraphList[i] = NULL;
        if (deg[i]) {
            if ((economyGraphList[i] = (EconomyEdge *)malloc((deg[i] + 1) * sizeof(EconomyEdge))) == NULL) printError(MEM_ALLOC, "economy list");
            economyGraphList[i][0].readId = 0;
        }
    }
    economyGraphList[0] = NULL;
    for (uint64_t x = 0; x < E; ++x) {
        const sage2gpu_edge &g = e[x];
        EconomyEdge *lu = economyGraphList[g.from], *lv = economyGraphList[g.to];
        lu[0].readId = lu[0].readId + 1;
        lu[lu[0].readId] = EconomyEdge(g.to, (uint8_t)g.type, 0, g.delta);
        lv[0].readId = lv[0].readId + 1;
        lv[lv[0].readId] = EconomyEdge(g.from, reverseEdgeType((uint8_t)g.type), 0, g.delta_twin);
    }
    #pragma omp parallel for
    for (i = 1; i <= U; i++)
        if (economyGraphList[i] != NULL && economyGraphList[i][0].readId > 1)
            sort(economyGraphList[i] + 1, economyGraphList[i] + economyGraphList[i][0].readId + 1, compareIdBased);

    logStream << "     Total contained by extension: " << n.contained_ext << "\n";
    logStream << "          Total contained by size: " << n.contained_size << "\n";
    logStream << "            Total left to explore: " << n.left_to_explore << "\n";
    logStream << "Function buildInitialOverlapGraph() in " << time(NULL) - seconds_s << " sec.\n";
    logStream.flush();
}

void EconomyGraph::buildOverlapGraphEconomy()
{
    logStream << "In function buildOverlapGraphEconomy().\n";
    logStream << "     Total edges inserted: " << g_graph_counters.edges_inserted_c << "\n";
    logStream << "Total number of hash miss: " << numberOfHashMiss << "\n";
    logStream << "  Transitive edge removed: " << g_graph_counters.transitive_removed << "\n";
    logStream << "Function buildOverlapGraphEconomy() in 0 sec.\n";
    logStream.flush();
}

void EconomyGraph::sortEconomyGraph()
{
    logStream << "\nIn function sortEdgesEconomy().\n";
    logStream << "Function sortEdgesEconomy() in 0 sec.\n";      // the lists were handed over in compareIdBased order
    logStream.flush();
}
