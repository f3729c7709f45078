// ref_harness.cpp -- TEST / BENCH INFRASTRUCTURE: our own main() around the UNMODIFIED reference classes
// for steps 1-3 (ReadLoader, HashTable, EconomyGraph, OverlapGraph::convertGraph), compiled against the
// sources where they lie under /root/reference (oracle/Makefile `make ref`).  It exists because the
// reference only times itself with time(NULL) at 1 s resolution and always mixes FASTQ parsing and
// text output into its steps.  Here the FASTQ is parsed first (reference InputReader, untimed), then
// the exact call sequence of main.cpp:44-118 is timed with a monotonic clock, no file I/O inside.
//
//   ref_steps123 <fastq> <k> [out_prefix]      prints one JSON line on stdout
#include <omp.h>
#include <time.h>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include "economyGraph/economyGraph.h"
#include "economyGraph/hashTable.h"
#include "inputReader/inputReader.h"
#include "inputReader/readLoader.h"
#include "overlapGraph/overlapGraph.h"
#include "utils.h"

ofstream logStream;                           // globals main.h defines (main.h:34-36)
uint64_t genomeSize = 0, averageReadLength = 0;

static double now()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv)
{
    if (argc < 3) { cerr << "usage: ref_steps123 <fastq> <k> [out_prefix]\n"; return 2; }
    const string fastq = argv[1];
    const uint16_t k = (uint16_t)atoi(argv[2]);
    const string out = argc > 3 ? argv[3] : "";
    logStream.open(out.empty() ? "/dev/null" : (out + ".log").c_str());

    double t = now();
    vector<string> seqs;
    {
        InputReader rd(fastq, "");
        uint64_t id = 0;
        while (rd.getNextRead(id)) { seqs.push_back(rd.read.sequence); id++; }
    }
    const double t_parse = now() - t;

    // step 1 (main.cpp:44-49; loop body of readLoader.cpp:145-160)
    t = now();
    ReadLoader *loader = new ReadLoader(k);
    for (size_t i = 0; i < seqs.size(); i++) {
        string read1 = seqs[i];
        if (read1.size() <= k) continue;
        if (isGoodRead(read1, k)) loader->insertReadIntoList(read1);
    }
    averageReadLength = loader->numberOfReads ? loader->totalBP / loader->numberOfReads : 0;
    const double t_pack = now() - t;
    t = now();
    loader->organizeReads();
    const double t_organize = now() - t;
    // step 2 (main.cpp:76-77)
    t = now();
    HashTable *hash = new HashTable(k, loader);
    hash->hashPrefixesAndSuffix();
    const double t_table = now() - t;
    // step 3 (main.cpp:108-118)
    t = now();
    EconomyGraph *eco = new EconomyGraph(k, hash);
    eco->buildInitialOverlapGraph();
    const double t_phase_ab = now() - t;
    t = now();
    eco->buildOverlapGraphEconomy();
    delete hash;
    const double t_phase_c = now() - t;
    t = now();
    eco->sortEconomyGraph();
    const double t_sort = now() - t;
    t = now();
    OverlapGraph *graph = new OverlapGraph(eco, loader);
    graph->convertGraph();
    delete eco;
    const double t_convert = now() - t;

    uint64_t edges = 0;
    for (uint64_t i = 1; i <= loader->numberOfUniqueReads; i++)
        for (Edge *u = graph->graph[i]; u != NULL; u = u->next)
            if (i <= u->ID) edges++;
    if (!out.empty()) {
        loader->saveReadsInFile(out + ".reads");
        graph->saveOverlapGraphInFile(out + ".graph3");
    }
    const double total = t_pack + t_organize + t_table + t_phase_ab + t_phase_c + t_sort + t_convert;
    cout << "{\"input_reads\": " << seqs.size() << ", \"good_reads\": " << loader->numberOfReads
         << ", \"unique_reads\": " << loader->numberOfUniqueReads << ", \"edges\": " << edges
         << ", \"threads\": " << omp_get_max_threads() << ", \"t_parse\": " << t_parse << ", \"t_pack\": " << t_pack
         << ", \"t_organize\": " << t_organize << ", \"t_table\": " << t_table << ", \"t_phase_ab\": " << t_phase_ab
         << ", \"t_phase_c\": " << t_phase_c << ", \"t_sort\": " << t_sort << ", \"t_convert\": " << t_convert
         << ", \"t_steps123\": " << total << "}" << endl;
    return 0;
}
