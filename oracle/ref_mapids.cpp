// ref_mapids.cpp -- TEST INFRASTRUCTURE: our own main() around the UNMODIFIED reference ReadLoader, compiled against
// the sources where they lie under /root/reference (oracle/Makefile `make ref`).  It builds the read list of
// <dataset.fastq> exactly as main.cpp:44-49 does and prints ReadLoader::getIdOfRead (readLoader.cpp:319-353) for
// every good read of <queries.fastq> ("bad" for reads isGoodRead rejects, the gate of matePair.cpp:176-179).
// tests/golden/make_golden.py turns its output into tests/golden/mapids.json, the pin of oracle/sgo_map_reads.
//
//   ref_mapids <dataset.fastq> <queries.fastq> <k>
#include <fstream>
#include <iostream>
#include <string>
#include "inputReader/inputReader.h"
#include "inputReader/readLoader.h"
#include "utils.h"

ofstream logStream;                           // globals main.h defines (main.h:34-36)
uint64_t genomeSize = 0, averageReadLength = 0;

int main(int argc, char **argv)
{
    if (argc < 4) { cerr << "usage: ref_mapids <dataset.fastq> <queries.fastq> <k>\n"; return 2; }
    const uint16_t k = (uint16_t)atoi(argv[3]);
    logStream.open("/dev/null");
    ReadLoader *loader = new ReadLoader(k);
    {
        InputReader rd(argv[1], "");
        uint64_t id = 0;
        while (rd.getNextRead(id)) {
            string read1 = rd.read.sequence;
            id++;
            if (read1.size() <= k) continue;
            if (isGoodRead(read1, k)) loader->insertReadIntoList(read1);
        }
    }
    loader->organizeReads();
    InputReader rq(argv[2], "");
    uint64_t id = 0;
    while (rq.getNextRead(id)) {
        string q = rq.read.sequence;
        id++;
        if (isGoodRead(q, k)) cout << loader->getIdOfRead(q) << "\n";
        else cout << "bad\n";
    }
    return 0;
}
