// phase_c_bench.cpp -- times the host phase-C walk (sage2_b200/csrc/host_phase_c.cpp) on a dump of its real input
// (written by libsage2gpu when SAGE2GPU_DUMP_PHASE_C=<file> is set, see tools/dump_phase_c.py) and checks the result
// against the output stored in the same dump.
//   g++ -O2 -std=c++17 -pthread -o /tmp/phase_c_bench tools/phase_c_bench.cpp sage2_b200/csrc/host_phase_c.cpp
//   /tmp/phase_c_bench gpurun_out/phasec_cfg4.bin [repeats]
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "../sage2_b200/csrc/host_phase_c.h"
#include "../tests/phase_c_lists_ref.h"

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: phase_c_bench <dump> [repeats]\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    uint64_t hdr[6];
    if (fread(hdr, 8, 6, f) != 6) return 1;
    const uint64_t nS = hdr[0], nC = hdr[1], nB = hdr[2], nOut = hdr[3];
    std::vector<uint32_t> s_ids(nS), off(nS + 1), selLen(nB);
    std::vector<uint16_t> s_len(nS);
    std::vector<uint64_t> cand(nC), selB(2 * nB), want(nOut);
    bool ok = fread(s_ids.data(), 4, nS, f) == nS && fread(s_len.data(), 2, nS, f) == nS && fread(off.data(), 4, nS + 1, f) == nS + 1 &&
              fread(cand.data(), 8, nC, f) == nC && fread(selB.data(), 8, 2 * nB, f) == 2 * nB && fread(selLen.data(), 4, nB, f) == nB &&
              fread(want.data(), 8, nOut, f) == nOut;
    fclose(f);
    if (!ok) { fprintf(stderr, "short dump\n"); return 1; }
    sg::PhaseCInput in;
    in.nS = nS; in.s_ids = s_ids.data(); in.s_len = s_len.data(); in.cand_off = off.data(); in.cand = cand.data();
    in.nB = nB; in.edgesB = selB.data(); in.edgesB_len = selLen.data();
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        sg::PhaseCOutput out;
        const float ms = sg::run_host_phase_c(in, out);
        best = std::min(best, ms);
        if (out.edges != want || out.inserted != hdr[4] || out.removed != hdr[5]) {
            fprintf(stderr, "MISMATCH: %zu words (want %zu), inserted %llu (want %llu), removed %llu (want %llu)\n", out.edges.size(), want.size(),
                    (unsigned long long)out.inserted, (unsigned long long)hdr[4], (unsigned long long)out.removed, (unsigned long long)hdr[5]);
            return 3;
        }
    }
    // the traversal alone, as libsage2gpu runs it for asymmetric candidate sets: from the lists the device prepares (built here
    // on the CPU, tests/phase_c_lists_ref.h).  It must explore in the same order as the walk.
    {
        std::vector<uint32_t> cnode(nC);
        auto node_of = [&](uint32_t id1) { return (uint32_t)(std::lower_bound(s_ids.begin(), s_ids.end(), id1 - 1) - s_ids.begin()); };
        auto in_s = [&](uint32_t id1) { const uint32_t n = node_of(id1); return n < nS && s_ids[n] == id1 - 1; };
        for (uint64_t q = 0; q < nC; ++q) cnode[q] = node_of((uint32_t)(cand[q] >> 32));
        pc_ref::Lists lists = pc_ref::build(in, cnode);
        for (uint64_t e = 0; e < nB; ++e) {
            const uint32_t a = (uint32_t)(selB[2 * e] >> 32), b = (uint32_t)selB[2 * e];
            if (in_s(a)) lists.has_b[node_of(a)] = 1;
            if (in_s(b)) lists.has_b[node_of(b)] = 1;
        }
        std::vector<uint32_t> order1, order2;
        float b1 = 1e30f, b2 = 1e30f;
        for (int r = 0; r < reps; ++r) b1 = std::min(b1, sg::run_host_phase_c_order(in, order1));
        for (int r = 0; r < reps; ++r) b2 = std::min(b2, sg::run_host_phase_c_order_lists(lists.view(nS), order2));
        printf("traversal only: walk %.2f ms, from the sorted lists %.2f ms, orders %s\n", b1, b2, order1 == order2 ? "identical" : "DIFFER");
        if (order1 != order2) return 4;
    }
    printf("nS %llu  candidates %llu  phase-B records %llu  output words %llu : best of %d = %.2f ms, identical output\n",
           (unsigned long long)nS, (unsigned long long)nC, (unsigned long long)nB, (unsigned long long)nOut, reps, best);
    return 0;
}
