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
    printf("nS %llu  candidates %llu  phase-B records %llu  output words %llu : best of %d = %.2f ms, identical output\n",
           (unsigned long long)nS, (unsigned long long)nC, (unsigned long long)nB, (unsigned long long)nOut, reps, best);
    return 0;
}
