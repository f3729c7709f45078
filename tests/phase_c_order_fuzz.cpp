// phase_c_order_fuzz.cpp -- TEST INFRASTRUCTURE.  Random candidate graphs for the two host traversals of phase C
// (sage2_b200/csrc/host_phase_c.cpp): the walk itself (run_host_phase_c_order: id map, insertions into both lists, sort at
// every exploration -- the transcription of economyGraph.cpp:513-638) and the traversal over the lists the device prepares
// (run_host_phase_c_order_lists; the lists are built here by tests/phase_c_lists_ref.h, the CPU restatement of
// phase_c_sorted_lists).  They must explore the reads in the same order on ANY input: asymmetric candidate sets, several
// candidates per pair, candidates of a read with itself, reads without candidates, with and without phase-B records.
//   usage: phase_c_order_fuzz <cases> <seed>      exit 0 = all orders identical
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <random>
#include <vector>
#include "../sage2_b200/csrc/host_phase_c.h"
#include "phase_c_lists_ref.h"

int main(int argc, char **argv)
{
    const int cases = argc > 1 ? atoi(argv[1]) : 200;
    std::mt19937_64 rng(argc > 2 ? strtoull(argv[2], nullptr, 10) : 1);
    auto rnd = [&](uint64_t n) { return (uint32_t)(rng() % n); };
    unsigned long long total_nodes = 0, total_cands = 0;
    for (int cs = 0; cs < cases; ++cs) {
        const uint32_t nS = 1 + rnd(cs % 5 == 0 ? 2000 : 60);
        const uint32_t U = nS + 1 + rnd(3 * nS + 5);                     // ids 1..U; S is a random subset
        std::vector<uint32_t> ids(U);
        for (uint32_t i = 0; i < U; ++i) ids[i] = i;
        std::shuffle(ids.begin(), ids.end(), rng);
        std::vector<uint32_t> s_ids(ids.begin(), ids.begin() + nS);      // 0-based, ascending
        std::sort(s_ids.begin(), s_ids.end());
        std::vector<uint8_t> in_s(U, 0);
        for (uint32_t x : s_ids) in_s[x] = 1;
        std::vector<uint16_t> s_len(nS);
        for (auto &l : s_len) l = (uint16_t)(40 + rnd(111));
        // candidates: density, symmetry and multiplicity vary from case to case
        const uint32_t max_deg = 1 + rnd(cs % 7 == 0 ? 40 : 8);
        const bool mostly_symmetric = cs % 3 == 0;
        std::vector<std::vector<uint64_t>> lists(nS);
        auto rev = [](uint32_t t) { return t == 0 ? 3u : (t == 3 ? 0u : t); };
        for (uint32_t s = 0; s < nS; ++s) {
            const uint32_t deg = rnd(max_deg + 1);
            for (uint32_t k = 0; k < deg; ++k) {
                const uint32_t p = rnd(8) == 0 ? s : rnd(nS);            // sometimes the read itself
                const uint32_t t = rnd(4);
                const uint32_t d = rnd(10) == 0 ? ((0u - (1 + rnd(30))) & 0xFFFFFu) : 1 + rnd(s_len[s] - 1);     // sometimes negative (20-bit two's complement)
                lists[s].push_back(((uint64_t)(s_ids[p] + 1) << 32) | ((uint64_t)t << 20) | d);
                if (mostly_symmetric && p != s && rnd(10) != 0) {
                    uint32_t ds = d;
                    if (ds & 0x80000u) ds |= 0xFFF00000u;
                    const uint32_t d2 = ((uint32_t)s_len[s] - ((uint32_t)s_len[p] - ds)) & 0xFFFFFu;
                    lists[p].push_back(((uint64_t)(s_ids[s] + 1) << 32) | ((uint64_t)rev(t) << 20) | d2);
                }
                if (rnd(6) == 0) lists[s].push_back(lists[s].back());    // the same candidate twice
            }
        }
        std::vector<uint32_t> off(nS + 1, 0);
        std::vector<uint64_t> cand;
        for (uint32_t s = 0; s < nS; ++s) {
            std::shuffle(lists[s].begin(), lists[s].end(), rng);
            cand.insert(cand.end(), lists[s].begin(), lists[s].end());
            off[s + 1] = (uint32_t)cand.size();
        }
        // phase-B records between an S read and a read outside S (the only kind that touches S: both end points of a
        // phase-B edge between two S reads would have left S)
        std::vector<uint64_t> selB;
        std::vector<uint32_t> selLen;
        std::vector<uint8_t> has_b(nS, 0);
        std::vector<uint32_t> outside;
        for (uint32_t i = 0; i < U; ++i) if (!in_s[i]) outside.push_back(i);
        if (!outside.empty() && cs % 2 == 0)
            for (uint32_t s = 0; s < nS; ++s)
                if (rnd(3) == 0) {
                    const uint32_t o = outside[rnd(outside.size())];
                    const uint32_t a = std::min(s_ids[s], o) + 1, b = std::max(s_ids[s], o) + 1;
                    selB.push_back(((uint64_t)a << 32) | b);
                    selB.push_back(((uint64_t)rnd(4) << 20) | (1 + rnd(30)));
                    selLen.push_back(100u | (100u << 16));
                    has_b[s] = 1;
                }
        sg::PhaseCInput in;
        in.nS = nS; in.s_ids = s_ids.data(); in.s_len = s_len.data(); in.cand_off = off.data(); in.cand = cand.data();
        in.nB = selB.size() / 2; in.edgesB = selB.data(); in.edgesB_len = selLen.data();
        std::vector<uint32_t> cnode(cand.size());
        for (size_t q = 0; q < cand.size(); ++q)
            cnode[q] = (uint32_t)(std::lower_bound(s_ids.begin(), s_ids.end(), (uint32_t)(cand[q] >> 32) - 1) - s_ids.begin());
        pc_ref::Lists L = pc_ref::build(in, cnode);
        L.has_b = has_b;
        std::vector<uint32_t> o1, o2;
        sg::run_host_phase_c_order(in, o1);
        sg::run_host_phase_c_order_lists(L.view(nS), o2);
        if (o1 != o2) {
            uint32_t first = 0;
            while (first < nS && o1[first] == o2[first]) ++first;
            fprintf(stderr, "case %d (nS %u, %zu candidates, %zu phase-B records): orders differ, first at S index %u (%u vs %u)\n", cs, nS, cand.size(),
                    selB.size() / 2, first, o1[first], o2[first]);
            return 1;
        }
        total_nodes += nS; total_cands += cand.size();
    }
    printf("%d random candidate graphs, %llu reads, %llu candidates: the two traversals explore in the same order\n", cases, total_nodes, total_cands);
    return 0;
}
