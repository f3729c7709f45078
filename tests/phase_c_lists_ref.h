// phase_c_lists_ref.h -- test infrastructure: what phase_c_device.cu (phase_c_sorted_lists) hands to the host traversal, built
// on the CPU from the same candidates.  Used by tests/host_emul.cpp and tools/phase_c_bench.cpp.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <vector>
#include "../sage2_b200/csrc/host_phase_c.h"

namespace pc_ref {

struct Lists {
    std::vector<uint32_t> off, ent;
    std::vector<uint8_t> has_b;
    sg::PhaseCLists view(uint64_t nS) const
    {
        sg::PhaseCLists l;
        l.nS = nS; l.off = off.data(); l.ent = ent.data(); l.has_b = has_b.data();
        return l;
    }
};

// cnode[q] = S index of candidate q's read2 (every candidate of phase C leads to an S read)
inline Lists build(const sg::PhaseCInput &in, const std::vector<uint32_t> &cnode)
{
    const uint32_t nS = (uint32_t)in.nS;
    auto rev = [](uint32_t t) { return t == 0 ? 3u : (t == 3 ? 0u : t); };
    struct Rec { uint32_t list; uint64_t key; };      // key: (overhang20 << 34 | other << 2 | type) << 1 | twin
    std::vector<Rec> recs;
    for (uint32_t s = 0; s < nS; ++s)
        for (uint32_t q = in.cand_off[s]; q < in.cand_off[s + 1]; ++q) {
            const uint64_t cw = in.cand[q];
            const uint32_t t = (uint32_t)(cw >> 20) & 3u, d20 = (uint32_t)(cw & 0xFFFFFu), sb = cnode[q];
            const uint32_t d = (d20 & 0x80000u) ? (d20 | 0xFFF00000u) : d20;
            const uint32_t d2 = ((uint32_t)in.s_len[s] - ((uint32_t)in.s_len[sb] - d)) & 0xFFFFFu;
            recs.push_back(Rec{ s, (((uint64_t)d20 << 34) | ((uint64_t)sb << 2) | t) << 1 });
            if (sb != s) recs.push_back(Rec{ sb, ((((uint64_t)d2 << 34) | ((uint64_t)s << 2) | rev(t)) << 1) | 1 });
        }
    std::stable_sort(recs.begin(), recs.end(), [](const Rec &x, const Rec &y) { return x.list != y.list ? x.list < y.list : x.key > y.key; });
    Lists L;
    L.off.assign((size_t)nS + 1, 0);
    L.ent.resize(recs.size());
    for (size_t i = 0; i < recs.size(); ++i) {
        L.ent[i] = ((uint32_t)(recs[i].key >> 3) << 1) | (uint32_t)(recs[i].key & 1);
        L.off[recs[i].list + 1]++;
    }
    for (uint32_t s = 0; s < nS; ++s) L.off[s + 1] += L.off[s];
    L.has_b.assign(nS, 0);
    return L;
}

}  // namespace pc_ref
