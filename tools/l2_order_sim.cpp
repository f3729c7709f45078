// l2_order_sim.cpp -- ANALYSIS AID (no product code path): how many of the search kernel's 128-byte line requests would hit
// an LRU cache of a given size when the reads are processed (a) in id order, as today, (b) in the order of a min-hash of
// their window keys.  Input: the unique reads of a data set as core.cuh records.
//   built and driven by tools/l2_order_sim.py
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <list>
#include <unordered_map>
#include <vector>
#include "../sage2_b200/csrc/core.cuh"

using namespace sg;

struct Lru {
    size_t cap;
    std::list<u64> order;
    std::unordered_map<u64, std::list<u64>::iterator> pos;
    u64 hits = 0, misses = 0;
    explicit Lru(size_t c) : cap(c) {}
    void touch(u64 line)
    {
        auto it = pos.find(line);
        if (it != pos.end()) { order.splice(order.begin(), order, it->second); ++hits; return; }
        ++misses;
        order.push_front(line);
        pos[line] = order.begin();
        if (order.size() > cap) { pos.erase(order.back()); order.pop_back(); }
    }
};

struct KeyHash { size_t operator()(const std::pair<u64, u64> &k) const { return (size_t)hash_key(k.first, k.second); } };

extern "C" void l2sim(const u64 *F, const u64 *RC, const uint16_t *len, u64 U, int SW, int k, u64 cache_lines, u64 inflight, double *out /*[8]*/)
{
    const int h = hash_len_for(k);
    std::unordered_map<std::pair<u64, u64>, std::vector<u32>, KeyHash> table;
    table.reserve(4 * U);
    for (u64 i = 0; i < U; ++i)
        for (int t = 0; t < 4; ++t) {
            u64 v0, v1;
            entry_key(F + i * SW, RC + i * SW, SW, len[i], h, t, v0, v1);
            table[std::make_pair(v0, v1)].push_back((u32)(i * 4 + t));
        }
    const u64 nsec = (4 * U + 2 * U + 3) / 4, slot_lines = nsec / 4 + 1;     // 1.5 x 4U slots, 4 per sector, 4 sectors per line
    std::vector<u64> minhash(U);
    for (u64 i = 0; i < U; ++i) {
        u64 m = ~0ull;
        for (int j = 0; j + h <= (int)len[i]; ++j) {
            u64 v0, v1;
            extract_key(F + i * SW, SW, j, h, v0, v1);
            m = std::min(m, hash_key(v0, v1));
        }
        minhash[i] = m;
    }
    for (int mode = 0; mode < 2; ++mode) {
        std::vector<u32> order(U);
        for (u64 i = 0; i < U; ++i) order[i] = (u32)i;
        if (mode == 1) std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return minhash[a] < minhash[b]; });
        Lru slot_c(cache_lines);      // one cache for everything, statistics by kind
        u64 slot_hits = 0, slot_req = 0, rec_hits = 0, rec_req = 0;
        // `inflight` reads advance together, 32 windows at a time each (the kernel's resident warps), then the next group
        for (u64 x0 = 0; x0 < U; x0 += inflight)
        for (int jb = 0; jb < 1024; jb += 32) {
          bool any = false;
          for (u64 x = x0; x < U && x < x0 + inflight; ++x) {
            const u64 i = order[x];
            for (int j = jb; j < jb + 32 && j + h <= (int)len[i]; ++j) {
                any = true;
                u64 v0, v1;
                extract_key(F + i * SW, SW, j, h, v0, v1);
                const u64 hsh = hash_key(v0, v1);
                const u64 before = slot_c.hits;
                slot_c.touch(home_sector(hsh, nsec) / 4);
                slot_hits += slot_c.hits - before; ++slot_req;
                auto it = table.find(std::make_pair(v0, v1));
                if (it == table.end() || it->second.size() >= (size_t)kHashThreshold) continue;
                for (u32 ent : it->second) {
                    const u64 r2 = ent >> 2;
                    if (r2 == i) continue;
                    const u64 line = slot_lines + (partner_uses_rc((int)(ent & 3)) ? U / 2 + 1 : 0) + r2 / 2;     // 64-byte records, two per line
                    const u64 b2 = slot_c.hits;
                    slot_c.touch(line);
                    rec_hits += slot_c.hits - b2; ++rec_req;
                }
            }
          }
          if (!any) break;
        }
        out[4 * mode + 0] = (double)slot_req; out[4 * mode + 1] = (double)slot_hits;
        out[4 * mode + 2] = (double)rec_req; out[4 * mode + 3] = (double)rec_hits;
    }
}
