// host_emul.cpp -- TEST INFRASTRUCTURE.  Drives the product's own host/device primitives
// (sage2_b200/csrc/core.cuh) and its host phase-C walk (host_phase_c.cpp) sequentially on the CPU,
// with std::sort / std::map standing in for the CUDA radix sort and open-addressing index.  It lets
// the `-m "not gpu"` suite check the bit arithmetic, the extension state machine, phase B and the
// phase-C walk against the oracle without a GPU.  It is NOT a fallback: nothing in sage2_b200 links it.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <map>
#include <vector>
#include "../sage2_b200/csrc/core.cuh"
#include "../sage2_b200/csrc/host_phase_c.h"
#include "phase_c_lists_ref.h"

using namespace sg;

namespace {

struct Emu {
    int SW = 1, k = 0, h = 0;
    u64 total = 0, good = 0, total_bp = 0, U = 0;
    std::vector<u64> F, RC;
    std::vector<uint16_t> len, freq;
    std::vector<u64> extR, extL;
    std::vector<uint8_t> explored_a, explored_b;
    std::vector<u64> edges;     // (w0,w1) final
    std::map<std::pair<u64, u64>, std::vector<u32>> table;
    std::vector<uint8_t> flag5;
    std::vector<u32> cont_max;
    u64 over = 0, distinct = 0, compare_calls = 0, inserted = 0, removed = 0, contained = 0, contained_size = 0, slow_reads = 0, fast_reads = 0;
    // ---- the sharded table (emulation of csrc/shard.cu: same buffers, same wire format) ----
    int max_len = 1, tb_rank = 0, tb_world = 1;
    std::map<u64, std::pair<u64, u64>> by_hash;      // owned keys by their 64-bit hash (what a tag probe can see)
    u64 pa_lo = 0, pa_hi = 0, restarts = 0;
    std::vector<u64> rt_queries, rt_wslot, an_resp;
    std::vector<u32> rt_qmap, rt_wentries, rt_ids, an_entries;
    std::vector<uint8_t> rt_redo;
    u64 rt_n = 0, rt_first = 0, rt_counts[kMaxWorld] = {};
    int rt_wstride = 1, rt_what = 0, rt_world = 1, rt_state = 0;
    bool rt_exact = false, rt_is_list = false, rt_for_c = false, have_phase_b = false;
};

// table lookup of one window of one read of a scan: number of bucket entries (0: absent or masked) through `ents`,
// or -1 when the answer failed its proof (tag collision; routed tag probes only).  s = position in the scanned batch.
typedef std::function<int(u64 s, const u64 *Xf, int j, u64 v0, u64 v1, const u32 *&ents)> Lookup;

int code_of(uint8_t c) { return ((c >> 1) ^ (c >> 2)) & 3; }
bool valid_char(uint8_t c) { c &= 0xDF; return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

void pack_record(const uint8_t *s, int len, int SW, u64 *rec, bool rc)
{
    for (int w = 0; w < SW; ++w) rec[w] = 0;
    for (int p = 0; p < len; ++p) {
        const int c = rc ? 3 - code_of(s[len - 1 - p]) : code_of(s[p]);
        rec[p >> 5] |= (u64)c << (62 - 2 * (p & 31));
    }
    rec[SW - 1] |= (u64)len;
}

// reverse complements + K3, once the unique reads are complete
void finish_reads(Emu &e)
{
    const int SW = e.SW, h = e.h;
    const u64 U = e.U = e.len.size();
    e.RC.resize(U * SW);
    for (u64 i = 0; i < U; ++i) revcomp_record(&e.F[i * SW], &e.RC[i * SW], SW, e.len[i]);
    auto &table = e.table;
    table.clear();
    for (u64 i = 0; i < U; ++i)
        for (int t = 0; t < 4; ++t) {
            u64 v0, v1;
            entry_key(&e.F[i * SW], &e.RC[i * SW], SW, e.len[i], h, t, v0, v1);
            table[std::make_pair(v0, v1)].push_back((u32)(i * 4 + t));
        }
    e.distinct = table.size(); e.over = 0;
    for (auto &kv : table) if (kv.second.size() >= (size_t)kHashThreshold) e.over++;
}

void prepare(Emu &e, const uint8_t *bases, const int64_t *off, int64_t n, int k, int rank = 0, int world = 1)
{
    e.k = k; e.h = hash_len_for(k); e.total = (u64)n;
    const int h = e.h;
    int max_len = 1;
    for (int64_t r = 0; r < n; ++r) max_len = std::max<int64_t>(max_len, std::min<int64_t>(off[r + 1] - off[r], 32 * kMaxWords - 8));
    e.max_len = max_len;
    int SW = words_for_len(max_len);
    static const int kStrides[] = { 2, 3, 4, 5, 6, 8, 12, 16, 32 };
    for (int s : kStrides) if (s >= SW) { SW = s; break; }
    e.SW = SW;

    // K1
    std::vector<std::vector<u64>> recs;
    for (int64_t r = 0; r < n; ++r) {
        const int64_t l = off[r + 1] - off[r];
        if (l <= k || l > 32 * SW - 8) continue;
        const uint8_t *s = bases + off[r];
        bool ok = true;
        for (int p = 0; p < l; ++p) ok &= valid_char(s[p]);
        if (!ok) continue;
        std::vector<u64> f(SW), q(SW);
        pack_record(s, (int)l, SW, f.data(), false);
        pack_record(s, (int)l, SW, q.data(), true);
        recs.push_back(q < f ? q : f);
        e.good++; e.total_bp += (u64)l;
    }
    if (world > 1) {
        // reads.cu, stage_organize_reads with world > 1: quantile splitters of a 4096-bin histogram of the leading 6 bases
        const int nbins = 1 << 12;
        std::vector<u64> hist(nbins, 0);
        for (auto &r : recs) hist[r[0] >> 52]++;
        std::vector<u32> bound(world + 1, (u32)nbins);
        bound[0] = 0;
        u64 pre = 0;
        int r = 1;
        const u64 n_good = recs.size();
        for (int b = 0; b < nbins && r < world; ++b) {
            while (r < world && pre >= (n_good * (u64)r + (u64)world - 1) / (u64)world) bound[r++] = (u32)b;
            pre += hist[b];
        }
        std::vector<std::vector<u64>> mine;
        for (auto &rec : recs) { const u32 b = (u32)(rec[0] >> 52); if (b >= bound[rank] && b < bound[rank + 1]) mine.push_back(rec); }
        recs.swap(mine);
    }
    // K2
    std::sort(recs.begin(), recs.end());
    for (size_t i = 0; i < recs.size(); ++i) {
        if (i == 0 || recs[i] != recs[i - 1]) {
            e.F.insert(e.F.end(), recs[i].begin(), recs[i].end());
            e.len.push_back((uint16_t)rec_len(recs[i].data(), SW));
            e.freq.push_back(0);
        }
        e.freq.back()++;
    }
    e.U = e.len.size();
    if (world == 1) finish_reads(e);
}

Lookup local_table(Emu &e)
{
    return [&e](u64, const u64 *, int, u64 v0, u64 v1, const u32 *&ents) -> int {
        auto it = e.table.find(std::make_pair(v0, v1));
        if (it == e.table.end() || it->second.size() >= (size_t)kHashThreshold) return 0;
        ents = it->second.data();
        return (int)it->second.size();
    };
}

void alloc_phase_a(Emu &e, int rank, int world)
{
    const u64 U = e.U;
    const u64 chunk = partition_chunk(U, world), padded = chunk * (u64)world;
    e.pa_lo = std::min<u64>(U, (u64)rank * chunk); e.pa_hi = std::min<u64>(U, e.pa_lo + chunk);
    e.extR.assign(padded, 0); e.extL.assign(padded, 0);
    e.flag5.assign(padded, 0); e.cont_max.assign(padded, 0);
    e.rt_redo.assign(chunk, 0);
    e.compare_calls = 0; e.slow_reads = 0; e.restarts = 0;
}

// ---- the superstring fast path of phase_a_fast_kernel (search_fast.cu) ------------------------------------------------
// Per side a "superstring" S = read i (right side) or its reverse complement (left side) followed by the bases the hits
// seen so far agree on beyond its end.  Every gated item is compared ONCE against S: a mismatch inside read i's span
// means "not a hit", a mismatch beyond it means two hits disagree (anomaly).  Per round of 32 items the hit that reaches
// furthest extends S and the other hits of the round are checked on the part S did not cover before.  No anomaly and
// at most one hit per side and window: the reference's chain (economyGraph.cpp:94-437) reduces to "first right hit,
// last left hit, not ambiguous".  Anything else is left to the general scan (returns false, no side effects kept).
struct FastOut { u32 Rid, Rtype, Rlen, Lid, Ltype, Llen, connections; u64 calls; std::vector<std::pair<u32, u32>> cont; };

static int base_at(const u64 *rec, int p) { return (int)((rec[p >> 5] >> (62 - 2 * (p & 31))) & 3); }

template <typename ItemT>
bool fast_certify(const Emu &e, u64 i, const std::vector<ItemT> &items, FastOut &o, size_t cap_items, int max_windows)
{
    const int SW = e.SW, k = e.k, h = e.h;
    const u64 *Xf = &e.F[i * SW], *Xr = &e.RC[i * SW];
    const int len1 = e.len[i];
    if (len1 - h + 1 > max_windows || items.size() > cap_items) return false;
    std::vector<uint8_t> S[2];          // [0] right side: read i; [1] left side: its reverse complement
    for (int p = 0; p < len1; ++p) { S[0].push_back((uint8_t)base_at(Xf, p)); S[1].push_back((uint8_t)base_at(Xr, p)); }
    int lastJ[2] = { -1, -1 };
    bool hasR = false, hasL = false;
    int bestRj = 0, bestLj = 0;
    o = FastOut();
    o.Rid = o.Rtype = o.Rlen = o.Lid = o.Ltype = o.Llen = o.connections = 0; o.calls = 0;
    struct H { int jj, side, s, len2, type; u32 rid2; const u64 *Q; };
    std::vector<H> hits;
    // the kernel's order: the gated right items from the last window down, then the gated left items from the first up
    std::vector<ItemT> order;
    for (size_t x = items.size(); x-- > 0;) {
        const u32 rid2 = items[x].ent >> 2;
        const int type = (int)(items[x].ent & 3);
        if (!(type & 1) && rid2 != (u32)i && gate_right(items[x].jj, len1, k)) order.push_back(items[x]);
    }
    for (size_t x = 0; x < items.size(); ++x) {
        const u32 rid2 = items[x].ent >> 2;
        const int type = (int)(items[x].ent & 3);
        if ((type & 1) && rid2 != (u32)i && gate_left(items[x].jj, k, h)) order.push_back(items[x]);
    }
    for (size_t r0 = 0; r0 < order.size(); r0 += 32) {
        hits.clear();
        for (size_t x = r0; x < std::min(order.size(), r0 + 32); ++x) {
            const int j = order[x].jj;
            const u32 rid2 = order[x].ent >> 2;
            const int type = (int)(order[x].ent & 3);
            const bool right = !(type & 1);
            const u64 *Q = (partner_uses_rc(type) ? &e.RC[0] : &e.F[0]) + (u64)rid2 * SW;
            const int len2 = e.len[rid2], side = right ? 0 : 1, s = right ? j : len1 - j - h, xlen = len1 - s;
            o.calls++;
            const int ov = std::min(len2, (int)S[side].size() - s);
            int first_bad = -1;
            for (int t = 0; t < ov; ++t) if (base_at(Q, t) != S[side][s + t]) { first_bad = t; break; }
            const bool contained = len2 <= xlen;
            if (first_bad >= 0 && first_bad < std::min(xlen, len2)) continue;           // differs inside read i: not a hit
            if (contained) { o.cont.push_back(std::make_pair(rid2, (u32)(i + 1))); continue; }
            if (first_bad >= 0) return false;                                             // disagrees with an earlier hit
            hits.push_back(H{ j, side, s, len2, type, rid2, Q });
        }
        for (int side = 0; side < 2; ++side) {
            int best = -1, nside = 0;
            for (size_t x = 0; x < hits.size(); ++x) {
                if (hits[x].side != side) continue;
                nside++;
                if (hits[x].jj == lastJ[side]) return false;                              // two hits of one side in one window
                for (size_t y = 0; y < x; ++y) if (hits[y].side == side && hits[y].jj == hits[x].jj) return false;
                if (best < 0 || hits[x].s + hits[x].len2 > hits[best].s + hits[best].len2) best = (int)x;
            }
            if (!nside) continue;
            const int old = (int)S[side].size();
            const H &m = hits[best];
            for (int pos = old; pos < m.s + m.len2; ++pos) S[side].push_back((uint8_t)base_at(m.Q, pos - m.s));
            for (const H &q : hits) {
                if (q.side != side) continue;
                for (int t = std::max(0, old - q.s); t < q.len2; ++t) if (base_at(q.Q, t) != S[side][q.s + t]) return false;
                lastJ[side] = q.jj;
            }
        }
        for (const H &q : hits) {
            if (q.side == 0) { if (!hasR || q.jj < bestRj) { o.Rid = q.rid2 + 1; o.Rtype = (u32)(q.type >> 1); o.Rlen = (u32)(q.len2 - (len1 - q.jj)); hasR = true; bestRj = q.jj; } }
            else if (!hasL || q.jj > bestLj) { o.Lid = q.rid2 + 1; o.Ltype = (u32)(q.type >> 1); o.Llen = (u32)(q.len2 - q.jj - h); hasL = true; bestLj = q.jj; }
        }
        o.connections += (u32)hits.size();
    }
    (void)hasL;
    return true;
}

// Phase A of the reads ids[0..n) (or first + [0..n)); redo[s] is set, and the read skipped, when a lookup fails its proof
void scan_reads(Emu &e, const u32 *ids, u64 first, u64 n, const Lookup &lookup, uint8_t *redo)
{
    const int SW = e.SW, k = e.k, h = e.h;
    // K4 phase A: the kernel's round structure (search.cu): the (window, bucket entry) items of a read are
    // consumed 32 at a time; a round is evaluated in FAST mode (every hit checked against the previous hit
    // of its side only) until the first anomaly, then converted to the reference's sequential chain.
    // With SAGE2_EMUL_EXACT_ONLY set the plain sequential chain runs instead (cross-check of the hybrid).
    auto &flag5 = e.flag5;
    auto &cont_max = e.cont_max;
    std::vector<u64> prevR(SW), prevL(SW);
    const bool exact_only = getenv("SAGE2_EMUL_EXACT_ONLY") != nullptr;
    const bool fast_first = !exact_only && getenv("SAGE2_EMUL_NO_FAST") == nullptr;
    struct Item { int jj; u32 ent; };
    struct Hit { int jj; bool right; u32 rid2; int type; int len2; const u64 *Q; };
    std::vector<Item> items;
    std::vector<Hit> hits;
    for (u64 sb = 0; sb < n; ++sb) {
        const u64 i = ids ? (u64)ids[sb] : first + sb;
        const u64 *Xf = &e.F[i * SW], *Xr = &e.RC[i * SW];
        const int len1 = e.len[i];
        items.clear();
        bool collision = false;
        for (int j = 0; j <= len1 - h && !collision; ++j) {
            u64 v0, v1;
            extract_key(Xf, SW, j, h, v0, v1);
            const u32 *ents = nullptr;
            const int cnt = lookup(sb, Xf, j, v0, v1, ents);
            if (cnt < 0) { collision = true; break; }
            for (int x = 0; x < cnt; ++x) items.push_back(Item{ j, ents[x] });
        }
        if (collision) { redo[sb] = 1; e.restarts++; continue; }
        if (fast_first) {       // phase_a_fast_kernel first; the general scan below only sees what it could not certify
            FastOut fo;
            if (fast_certify(e, i, items, fo, 192, 128)) {
                for (auto &cm : fo.cont) cont_max[cm.first] = std::max(cont_max[cm.first], cm.second);
                e.compare_calls += fo.calls;
                e.fast_reads++;
                flag5[i] = fo.connections > kConnectionsLimit;
                e.extR[i] = ext_pack(fo.Rid, fo.Rtype, fo.Rlen);
                e.extL[i] = ext_pack(fo.Lid, fo.Ltype, fo.Llen);
                continue;
            }
        }
        ExtState st;
        ext_init(st);
        bool exact = exact_only, hasR = false, hasL = false;
        u32 Rid = 0, Rtype = 0, Rlen = 0, Lid = 0, Ltype = 0, Llen = 0, connections = 0;
        int cJR = 0, cLenR = 0, firstJR = 0, cJL = 0, cLenL = 0, curWin = -1;
        for (size_t r0 = 0; r0 < items.size(); r0 += 32) {
            hits.clear();
            for (size_t x = r0; x < std::min(items.size(), r0 + 32); ++x) {
                const int j = items[x].jj;
                const u32 rid2 = items[x].ent >> 2;
                const int type = (int)(items[x].ent & 3);
                const bool right = !(type & 1);
                if (rid2 == (u32)i || !(right ? gate_right(j, len1, k) : gate_left(j, k, h))) continue;
                const u64 *Q = (partner_uses_rc(type) ? &e.RC[0] : &e.F[0]) + (u64)rid2 * SW;
                const int len2 = e.len[rid2];
                bool cont;
                e.compare_calls++;
                const bool ok = overlap_equal(right ? Xf : Xr, len1, right ? j : len1 - j - h, Q, len2, SW, cont);
                if (ok && cont) cont_max[rid2] = std::max(cont_max[rid2], (u32)(i + 1));
                if (ok && !cont) hits.push_back(Hit{ j, right, rid2, type, len2, Q });
            }
            if (hits.empty()) continue;
            if (!exact) {
                bool anomaly = false;
                for (size_t x = 0; x < hits.size() && !anomaly; ++x) {
                    const Hit &hx = hits[x];
                    const u64 *prec = nullptr;
                    int prevJ = 0, prevLen = 0;
                    for (size_t y = x; y-- > 0;)
                        if (hits[y].right == hx.right) { prec = hits[y].Q; prevJ = hits[y].jj; prevLen = hits[y].len2; break; }
                    if (!prec && (hx.right ? hasR : hasL)) {
                        prec = hx.right ? prevR.data() : prevL.data();
                        prevJ = hx.right ? cJR : cJL; prevLen = hx.right ? cLenR : cLenL;
                    }
                    if (!prec) continue;
                    bool c2;
                    if (prevJ == hx.jj) anomaly = true;
                    else if (hx.right) anomaly = !overlap_equal(prec, prevLen, hx.jj - prevJ, hx.Q, hx.len2, SW, c2);
                    else anomaly = !overlap_equal(hx.Q, hx.len2, hx.jj - prevJ, prec, prevLen, SW, c2);
                }
                if (!anomaly) {
                    for (const Hit &hx : hits) {
                        if (hx.right) {
                            if (!hasR) { Rid = hx.rid2 + 1; Rtype = (u32)(hx.type >> 1); Rlen = (u32)(hx.len2 - (len1 - hx.jj)); firstJR = hx.jj; hasR = true; }
                            cJR = hx.jj; cLenR = hx.len2; copy_rec(prevR.data(), hx.Q, SW);
                        } else {
                            Lid = hx.rid2 + 1; Ltype = (u32)(hx.type >> 1); Llen = (u32)(hx.len2 - hx.jj - h); hasL = true;
                            cJL = hx.jj; cLenL = hx.len2; copy_rec(prevL.data(), hx.Q, SW);
                        }
                    }
                    connections += (u32)hits.size();
                    continue;
                }
                exact = true;
                e.slow_reads++;
                curWin = hits[0].jj;
                ext_init(st);
                st.Rid = Rid; st.Rtype = Rtype; st.Rlen = Rlen; st.Lid = Lid; st.Ltype = Ltype; st.Llen = Llen;
                st.prevJR = cJR; st.prevLenR = cLenR; st.prevPL = len1 - cJL - h; st.prevLenL = cLenL;
                st.markAmbigR = (hasR && cJR == curWin) ? 1 : 0;
                st.markFirstR = (hasR && firstJR == curWin) ? 1 : 0;
                st.markAmbigL = (hasL && cJL == curWin) ? 1 : 0;
                st.connections = connections;
            }
            for (const Hit &hx : hits) {
                if (hx.jj != curWin) { curWin = hx.jj; ext_new_window(st); }
                if (hx.right) ext_right_hit(st, prevR.data(), hx.Q, SW, hx.rid2 + 1, hx.type >> 1, hx.jj, len1, hx.len2);
                else ext_left_hit(st, prevL.data(), hx.Q, SW, hx.rid2 + 1, hx.type >> 1, hx.jj, h, len1, hx.len2);
            }
        }
        if (exact) {
            flag5[i] = st.connections > kConnectionsLimit;
            const bool amb = st.itsAmbigR == 1 || st.itsAmbigL == 1;
            e.extR[i] = ext_pack(st.Rid, st.Rtype, amb ? 0u : st.Rlen);
            e.extL[i] = ext_pack(st.Lid, st.Ltype, amb ? 0u : st.Llen);
        } else {
            flag5[i] = connections > kConnectionsLimit;
            e.extR[i] = ext_pack(Rid, Rtype, Rlen);
            e.extL[i] = ext_pack(Lid, Ltype, Llen);
        }
    }
}

// ---- phase_c_device.cu restated: lists, marks and filtering of phase C from the exploration order ----------------------
static u32 rev_t(u32 t) { return t == 0 ? 3u : (t == 3 ? 0u : t); }
static bool tr_rule(u32 t1, u32 t2) { return ((t1 == 0 || t1 == 2) && (t2 == 0 || t2 == 1)) || ((t1 == 1 || t1 == 3) && (t2 == 2 || t2 == 3)); }
void phase_c_from_order(const Emu &e, const PhaseCInput &in, const std::vector<u32> &order, std::vector<u64> &out, u64 &inserted, u64 &removed)
{
    std::map<u32, u32> sidx;                                     // read id (1-based) -> S index
    for (u64 s = 0; s < in.nS; ++s) sidx[in.s_ids[s] + 1] = (u32)s;
    std::map<u32, std::vector<u64>> lists;                       // owner id -> half-edge records id << 32 | type << 20 | overhang
    for (u64 x = 0; x < in.nB; ++x) {
        const u64 w0 = in.edgesB[2 * x], w1 = in.edgesB[2 * x + 1];
        const u32 a = (u32)(w0 >> 32), b = (u32)w0, type = (u32)(w1 >> 20) & 3u, len = (u32)(w1 & 0xFFFFFu);
        const u32 la = in.edgesB_len[x] & 0xFFFFu, lb = in.edgesB_len[x] >> 16;
        lists[a].push_back(((u64)b << 32) | ((u64)type << 20) | len);
        lists[b].push_back(((u64)a << 32) | ((u64)rev_t(type) << 20) | ((la - (lb - len)) & 0xFFFFFu));
    }
    inserted = 0; removed = 0;
    for (u64 s = 0; s < in.nS; ++s) {
        const u32 a = in.s_ids[s] + 1, la = e.len[a - 1];
        for (u32 q = in.cand_off[s]; q < in.cand_off[s + 1]; ++q) {
            const u64 cw = in.cand[q];
            const u32 b = (u32)(cw >> 32), t = (u32)(cw >> 20) & 3u;
            u32 d = (u32)(cw & 0xFFFFFu);
            if (d & 0x80000u) d |= 0xFFF00000u;
            if (!(order[s] < order[sidx.at(b)])) continue;
            const u32 lb = e.len[b - 1];
            lists[a].push_back(cw);
            lists[b].push_back(((u64)a << 32) | ((u64)rev_t(t) << 20) | ((la - (lb - d)) & 0xFFFFFu));
            inserted += 2;
        }
    }
    auto key_of = [](u64 cw) { return ((cw & 0xFFFFFull) << 34) | ((cw >> 32) << 2) | ((cw >> 20) & 3ull); };
    for (u64 s = 0; s < in.nS; ++s) {
        const u32 n = in.s_ids[s] + 1;
        auto it = lists.find(n);
        if (it == lists.end()) continue;
        std::vector<u64> key;
        for (u64 cw : it->second) key.push_back(key_of(cw));
        std::sort(key.begin(), key.end(), std::greater<u64>());               // length desc, id desc, type desc (:853-871)
        std::map<u32, int> st;
        for (u64 kx : key) st[(u32)(kx >> 2)] = 1;
        for (u64 kx : key) {
            const u32 ida = (u32)(kx >> 2), t1 = (u32)kx & 3u;
            if (st[ida] != 1) continue;
            auto jt = lists.find(ida);
            if (jt == lists.end()) continue;
            for (u64 cw : jt->second) {
                auto f = st.find((u32)(cw >> 32));
                if (f != st.end() && f->second == 1 && tr_rule(t1, (u32)(cw >> 20) & 3u)) f->second = 2;
            }
        }
        for (u64 kx : key) {
            const u32 id = (u32)(kx >> 2);
            if (st[id] == 2) { removed++; continue; }
            if (id > n) { out.push_back(((u64)n << 32) | id); out.push_back(((kx & 3ull) << 20) | (kx >> 34)); }
        }
    }
}

// rank's slice of the reads (sage2gpu_phase_a_partition); arrays padded to world * chunk
void phase_a(Emu &e, int rank, int world)
{
    alloc_phase_a(e, rank, world);
    std::vector<uint8_t> redo(e.pa_hi - e.pa_lo + 1, 0);
    scan_reads(e, nullptr, e.pa_lo, e.pa_hi - e.pa_lo, local_table(e), redo.data());
}

// ---- the sharded table: csrc/shard.cu on the CPU ---------------------------------------------------------------
void shard_build(Emu &e, int rank, int world)
{
    e.tb_rank = rank; e.tb_world = world;
    e.by_hash.clear();
    for (auto it = e.table.begin(); it != e.table.end();) {
        const u64 hsh = hash_key(it->first.first, it->first.second);
        if (key_owner(hsh, world) != rank) it = e.table.erase(it);
        else { e.by_hash[hsh] = it->first; ++it; }
    }
    e.distinct = e.table.size(); e.over = 0;
    for (auto &kv : e.table) if (kv.second.size() >= (size_t)kHashThreshold) e.over++;
}

void route_begin(Emu &e, int what, u64 first, u64 count, int exact, int world)
{
    e.rt_what = what; e.rt_exact = exact != 0; e.rt_world = world; e.rt_is_list = what != 0; e.rt_for_c = false;
    e.rt_ids.clear();
    if (what == 0) { e.rt_first = first; e.rt_n = count; }
    else if (what == 1) { for (u64 i = 0; i < e.U; ++i) if (e.explored_b[i] == 0) e.rt_ids.push_back((u32)i); e.rt_first = 0; e.rt_n = e.rt_ids.size(); }
    else { for (u64 i = 0; i < e.pa_hi - e.pa_lo; ++i) if (e.rt_redo[i]) e.rt_ids.push_back((u32)(e.pa_lo + i)); e.rt_first = 0; e.rt_n = e.rt_ids.size(); }
    e.rt_wstride = std::max(1, e.max_len - e.h + 1);
    const int qw = exact ? 2 : 1;
    std::vector<std::vector<u64>> q(world);
    std::vector<std::vector<u32>> m(world);
    for (u64 s = 0; s < e.rt_n; ++s) {
        const u64 i = e.rt_is_list ? (u64)e.rt_ids[s] : e.rt_first + s;
        const u64 *Xf = &e.F[i * e.SW];
        for (int j = 0; j <= (int)e.len[i] - e.h; ++j) {
            u64 v0, v1;
            extract_key(Xf, e.SW, j, e.h, v0, v1);
            const u64 hsh = hash_key(v0, v1);
            const int g = key_owner(hsh, world);
            if (exact) { q[g].push_back(v0); q[g].push_back(v1); } else q[g].push_back(hsh);
            m[g].push_back((u32)(s * (u64)e.rt_wstride + (u64)j));
        }
    }
    e.rt_queries.clear(); e.rt_qmap.clear();
    for (int g = 0; g < world; ++g) {
        e.rt_counts[g] = m[g].size();
        e.rt_queries.insert(e.rt_queries.end(), q[g].begin(), q[g].end());
        e.rt_qmap.insert(e.rt_qmap.end(), m[g].begin(), m[g].end());
    }
    (void)qw;
    e.rt_wslot.assign(e.rt_n * (u64)e.rt_wstride, 0);
    e.rt_state = 1;
}

void shard_answer(Emu &e, const u64 *queries, const u64 *counts_per_source, int exact, int world, u64 *entry_counts)
{
    const char *fe = getenv("SAGE2_EMUL_FAKE_TAG_COLLISIONS");
    const u64 fake_mask = fe ? strtoull(fe, nullptr, 0) : 0ull;
    e.an_resp.clear(); e.an_entries.clear();
    u64 p = 0;
    for (int s = 0; s < world; ++s) {
        const u64 ebase = e.an_entries.size();
        for (u64 x = 0; x < counts_per_source[s]; ++x, ++p) {
            const std::vector<u32> *bucket = nullptr;
            if (exact) {
                auto it = e.table.find(std::make_pair(queries[2 * p], queries[2 * p + 1]));
                if (it != e.table.end() && it->second.size() < (size_t)kHashThreshold) bucket = &it->second;
            } else {
                const u64 hsh = queries[p];
                if (fake_mask && (hsh & fake_mask) == 0 && !e.table.empty()) bucket = &e.table.begin()->second;   // a wrong bucket, like a tag collision
                else {
                    auto it = e.by_hash.find(hsh);
                    if (it != e.by_hash.end()) bucket = &e.table[it->second];
                }
            }
            u64 answer = 0;
            if (bucket) {
                const size_t c = bucket->size();
                if (c == 1 || c >= (size_t)kHashThreshold) answer = answer_encode((u32)c, (*bucket)[0]);
                else {
                    answer = answer_encode((u32)c, e.an_entries.size() - ebase);
                    e.an_entries.insert(e.an_entries.end(), bucket->begin(), bucket->end());
                }
            }
            e.an_resp.push_back(answer);
        }
        entry_counts[s] = e.an_entries.size() - ebase;
    }
}

void route_finish(Emu &e, const u64 *resp, const u32 *entries, const u64 *entry_counts)
{
    u64 E = 0, p = 0;
    for (int g = 0; g < e.rt_world; ++g) {
        for (u64 x = 0; x < e.rt_counts[g]; ++x, ++p) {
            u64 w = resp[p];
            const u32 c = slot_get_count(w);
            if (c >= 2 && c < (u32)kHashThreshold) w += E;
            e.rt_wslot[e.rt_qmap[p]] = w;
        }
        E += entry_counts[g];
    }
    e.rt_wentries.assign(entries, entries + E);
    e.rt_state = 2;
    e.rt_for_c = e.rt_what == 1;
}

// the routed batch as a lookup; untrusted answers are proven like search.cu does (first entry of the bucket /
// representative of a masked key against the window's key)
Lookup routed_table(Emu &e)
{
    return [&e](u64 s, const u64 *, int j, u64 v0, u64 v1, const u32 *&ents) -> int {
        const u64 w = e.rt_wslot[s * (u64)e.rt_wstride + (u64)j];
        const u32 c = slot_get_count(w);
        if (c == 0) return 0;
        static thread_local u32 single;
        const u32 *first_entry;
        if (c == 1 || c >= (u32)kHashThreshold) { single = (u32)slot_get_payload(w); first_entry = &single; }
        else first_entry = &e.rt_wentries[slot_get_payload(w)];
        if (!e.rt_exact) {
            const u64 r = *first_entry >> 2;
            u64 w0, w1;
            entry_key(&e.F[r * e.SW], &e.RC[r * e.SW], e.SW, e.len[r], e.h, (int)(*first_entry & 3), w0, w1);
            if (w0 != v0 || w1 != v1) return -1;
        }
        if (c >= (u32)kHashThreshold) return 0;
        ents = first_entry;
        return (int)c;
    };
}

u64 phase_a_routed(Emu &e)
{
    const u64 before = e.restarts;
    std::vector<uint8_t> redo2(e.rt_n + 1, 0);
    uint8_t *redo = e.rt_is_list ? redo2.data() : e.rt_redo.data() + (e.rt_first - e.pa_lo);
    scan_reads(e, e.rt_is_list ? e.rt_ids.data() : nullptr, e.rt_first, e.rt_n, routed_table(e), redo);
    e.rt_state = 0;
    return e.restarts - before;
}

void phase_b(Emu &e);
void finish_after_b(Emu &e);
void finish(Emu &e)
{
    if (!e.have_phase_b) phase_b(e);
    finish_after_b(e);
    e.have_phase_b = false;
}

void phase_b(Emu &e)
{
    const u64 U = e.U;
    auto &flag5 = e.flag5;
    auto &cont_max = e.cont_max;
    e.have_phase_b = true;
    e.contained = e.contained_size = 0;
    // phase B
    e.explored_a.resize(U); e.explored_b.resize(U);
    for (u64 i = 0; i < U; ++i) {
        uint8_t s = state_after_a((u32)i + 1, cont_max[i], flag5[i]);
        e.explored_a[i] = s;
        if (s != 6) { if (phase_b_qualifies(e.extR.data(), e.extL.data(), i)) { s = 4; e.contained++; } }
        else e.contained_size++;
        e.explored_b[i] = s;
    }
}

void finish_after_b(Emu &e)
{
    const int SW = e.SW, k = e.k, h = e.h;
    const u64 U = e.U;
    const Lookup lookup = (e.tb_world > 1 || e.rt_for_c) ? routed_table(e) : local_table(e);
    std::vector<u64> edgesB;
    for (u64 i = 0; i < U; ++i) {
        EdgeRec r[2];
        const int nr = phase_b_edges(e.extR.data(), e.extL.data(), e.explored_b.data(), e.len.data(), i, r);
        for (int t = 0; t < nr; ++t) { edgesB.push_back(r[t].w0); edgesB.push_back(r[t].w1); }
    }
    // K5 candidates + host walk
    std::vector<u32> s_ids, cand_off;
    std::vector<u64> cand;
    for (u64 i = 0; i < U; ++i) {
        if (e.explored_b[i] != 0) continue;
        const u64 sb = s_ids.size();
        s_ids.push_back((u32)i);
        cand_off.push_back((u32)cand.size());
        const u64 *Xf = &e.F[i * SW], *Xr = &e.RC[i * SW];
        const int len1 = e.len[i];
        for (int j = 0; j <= len1 - h; ++j) {
            u64 v0, v1;
            extract_key(Xf, SW, j, h, v0, v1);
            const u32 *ents = nullptr;
            const int cnt = lookup(sb, Xf, j, v0, v1, ents);
            for (int x = 0; x < cnt; ++x) {
                const u32 ent = ents[x];
                const u32 rid2 = ent >> 2;
                const int type = (int)(ent & 3);
                const bool right = !(type & 1);
                if (rid2 == (u32)i || e.explored_b[rid2] != 0 || !(right ? gate_right(j, len1, k) : gate_left(j, k, h))) continue;
                const u64 *Q = (partner_uses_rc(type) ? &e.RC[0] : &e.F[0]) + (u64)rid2 * SW;
                bool cont;
                if (overlap_equal(right ? Xf : Xr, len1, right ? j : len1 - j - h, Q, e.len[rid2], SW, cont))
                    cand.push_back(candidate_record(type, j, h, len1, e.len[rid2], rid2));
            }
        }
    }
    cand_off.push_back((u32)cand.size());
    std::vector<u64> all;
    if (!s_ids.empty()) {
        // what graph.cu hands over: lengths of the S reads, and only the phase-B records that touch S or a
        // phase-B neighbour of S, each with the lengths of its endpoints
        std::vector<uint16_t> s_len;
        for (u32 sid : s_ids) s_len.push_back(e.len[sid]);
        std::vector<uint8_t> needed(U, 0);
        for (u32 sid : s_ids) needed[sid] = 1;
        for (size_t x = 0; x < edgesB.size(); x += 2) {
            const u32 a = (u32)(edgesB[x] >> 32) - 1, b = (u32)edgesB[x] - 1;
            if (e.explored_b[a] == 0) needed[b] = 1;
            if (e.explored_b[b] == 0) needed[a] = 1;
        }
        std::vector<u64> selB;
        std::vector<u32> selLen;
        for (size_t x = 0; x < edgesB.size(); x += 2) {
            const u32 a = (u32)(edgesB[x] >> 32) - 1, b = (u32)edgesB[x] - 1;
            if (!needed[a] && !needed[b]) continue;
            selB.push_back(edgesB[x]); selB.push_back(edgesB[x + 1]);
            selLen.push_back((u32)e.len[a] | ((u32)e.len[b] << 16));
        }
        PhaseCInput in;
        in.nS = s_ids.size(); in.s_ids = s_ids.data(); in.s_len = s_len.data(); in.cand_off = cand_off.data();
        in.cand = cand.data(); in.nB = selB.size() / 2; in.edgesB = selB.data(); in.edgesB_len = selLen.data();
        PhaseCOutput out;
        run_host_phase_c(in, out);
        e.inserted = out.inserted; e.removed = out.removed;
        all = out.edges;
        // phase_c_device.cu on the CPU: the same records from the traversal ORDER alone (run_host_phase_c_order): half
        // edges of every overlap from the end point explored first, lists sorted, marked and filtered read by read
        if (!getenv("SAGE2_EMUL_NO_ORDER_CHECK")) {
            std::vector<u32> order;
            // as graph.cu hands it over: S index of every candidate's read2, S reads that own phase-B records
            std::map<u32, u32> sidx;
            for (u64 sx = 0; sx < in.nS; ++sx) sidx[in.s_ids[sx] + 1] = (u32)sx;
            std::vector<u32> cnode(cand.size());
            for (size_t q = 0; q < cand.size(); ++q) cnode[q] = sidx.at((u32)(cand[q] >> 32));
            std::vector<uint8_t> hasb(in.nS, 0);
            for (u64 x = 0; x < in.nB; ++x) {
                const u32 a = (u32)(selB[2 * x] >> 32), b = (u32)selB[2 * x];
                if (sidx.count(a)) hasb[sidx[a]] = 1;
                if (sidx.count(b)) hasb[sidx[b]] = 1;
            }
            // the traversal as graph.cu runs it: from the lists phase_c_sorted_lists prepares (here: tests/phase_c_lists_ref.h)
            pc_ref::Lists lists = pc_ref::build(in, cnode);
            lists.has_b = hasb;
            run_host_phase_c_order_lists(lists.view(in.nS), order);
            if (!getenv("SAGE2_EMUL_NO_PLAIN_ORDER_CHECK")) {       // the walk itself (its own id map, sort and insertions) must explore in the same order
                std::vector<u32> order2;
                run_host_phase_c_order(in, order2);
                if (order2 != order) { fprintf(stderr, "host_emul: traversal order from the sorted lists differs from the walk's\n"); abort(); }
            }
            std::vector<u64> got;
            u64 ins2 = 0, rem2 = 0;
            phase_c_from_order(e, in, order, got, ins2, rem2);
            std::vector<std::pair<u64, u64>> A, B;
            for (size_t x = 0; x < all.size(); x += 2) A.emplace_back(all[x], all[x + 1]);
            for (size_t x = 0; x < got.size(); x += 2) B.emplace_back(got[x], got[x + 1]);
            std::sort(A.begin(), A.end()); std::sort(B.begin(), B.end());
            if (A != B || ins2 != out.inserted || rem2 != out.removed) {
                fprintf(stderr, "host_emul: phase C from the traversal order differs from the walk (%zu vs %zu records, inserted %llu vs %llu, removed %llu vs %llu)\n",
                        B.size(), A.size(), (unsigned long long)ins2, (unsigned long long)out.inserted, (unsigned long long)rem2, (unsigned long long)out.removed);
                abort();
            }
        }
    }
    for (size_t x = 0; x < edgesB.size(); x += 2)
        if (e.explored_b[(edgesB[x] >> 32) - 1] != 0) { all.push_back(edgesB[x]); all.push_back(edgesB[x + 1]); }
    // K6
    std::vector<std::pair<u64, u64>> recs2;
    for (size_t x = 0; x < all.size(); x += 2) recs2.emplace_back(all[x], all[x + 1]);
    std::sort(recs2.begin(), recs2.end());
    for (size_t x = 0; x < recs2.size(); ++x) {
        if (x > 0 && recs2[x].first == recs2[x - 1].first && (recs2[x].second >> 20) == (recs2[x - 1].second >> 20)) continue;
        e.edges.push_back(recs2[x].first); e.edges.push_back(recs2[x].second);
    }
}

}  // namespace

extern "C" {

void *hemu_run(const uint8_t *bases, const int64_t *off, int64_t n, int k)
{
    Emu *e = new Emu();
    prepare(*e, bases, off, n, k);
    phase_a(*e, 0, 1);
    finish(*e);
    return e;
}
// the multi-GPU split: prepare (steps 1-2), rank's phase-A slice, [exchange by the caller], finish
void *hemu_prepare(const uint8_t *bases, const int64_t *off, int64_t n, int k)
{
    Emu *e = new Emu();
    prepare(*e, bases, off, n, k);
    return e;
}
// every stage partitioned (include/sage2gpu.h): this rank's key range of the reads ...
void *hemu_prepare_partition(const uint8_t *bases, const int64_t *off, int64_t n, int k, int rank, int world, uint64_t *unique_local)
{
    Emu *e = new Emu();
    prepare(*e, bases, off, n, k, rank, world);
    *unique_local = e->U;
    return e;
}
// ... arrays of the total size with this rank's run in place (the caller all-gathers them) ...
void hemu_reads_gather_layout(void *p, const uint64_t *counts, int rank, int world, uint64_t **F, uint16_t **len, uint16_t **freq)
{
    Emu *e = (Emu *)p;
    u64 tot = 0, base = 0;
    for (int q = 0; q < world; ++q) { if (q < rank) base += counts[q]; tot += counts[q]; }
    std::vector<u64> Fg(tot * e->SW, 0);
    std::vector<uint16_t> lg(tot, 0), fg(tot, 0);
    std::copy(e->F.begin(), e->F.end(), Fg.begin() + base * e->SW);
    std::copy(e->len.begin(), e->len.end(), lg.begin() + base);
    std::copy(e->freq.begin(), e->freq.end(), fg.begin() + base);
    e->F.swap(Fg); e->len.swap(lg); e->freq.swap(fg);
    *F = e->F.data(); *len = e->len.data(); *freq = e->freq.data();
}
// ... and the reads are complete: reverse complements, table
void hemu_reads_gather_finish(void *p) { finish_reads(*(Emu *)p); }
void hemu_phase_a(void *p, int rank, int world) { phase_a(*(Emu *)p, rank, world); }
// pointers to the padded phase-A arrays (u64, u64, u8, u32) and their length
uint64_t hemu_phase_a_arrays(void *p, uint64_t **extR, uint64_t **extL, uint8_t **flag5, uint32_t **cont_max)
{
    Emu *e = (Emu *)p;
    *extR = e->extR.data(); *extL = e->extL.data(); *flag5 = e->flag5.data(); *cont_max = e->cont_max.data();
    return e->extR.size();
}
void hemu_finish(void *p) { finish(*(Emu *)p); }
// the sharded table: the entry points of include/sage2gpu.h with host buffers
void hemu_shard_build(void *p, int rank, int world) { shard_build(*(Emu *)p, rank, world); }
void hemu_phase_a_sharded_begin(void *p, int rank, int world, uint64_t *first, uint64_t *count)
{
    Emu *e = (Emu *)p;
    alloc_phase_a(*e, rank, world);
    *first = e->pa_lo; *count = e->pa_hi - e->pa_lo;
}
void hemu_route_begin(void *p, int what, uint64_t first, uint64_t count, int exact, int world, uint64_t **queries, uint64_t *counts, uint64_t *n_reads)
{
    Emu *e = (Emu *)p;
    route_begin(*e, what, first, count, exact, world);
    *queries = e->rt_queries.data(); *n_reads = e->rt_n;
    for (int g = 0; g < world; ++g) counts[g] = e->rt_counts[g];
}
void hemu_shard_answer(void *p, const uint64_t *queries, const uint64_t *counts_per_source, int exact, int world, uint64_t **resp,
                       uint32_t **entries, uint64_t *entry_counts)
{
    Emu *e = (Emu *)p;
    shard_answer(*e, queries, counts_per_source, exact, world, entry_counts);
    *resp = e->an_resp.data(); *entries = e->an_entries.data();
}
void hemu_route_finish(void *p, const uint64_t *resp, const uint32_t *entries, const uint64_t *entry_counts) { route_finish(*(Emu *)p, resp, entries, entry_counts); }
uint64_t hemu_phase_a_routed(void *p) { return phase_a_routed(*(Emu *)p); }
void hemu_phase_b(void *p) { phase_b(*(Emu *)p); }
void hemu_free(void *p) { delete (Emu *)p; }
// sizes: [0]=U [1]=SW [2]=good [3]=total_bp [4]=n_edges [5]=over [6]=distinct [7]=compare_calls [8]=inserted
//        [9]=removed [10]=contained [11]=contained_size [12]=reads that left the fast mode
void hemu_sizes(void *p, uint64_t *o)
{
    Emu *e = (Emu *)p;
    o[0] = e->U; o[1] = (u64)e->SW; o[2] = e->good; o[3] = e->total_bp; o[4] = e->edges.size() / 2; o[5] = e->over;
    o[6] = e->distinct; o[7] = e->compare_calls; o[8] = e->inserted; o[9] = e->removed; o[10] = e->contained;
    o[11] = e->contained_size; o[12] = e->slow_reads; o[13] = e->fast_reads;
}
void hemu_copy(void *p, uint64_t *F, uint64_t *RC, uint16_t *len, uint16_t *freq, uint64_t *extR, uint64_t *extL,
               uint8_t *expl_a, uint8_t *expl_b, uint64_t *edges)
{
    Emu *e = (Emu *)p;
    if (F) memcpy(F, e->F.data(), e->F.size() * 8);
    if (RC) memcpy(RC, e->RC.data(), e->RC.size() * 8);
    if (len) memcpy(len, e->len.data(), e->len.size() * 2);
    if (freq) memcpy(freq, e->freq.data(), e->freq.size() * 2);
    if (extR) memcpy(extR, e->extR.data(), e->U * 8);
    if (extL) memcpy(extL, e->extL.data(), e->U * 8);
    if (expl_a) memcpy(expl_a, e->explored_a.data(), e->explored_a.size());
    if (expl_b) memcpy(expl_b, e->explored_b.data(), e->explored_b.size());
    if (edges) memcpy(edges, e->edges.data(), e->edges.size() * 8);
}

// unit-level entry points
uint64_t hemu_get_bases(const uint64_t *rec, int SW, int s, int nb) { return get_bases(rec, SW, s, nb); }
void hemu_pack(const uint8_t *s, int len, int SW, uint64_t *rec, int rc) { pack_record(s, len, SW, rec, rc != 0); }
void hemu_revcomp(const uint64_t *F, uint64_t *R, int SW, int len) { revcomp_record(F, R, SW, len); }
void hemu_key(const uint64_t *rec, int SW, int j, int h, uint64_t *v) { extract_key(rec, SW, j, h, v[0], v[1]); }
int hemu_overlap(const uint64_t *X, int lenX, int start, const uint64_t *Y, int lenY, int SW, int *contained)
{
    bool c;
    const bool ok = overlap_equal(X, lenX, start, Y, lenY, SW, c);
    *contained = c;
    return ok;
}

}  // extern "C"
