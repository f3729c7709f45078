// host_phase_c.cpp -- EconomyGraph::buildOverlapGraphEconomy (economyGraph/economyGraph.cpp:495-574)
// with insertAllEdgesOfRead's search replaced by the GPU's candidate lists.
//
// Why this is on the host: the walk is a FIFO breadth-first traversal whose edge insertions depend
// on which reads were explored before (:605), followed by Myers marking whose outcome depends on list
// order (:643-679).  SURVEY.md section 8(f) N1 lists moving it to the device as the next step.
//
// Only the nodes the walk can touch get adjacency lists: S (state 0 after phase B) and the phase-B
// neighbours of S (whose lists markTransitiveEdge scans, :653).  Everything else keeps the device's
// phase-B records untouched.  What leaves this function: for every a in S, the entries of list[a]
// with id > a after the walk, i.e. exactly what convertGraph would consume (overlapGraph.cpp:103).
#include "host_phase_c.h"

#include <algorithm>
#include <chrono>
#include <thread>

namespace sg {
namespace {

struct HEdge { uint32_t id; uint32_t node; uint8_t type, mark; uint32_t length; };     // node = local index of `id`, kNone if not in play
constexpr uint32_t kNone = 0xFFFFFFFFu;

inline uint32_t rev_type(uint32_t t) { return t == 0 ? 3u : (t == 3 ? 0u : t); }

// compareLengthBased, economyGraph.cpp:853-871
inline bool by_length_desc(const HEdge &a, const HEdge &b)
{
    if (a.length != b.length) return a.length > b.length;
    if (a.id != b.id) return a.id > b.id;
    return a.type > b.type;
}

// read id -> local node index: open addressing on a power-of-two table (the walk does millions of lookups)
struct IdMap {
    std::vector<uint32_t> key, val;
    uint32_t mask = 0;
    void init(size_t expected)
    {
        size_t cap = 64;
        while (cap < 2 * expected + 16) cap <<= 1;
        key.assign(cap, 0); val.assign(cap, 0);
        mask = (uint32_t)(cap - 1);
    }
    static uint32_t h(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
    uint32_t find(uint32_t id) const      // ids are >= 1; 0 marks an empty slot
    {
        for (uint32_t s = h(id) & mask;; s = (s + 1) & mask) {
            if (key[s] == id) return val[s];
            if (key[s] == 0) return kNone;
        }
    }
    void put(uint32_t id, uint32_t v)
    {
        for (uint32_t s = h(id) & mask;; s = (s + 1) & mask)
            if (key[s] == 0) { key[s] = id; val[s] = v; return; }
    }
};

struct Walk {
    const PhaseCInput &in;
    IdMap slot;                                    // read id (1-based) -> local node
    std::vector<std::vector<HEdge>> adj;
    std::vector<uint8_t> state;                    // 0/1/2 for S nodes, 4 for everything else
    std::vector<uint8_t> marked;
    std::vector<uint32_t> node_id, node_len;
    uint64_t inserted = 0, removed = 0;

    explicit Walk(const PhaseCInput &i) : in(i) {}

    uint32_t node(uint32_t id, uint8_t st, uint32_t len)
    {
        const uint32_t f = slot.find(id);
        if (f != kNone) return f;
        const uint32_t n = (uint32_t)adj.size();
        slot.put(id, n);
        adj.emplace_back();
        state.push_back(st);
        marked.push_back(0);
        node_id.push_back(id);
        node_len.push_back(len);
        return n;
    }

    // insertEdgeEconomy, economyGraph.cpp:813-849 (both endpoints are S nodes here)
    void insert_edge(uint32_t nu, uint32_t nv, uint32_t delta, uint32_t type)
    {
        const uint32_t u = node_id[nu], v = node_id[nv];
        const uint32_t lu = node_len[nu], lv = node_len[nv];
        const uint32_t delta2 = lu - (lv - delta);
        adj[nu].push_back(HEdge{ v, nv, (uint8_t)type, 0, delta & 0xFFFFFu });
        adj[nv].push_back(HEdge{ u, nu, (uint8_t)rev_type(type), 0, delta2 & 0xFFFFFu });
    }

    // insertAllEdgesOfRead, economyGraph.cpp:580-638
    void insert_all(uint32_t n1)
    {
        if (state[n1] != 0) return;
        state[n1] = 1;
        const uint32_t s = n1;    // S nodes occupy local slots 0..nS-1 in s_ids order
        uint64_t cnt = 0;
        for (uint32_t q = in.cand_off[s]; q < in.cand_off[s + 1]; ++q) {
            const uint64_t cw = in.cand[q];
            const uint32_t n2 = slot.find((uint32_t)(cw >> 32));
            if (state[n2] != 0) continue;                                   // :605
            uint32_t delta = (uint32_t)(cw & 0xFFFFFu);
            if (delta & 0x80000u) delta |= 0xFFF00000u;                     // sign-extend the int32 overhang
            insert_edge(n1, n2, delta, (uint32_t)((cw >> 20) & 3u));
            cnt++;
        }
        if (adj[n1].size() > 1) std::sort(adj[n1].begin(), adj[n1].end(), by_length_desc);   // :634
        inserted += 2 * cnt;
    }

    // markTransitiveEdge, economyGraph.cpp:643-679.  The walk only records that nf's marking is due (its position in the
    // traversal is what later pushes depend on); the marking itself is computed after the walk, see run_host_phase_c.
    void mark_transitive(uint32_t nf) { state[nf] = 2; }

    void compute_marks(uint32_t nf, std::vector<uint8_t> &scratch)      // scratch: one zeroed byte per node
    {
        std::vector<HEdge> &lf = adj[nf];
        for (auto &e : lf) scratch[e.node] = 1;
        for (auto &e : lf) {
            const uint32_t na = e.node;
            if (scratch[na] != 1) continue;
            for (auto &f : adj[na]) {
                const uint32_t nb = f.node;
                if (nb == kNone) continue;            // not adjacent to any S read: cannot be in play
                if (scratch[nb] != 1) continue;
                const uint32_t t1 = e.type, t2 = f.type;
                if ((t1 == 0 || t1 == 2) && (t2 == 0 || t2 == 1)) scratch[nb] = 2;
                else if ((t1 == 1 || t1 == 3) && (t2 == 2 || t2 == 3)) scratch[nb] = 2;
            }
        }
        for (auto &e : lf) if (scratch[e.node] == 2) e.mark = 1;
        for (auto &e : lf) scratch[e.node] = 0;
        scratch[nf] = 0;
    }

    // removeTransitiveEdges, economyGraph.cpp:681-707 (after all marks are known)
    void remove_marked(uint32_t n)
    {
        std::vector<HEdge> &l = adj[n];
        size_t w = 0;
        for (size_t r = 0; r < l.size(); ++r) if (!l[r].mark) l[w++] = l[r];
        removed += l.size() - w;
        l.resize(w);
    }
};

}  // namespace

float run_host_phase_c(const PhaseCInput &in, PhaseCOutput &out)
{
    const auto t0 = std::chrono::steady_clock::now();
    Walk w(in);
    w.slot.init(in.nS + 2 * in.nB);
    for (uint64_t s = 0; s < in.nS; ++s) w.node(in.s_ids[s] + 1, 0, in.s_len[s]);
    const uint32_t nS = (uint32_t)in.nS;

    // phase-B neighbours of S, then every phase-B entry of every needed node
    for (uint64_t e = 0; e < in.nB; ++e) {
        const uint32_t a = (uint32_t)(in.edgesB[2 * e] >> 32), b = (uint32_t)in.edgesB[2 * e];
        const uint32_t ia = w.slot.find(a), ib = w.slot.find(b);
        const bool a_in_s = ia != kNone && ia < nS;
        const bool b_in_s = ib != kNone && ib < nS;
        if (a_in_s && ib == kNone) w.node(b, 4, in.edgesB_len[e] >> 16);
        if (b_in_s && ia == kNone) w.node(a, 4, in.edgesB_len[e] & 0xFFFFu);
    }
    for (uint64_t e = 0; e < in.nB; ++e) {
        const uint64_t w0 = in.edgesB[2 * e], w1 = in.edgesB[2 * e + 1];
        const uint32_t a = (uint32_t)(w0 >> 32), b = (uint32_t)w0;
        const uint32_t type = (uint32_t)(w1 >> 20) & 3u, length = (uint32_t)(w1 & 0xFFFFFu);
        const uint32_t ia = w.slot.find(a), ib = w.slot.find(b);
        if (ia != kNone) w.adj[ia].push_back(HEdge{ b, ib, (uint8_t)type, 0, length });
        if (ib != kNone) {
            const uint32_t la = in.edgesB_len[e] & 0xFFFFu, lb = in.edgesB_len[e] >> 16;
            w.adj[ib].push_back(HEdge{ a, ia, (uint8_t)rev_type(type), 0, (la - (lb - length)) & 0xFFFFFu });
        }
    }

    // buildOverlapGraphEconomy, economyGraph.cpp:513-564
    std::vector<uint32_t> queue;
    queue.reserve(nS);
    for (uint32_t i = 0; i < nS; ++i) {
        if (w.state[i] != 0) continue;
        queue.clear();
        size_t start = 0;
        queue.push_back(i);
        while (start < queue.size()) {
            const uint32_t n1 = queue[start++];
            if (w.state[n1] == 0) w.insert_all(n1);
            if (w.adj[n1].empty()) continue;                                 // :525
            if (w.state[n1] == 1) {
                for (size_t x = 0; x < w.adj[n1].size(); ++x) {
                    const uint32_t n2 = w.adj[n1][x].node;
                    if (w.state[n2] == 0) { queue.push_back(n2); w.insert_all(n2); }
                }
                w.mark_transitive(n1);
            }
            if (w.state[n1] == 2) {
                for (size_t x = 0; x < w.adj[n1].size(); ++x) {
                    const uint32_t n2 = w.adj[n1][x].node;
                    if (w.state[n2] != 1) continue;
                    for (size_t y = 0; y < w.adj[n2].size(); ++y) {
                        const uint32_t n3 = w.adj[n2][y].node;
                        if (w.state[n3] == 0) { queue.push_back(n3); w.insert_all(n3); }
                    }
                    w.mark_transitive(n2);
                }
            }
        }
    }

    // Marking and removal, after the walk.  When the reference marks a node, every neighbour already has all its
    // edges (insertAllEdgesOfRead ran for it and nothing is appended to an explored node's list), the node's own list
    // was sorted at the end of its insertAllEdgesOfRead, and a node's marked edges are removed only after all its
    // neighbours were marked: so the marks are a function of the complete lists alone and every node can be marked
    // independently -- here on all host cores (the quadratic part of the phase: sum over nodes of degree^2).
    {
        std::vector<uint32_t> todo;
        for (uint32_t i = 0; i < nS; ++i) if (w.state[i] == 2) todo.push_back(i);
        unsigned T = std::thread::hardware_concurrency();
        if (T > 32) T = 32;
        if (T < 1 || todo.size() < 4096) T = 1;
        const size_t n_nodes = w.adj.size();
        auto work = [&](unsigned t) {
            std::vector<uint8_t> scratch(n_nodes, 0);
            for (size_t x = t; x < todo.size(); x += T) w.compute_marks(todo[x], scratch);
        };
        if (T == 1) work(0);
        else {
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < T; ++t) pool.emplace_back(work, t);
            for (auto &th : pool) th.join();
        }
        for (uint32_t i : todo) w.remove_marked(i);
    }

    // what convertGraph consumes from the lists of S reads (overlapGraph.cpp:93-112)
    out.edges.clear();
    for (uint32_t i = 0; i < nS; ++i) {
        const uint32_t a = w.node_id[i];
        for (const HEdge &e : w.adj[i])
            if (e.id > a) {
                out.edges.push_back(((uint64_t)a << 32) | e.id);
                out.edges.push_back(((uint64_t)e.type << 20) | e.length);
            }
    }
    out.inserted = w.inserted;
    out.removed = w.removed;
    return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace sg
