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

#include <stdio.h>
#include <stdlib.h>
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

// Adjacency lists live in ONE pool: list n is pool[start[n] .. start[n] + size[n]) with room for every entry it can ever
// receive (its own candidates, the candidates of others that name it, its phase-B records), so the walk never allocates.
// The big buffers survive between calls (grow-only, one set per calling thread): a fresh 50 MB block per call costs more in
// page faults than the walk itself.
struct Workspace {
    HEdge *pool = nullptr;
    size_t pool_cap = 0;
    std::vector<uint32_t> cnode, room, eb_a, eb_b;
    HEdge *get_pool(size_t n)
    {
        if (n > pool_cap) {
            free(pool);
            pool_cap = n + n / 4 + 1024;
            pool = (HEdge *)malloc(pool_cap * sizeof(HEdge));       // entries are always written before they are read
        }
        return pool;
    }
    ~Workspace() { free(pool); }
};

struct Walk {
    const PhaseCInput &in;
    IdMap slot;                                    // read id (1-based) -> local node
    HEdge *pool = nullptr;
    std::vector<uint64_t> start;
    std::vector<uint32_t> size;
    const uint32_t *cnode = nullptr;               // candidate q -> local node of its read2 (resolved by all cores before the walk)
    std::vector<uint8_t> state;                    // 0/1/2 for S nodes, 4 for everything else
    std::vector<uint32_t> node_id, node_len;
    uint64_t inserted = 0, removed = 0;
    std::vector<uint32_t> *order = nullptr;        // when set: order[s] = position of S read s in the exploration sequence

    explicit Walk(const PhaseCInput &i) : in(i) {}

    uint32_t node(uint32_t id, uint8_t st, uint32_t len)
    {
        const uint32_t f = slot.find(id);
        if (f != kNone) return f;
        const uint32_t n = (uint32_t)state.size();
        slot.put(id, n);
        state.push_back(st);
        node_id.push_back(id);
        node_len.push_back(len);
        return n;
    }
    HEdge *list(uint32_t n) { return pool + start[n]; }
    void push(uint32_t n, const HEdge &e) { pool[start[n] + size[n]++] = e; }

    // insertEdgeEconomy, economyGraph.cpp:813-849 (both endpoints are S nodes here)
    void insert_edge(uint32_t nu, uint32_t nv, uint32_t delta, uint32_t type)
    {
        const uint32_t u = node_id[nu], v = node_id[nv];
        const uint32_t lu = node_len[nu], lv = node_len[nv];
        const uint32_t delta2 = lu - (lv - delta);
        push(nu, HEdge{ v, nv, (uint8_t)type, 0, delta & 0xFFFFFu });
        push(nv, HEdge{ u, nu, (uint8_t)rev_type(type), 0, delta2 & 0xFFFFFu });
    }

    // insertAllEdgesOfRead, economyGraph.cpp:580-638
    void insert_all(uint32_t n1, uint32_t &counter)
    {
        if (state[n1] != 0) return;
        state[n1] = 1;
        if (order) (*order)[n1] = counter++;
        const uint32_t s = n1;    // S nodes occupy local slots 0..nS-1 in s_ids order
        uint64_t cnt = 0;
        for (uint32_t q = in.cand_off[s]; q < in.cand_off[s + 1]; ++q) {
            const uint32_t n2 = cnode[q];
            if (n2 == kNone || state[n2] != 0) continue;                    // :605
            const uint64_t cw = in.cand[q];
            uint32_t delta = (uint32_t)(cw & 0xFFFFFu);
            if (delta & 0x80000u) delta |= 0xFFF00000u;                     // sign-extend the int32 overhang
            insert_edge(n1, n2, delta, (uint32_t)((cw >> 20) & 3u));
            cnt++;
        }
        if (size[n1] > 1) std::sort(list(n1), list(n1) + size[n1], by_length_desc);   // :634
        inserted += 2 * cnt;
    }

    // markTransitiveEdge, economyGraph.cpp:643-679.  The walk only records that nf's marking is due (its position in the
    // traversal is what later pushes depend on); the marking itself is computed after the walk, see run_host_phase_c.
    void mark_transitive(uint32_t nf) { state[nf] = 2; }

    void compute_marks(uint32_t nf, std::vector<uint8_t> &scratch)      // scratch: one zeroed byte per node
    {
        HEdge *lf = list(nf);
        const uint32_t nfe = size[nf];
        for (uint32_t x = 0; x < nfe; ++x) scratch[lf[x].node] = 1;
        for (uint32_t x = 0; x < nfe; ++x) {
            const HEdge &e = lf[x];
            const uint32_t na = e.node;
            if (scratch[na] != 1) continue;
            const HEdge *la = list(na);
            const uint32_t nae = size[na];
            for (uint32_t y = 0; y < nae; ++y) {
                const HEdge &f = la[y];
                const uint32_t nb = f.node;
                if (nb == kNone) continue;            // not adjacent to any S read: cannot be in play
                if (scratch[nb] != 1) continue;
                const uint32_t t1 = e.type, t2 = f.type;
                if ((t1 == 0 || t1 == 2) && (t2 == 0 || t2 == 1)) scratch[nb] = 2;
                else if ((t1 == 1 || t1 == 3) && (t2 == 2 || t2 == 3)) scratch[nb] = 2;
            }
        }
        for (uint32_t x = 0; x < nfe; ++x) if (scratch[lf[x].node] == 2) lf[x].mark = 1;
        for (uint32_t x = 0; x < nfe; ++x) scratch[lf[x].node] = 0;
        scratch[nf] = 0;
    }

    // removeTransitiveEdges, economyGraph.cpp:681-707 (after all marks are known)
    void remove_marked(uint32_t n)
    {
        HEdge *l = list(n);
        uint32_t w = 0;
        for (uint32_t r = 0; r < size[n]; ++r) if (!l[r].mark) l[w++] = l[r];
        removed += size[n] - w;
        size[n] = w;
    }
};

// fn(t, T) on T host threads (T = 1 for small inputs)
template <typename Fn>
void on_all_cores(size_t work_items, Fn fn)
{
    static const unsigned forced = [] { const char *e = getenv("SAGE2GPU_HOST_THREADS"); return e ? (unsigned)atoi(e) : 0u; }();
    unsigned T = forced ? forced : std::thread::hardware_concurrency();
    if (T > 32) T = 32;
    if (T < 1 || work_items < 4096) T = 1;
    if (T == 1) { fn(0u, 1u); return; }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < T; ++t) pool.emplace_back(fn, t, T);
    for (auto &th : pool) th.join();
}

}  // namespace

// buildOverlapGraphEconomy, economyGraph.cpp:513-564
static void traverse(Walk &w, uint32_t nS, std::vector<uint32_t> &queue)
{
    uint32_t counter = 0;
    for (uint32_t i = 0; i < nS; ++i) {
        if (w.state[i] != 0) continue;
        queue.clear();
        size_t qs = 0;
        queue.push_back(i);
        while (qs < queue.size()) {
            const uint32_t n1 = queue[qs++];
            if (w.state[n1] == 0) w.insert_all(n1, counter);
            if (w.size[n1] == 0) continue;          // :525
            if (w.state[n1] == 1) {
                for (uint32_t x = 0; x < w.size[n1]; ++x) {
                    const uint32_t n2 = w.list(n1)[x].node;
                    if (w.state[n2] == 0) { queue.push_back(n2); w.insert_all(n2, counter); }
                }
                w.mark_transitive(n1);
            }
            if (w.state[n1] == 2) {
                for (uint32_t x = 0; x < w.size[n1]; ++x) {
                    const uint32_t n2 = w.list(n1)[x].node;
                    if (w.state[n2] != 1) continue;
                    for (uint32_t y = 0; y < w.size[n2]; ++y) {
                        const uint32_t n3 = w.list(n2)[y].node;
                        if (w.state[n3] == 0) { queue.push_back(n3); w.insert_all(n3, counter); }
                    }
                    w.mark_transitive(n2);
                }
            }
        }
    }
}

static float run_walk(const PhaseCInput &in, PhaseCOutput *outp, std::vector<uint32_t> *order)
{
    PhaseCOutput dummy;
    PhaseCOutput &out = outp ? *outp : dummy;
    const auto t0 = std::chrono::steady_clock::now();
    Walk w(in);
    if (order) { order->assign(in.nS, 0); w.order = order; }
    w.slot.init(in.nS + 2 * in.nB);
    w.state.reserve(in.nS + 2 * in.nB); w.node_id.reserve(in.nS + 2 * in.nB); w.node_len.reserve(in.nS + 2 * in.nB);
    for (uint64_t s = 0; s < in.nS; ++s) w.node(in.s_ids[s] + 1, 0, in.s_len[s]);
    const uint32_t nS = (uint32_t)in.nS;
    const uint64_t nC = in.nS ? in.cand_off[in.nS] : 0;

    // phase-B neighbours of S
    for (uint64_t e = 0; e < in.nB; ++e) {
        const uint32_t a = (uint32_t)(in.edgesB[2 * e] >> 32), b = (uint32_t)in.edgesB[2 * e];
        const uint32_t ia = w.slot.find(a), ib = w.slot.find(b);
        const bool a_in_s = ia != kNone && ia < nS;
        const bool b_in_s = ib != kNone && ib < nS;
        if (a_in_s && ib == kNone) w.node(b, 4, in.edgesB_len[e] >> 16);
        if (b_in_s && ia == kNone) w.node(a, 4, in.edgesB_len[e] & 0xFFFFu);
    }
    const size_t n_nodes = w.state.size();

    static thread_local Workspace ws;

    // candidate -> node of its read2, on all cores (the id map is read-only from here on)
    ws.cnode.resize(nC);
    uint32_t *cnode = ws.cnode.data();
    w.cnode = cnode;
    on_all_cores(nC, [&](unsigned t, unsigned T) {
        const uint64_t lo = nC * t / T, hi = nC * (t + 1) / T;
        for (uint64_t q = lo; q < hi; ++q) cnode[q] = w.slot.find((uint32_t)(in.cand[q] >> 32));
    });

    // room of every list: own candidates + candidates naming the node + phase-B records
    std::vector<uint32_t> &room = ws.room, &eb_a = ws.eb_a, &eb_b = ws.eb_b;
    room.assign(n_nodes, 0);
    for (uint32_t s = 0; s < nS; ++s) room[s] = in.cand_off[s + 1] - in.cand_off[s];
    for (uint64_t q = 0; q < nC; ++q) if (cnode[q] != kNone) room[cnode[q]]++;
    eb_a.resize(in.nB); eb_b.resize(in.nB);
    for (uint64_t e = 0; e < in.nB; ++e) {
        eb_a[e] = w.slot.find((uint32_t)(in.edgesB[2 * e] >> 32));
        eb_b[e] = w.slot.find((uint32_t)in.edgesB[2 * e]);
        if (eb_a[e] != kNone) room[eb_a[e]]++;
        if (eb_b[e] != kNone) room[eb_b[e]]++;
    }
    w.start.resize(n_nodes + 1);
    w.size.assign(n_nodes, 0);
    uint64_t total = 0;
    for (size_t n = 0; n < n_nodes; ++n) { w.start[n] = total; total += room[n]; }
    w.start[n_nodes] = total;
    w.pool = ws.get_pool(total);
    if (total && !w.pool) return -1.f;

    // every phase-B entry of every needed node
    for (uint64_t e = 0; e < in.nB; ++e) {
        const uint64_t w0 = in.edgesB[2 * e], w1 = in.edgesB[2 * e + 1];
        const uint32_t a = (uint32_t)(w0 >> 32), b = (uint32_t)w0;
        const uint32_t type = (uint32_t)(w1 >> 20) & 3u, length = (uint32_t)(w1 & 0xFFFFFu);
        const uint32_t ia = eb_a[e], ib = eb_b[e];
        if (ia != kNone) w.push(ia, HEdge{ b, ib, (uint8_t)type, 0, length });
        if (ib != kNone) {
            const uint32_t la = in.edgesB_len[e] & 0xFFFFu, lb = in.edgesB_len[e] >> 16;
            w.push(ib, HEdge{ a, ia, (uint8_t)rev_type(type), 0, (la - (lb - length)) & 0xFFFFFu });
        }
    }

    const auto t_setup = std::chrono::steady_clock::now();
    { std::vector<uint32_t> queue; queue.reserve(nS); traverse(w, nS, queue); }

    const auto t_walk = std::chrono::steady_clock::now();
    if (!outp) {        // the caller only wants the exploration order: lists, marks and filtering are rebuilt on the device
        if (getenv("SAGE2GPU_PHASE_C_TIMING"))
            fprintf(stderr, "[phase C host, order only] nS %llu nC %llu | setup %.2f walk %.2f ms\n", (unsigned long long)in.nS, (unsigned long long)nC,
                    std::chrono::duration<float, std::milli>(t_setup - t0).count(), std::chrono::duration<float, std::milli>(t_walk - t_setup).count());
        return std::chrono::duration<float, std::milli>(t_walk - t0).count();
    }
    // Marking and removal, after the walk.  When the reference marks a node, every neighbour already has all its
    // edges (insertAllEdgesOfRead ran for it and nothing is appended to an explored node's list), the node's own list
    // was sorted at the end of its insertAllEdgesOfRead, and a node's marked edges are removed only after all its
    // neighbours were marked: so the marks are a function of the complete lists alone and every node can be marked
    // independently -- here on all host cores (the quadratic part of the phase: sum over nodes of degree^2).
    {
        std::vector<uint32_t> todo;
        for (uint32_t i = 0; i < nS; ++i) if (w.state[i] == 2 && w.size[i] > 0) todo.push_back(i);
        on_all_cores(todo.size(), [&](unsigned t, unsigned T) {
            std::vector<uint8_t> scratch(n_nodes, 0);
            for (size_t x = t; x < todo.size(); x += T) w.compute_marks(todo[x], scratch);
        });
        for (uint32_t i : todo) w.remove_marked(i);
    }

    const auto t_mark = std::chrono::steady_clock::now();
    // what convertGraph consumes from the lists of S reads (overlapGraph.cpp:93-112)
    out.edges.clear();
    {
        size_t kept = 0;
        for (uint32_t i = 0; i < nS; ++i) kept += w.size[i];
        out.edges.reserve(kept + 2);      // every kept entry appears from both ends; the larger-id end emits it
    }
    for (uint32_t i = 0; i < nS; ++i) {
        const uint32_t a = w.node_id[i];
        const HEdge *l = w.list(i);
        for (uint32_t x = 0; x < w.size[i]; ++x)
            if (l[x].id > a) {
                out.edges.push_back(((uint64_t)a << 32) | l[x].id);
                out.edges.push_back(((uint64_t)l[x].type << 20) | l[x].length);
            }
    }
    out.inserted = w.inserted;
    out.removed = w.removed;
    const auto t_end = std::chrono::steady_clock::now();
    if (getenv("SAGE2GPU_PHASE_C_TIMING")) {
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<float, std::milli>(b - a).count(); };
        fprintf(stderr, "[phase C host] nS %llu nC %llu nB %llu nodes %zu pool %llu | setup %.2f walk %.2f mark %.2f emit %.2f ms\n",
                (unsigned long long)in.nS, (unsigned long long)nC, (unsigned long long)in.nB, n_nodes, (unsigned long long)total,
                ms(t0, t_setup), ms(t_setup, t_walk), ms(t_walk, t_mark), ms(t_mark, t_end));
    }
    return std::chrono::duration<float, std::milli>(t_end - t0).count();
}

float run_host_phase_c(const PhaseCInput &in, PhaseCOutput &out) { return run_walk(in, &out, nullptr); }

// The walk's traversal alone: order[s] = when S read s was explored (insertAllEdgesOfRead ran for it).  That order is all the
// final lists depend on: an overlap is inserted by whichever end point is explored first (:605).  (The product takes the order
// from run_host_phase_c_order_lists; this one is what the tests hold it against.)
float run_host_phase_c_order(const PhaseCInput &in, std::vector<uint32_t> &order) { return run_walk(in, nullptr, &order); }

// The traversal of economyGraph.cpp:513-564 from the lists the device has sorted (PhaseCLists).  The list of read n as
// insertAllEdgesOfRead leaves it (:580-638) = its own candidates towards reads that are still unexplored (:605) + the
// twins that the reads explored BEFORE it pushed into its list, i.e. the twin entries whose owner is explored by now (an
// explored owner met n unexplored, so it did insert).  Nothing is appended to a list after its read is explored.  So the
// list is the device's static list under a filter on `state`, written once, in exploration order, into a compact pool;
// the queue logic then reads those compact lists.  The traversal never follows phase-B records (they lead outside S);
// only whether a list is empty matters (:525, has_b).
float run_host_phase_c_order_lists(const PhaseCLists &in, std::vector<uint32_t> &order)
{
    const auto t0 = std::chrono::steady_clock::now();
    const uint32_t nS = (uint32_t)in.nS;
    order.assign(nS, 0);
    if (nS == 0) return 0.f;
    static thread_local std::vector<uint32_t> pool, start, size, queue;
    static thread_local std::vector<uint8_t> state;
    pool.resize((size_t)in.off[nS] + 1);
    start.assign(nS, 0); size.assign(nS, 0); state.assign(nS, 0);
    queue.clear(); queue.reserve(nS);
    uint32_t *P = pool.data();
    uint8_t *st = state.data();
    uint32_t top = 0, counter = 0;
    auto explore = [&](uint32_t n) {      // insertAllEdgesOfRead
        st[n] = 1;
        order[n] = counter++;
        const uint32_t s0 = top;
        for (const uint32_t *e = in.ent + in.off[n], *ee = in.ent + in.off[n + 1]; e < ee; ++e) {
            const uint32_t other = *e >> 1;
            P[top] = other;
            top += (uint32_t)((st[other] != 0) == (bool)(*e & 1u));      // own: partner unexplored; twin: owner explored
        }
        start[n] = s0; size[n] = top - s0;
    };
    for (uint32_t i = 0; i < nS; ++i) {
        if (st[i] != 0) continue;
        queue.clear();
        size_t qs = 0;
        queue.push_back(i);
        while (qs < queue.size()) {
            const uint32_t n1 = queue[qs++];
            if (st[n1] == 0) explore(n1);
            if (size[n1] == 0 && !in.has_b[n1]) continue;          // :525
            if (st[n1] == 1) {
                const uint32_t *l = P + start[n1];
                for (uint32_t x = 0, e = size[n1]; x < e; ++x) {
                    const uint32_t n2 = l[x];
                    if (st[n2] == 0) { queue.push_back(n2); explore(n2); }
                }
                st[n1] = 2;         // markTransitiveEdge is due (the marks themselves are computed on the device)
            }
            if (st[n1] == 2) {
                const uint32_t *l = P + start[n1];
                for (uint32_t x = 0, e = size[n1]; x < e; ++x) {
                    const uint32_t n2 = l[x];
                    if (st[n2] != 1) continue;
                    const uint32_t *l2 = P + start[n2];
                    for (uint32_t y = 0, e2 = size[n2]; y < e2; ++y) {
                        const uint32_t n3 = l2[y];
                        if (st[n3] == 0) { queue.push_back(n3); explore(n3); }
                    }
                    st[n2] = 2;
                }
            }
        }
    }
    const float ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (getenv("SAGE2GPU_PHASE_C_TIMING"))
        fprintf(stderr, "[phase C host, order from the device's lists] nS %u entries %u | %.2f ms\n", nS, in.off[nS], ms);
    return ms;
}

}  // namespace sg
