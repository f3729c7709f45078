// host_phase_c.h -- the serial phase-C walk (economyGraph.cpp:495-707) on the host, fed with the
// candidate lists the GPU produced.  Product code (part of libsage2gpu); shares nothing with oracle/.
#pragma once
#include <stdint.h>
#include <vector>

namespace sg {

struct PhaseCInput {
    uint64_t nS;
    const uint32_t *s_ids;      // [nS] 0-based ids of reads in state 0 after phase B, ascending
    const uint16_t *s_len;      // [nS] their lengths
    const uint32_t *cand_off;   // [nS+1]
    const uint64_t *cand;       // read2(1-based)<<32 | edgeType<<20 | overhang20, reference order
    uint64_t nB;
    const uint64_t *edgesB;     // [2*nB] canonical phase-B records (w0,w1), any order: every record with an
                                // endpoint in S or adjacent (by a phase-B record) to a read in S
    const uint32_t *edgesB_len; // [nB] len(from) | len(to) << 16
    // optional, for run_host_phase_c_order: what the device knows anyway, so that the traversal needs no id map and no
    // phase-B adjacency (phase-B records are inert in the traversal: they lead to reads that are not in S)
    const uint32_t *cand_node = nullptr;   // [nC] index in s_ids of every candidate's read2
    const uint8_t *has_b = nullptr;        // [nS] the read has phase-B records (its list is not empty, economyGraph.cpp:525)
    const uint32_t *comp = nullptr;        // [nS] optional: label (an index < nS) of the read's connected component in the candidate graph;
                                           // components are walked independently (the order inside a component is all that matters)
    uint32_t comp_min_nodes = 4096;        // below this many S reads the components are not worth the threads
};

struct PhaseCOutput {
    std::vector<uint64_t> edges;    // (w0,w1) canonical records owned by state-0 reads (from = that read)
    uint64_t inserted = 0, removed = 0;
};

// returns elapsed host milliseconds
float run_host_phase_c(const PhaseCInput &in, PhaseCOutput &out);
// the traversal alone: order[s] = position of S read s in the exploration sequence (lists / marks / filtering elsewhere)
float run_host_phase_c_order(const PhaseCInput &in, std::vector<uint32_t> &order);

}  // namespace sg
