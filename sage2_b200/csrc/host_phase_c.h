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
};

// The traversal's input when the device has built and sorted every list a read can ever hold (phase_c_device.cu,
// phase_c_sorted_lists): for S read n, in compareLengthBased order (economyGraph.cpp:853-871),
//   * its own candidates, as n would insert them (entry = partner << 1), and
//   * the twins of the candidates of OTHER reads that name n, as those reads would push them into n's list
//     (insertEdgeEconomy, :813-849; entry = owner << 1 | 1).
// partner / owner are S indices.  Which of them exist when n is explored is decided by the traversal alone.
struct PhaseCLists {
    uint64_t nS = 0;
    const uint32_t *off = nullptr;       // [nS+1]
    const uint32_t *ent = nullptr;       // [off[nS]]
    const uint8_t *has_b = nullptr;      // [nS] the read has phase-B records (its list is not empty, economyGraph.cpp:525)
};

struct PhaseCOutput {
    std::vector<uint64_t> edges;    // (w0,w1) canonical records owned by state-0 reads (from = that read)
    uint64_t inserted = 0, removed = 0;
};

// returns elapsed host milliseconds
float run_host_phase_c(const PhaseCInput &in, PhaseCOutput &out);
// the traversal alone: order[s] = position of S read s in the exploration sequence (lists / marks / filtering elsewhere)
float run_host_phase_c_order(const PhaseCInput &in, std::vector<uint32_t> &order);

// the same traversal from the device's pre-sorted lists: no sort, no insertion into other reads' lists, no id map
float run_host_phase_c_order_lists(const PhaseCLists &in, std::vector<uint32_t> &order);

}  // namespace sg
