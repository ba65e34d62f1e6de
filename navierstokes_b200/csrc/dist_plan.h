// dist_plan.h -- host-side plan of a row-partitioned operator (see dist_plan.cpp).
#pragma once
#include <map>
#include <vector>

struct nsk_plan_s {
    int nranks = 1, rank = 0, depth = 1;
    std::vector<int> row_starts;
    int own_begin = 0, own_end = 0, n_global = 0;
    int stage = 0;  // add_rows calls so far: rows of owned (1), ring 1 (2), ... supplied
    bool finalized = false;
    std::vector<int> frontier;             // global ids whose rows are wanted next
    std::vector<int> ptr, cols;            // accumulated rows, global column ids
    std::vector<double> vals;
    std::vector<std::vector<int>> rings;   // rings[r], r = 1..depth: ascending global ids
    std::vector<int> known;                // all ghost ids so far, ascending
    // after finalize
    std::vector<int> ring_start;           // depth+2 entries, ring_start[0] = 0, [1] = n_owned
    std::vector<int> local_cols;           // cols renumbered
    std::vector<int> ghost_gids;           // local order
    struct Req {
        std::vector<int> gids;             // ring-major
        std::vector<int> ring_count;       // depth
        std::vector<int> ring_local_start; // depth: where the slice lands in a local vector
    };
    struct Send {
        std::vector<int> local_idx;        // owned local indices, in the peer's ring-major order
        std::vector<int> ring_count;       // depth
    };
    std::map<int, Req> req;
    std::map<int, Send> sends;
};
