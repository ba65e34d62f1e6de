// wave_common.h -- host-side dependency geometry shared by the fused matrix-powers kernels.
#pragma once
#include <vector>

#include "nsk_internal.h"

constexpr int WF_GROUP = 16;  // tiles per completion counter (measured on 256^3, k=4: 4 -> 0.760 ms, 16 -> 0.701, 32 -> 0.702)

struct WaveDeps {
    int ntiles = 0, ngroups = 0;
    int reach = 0;                  // max over tiles of (last tile position needed) - (own position)
    std::vector<int> tile_at_pos;   // tile index at each position of global row order
    std::vector<int> pos_of_tile;
    std::vector<int> glo, ghi;      // per tile: first / last position group its columns refer to
};

bool nsk_wave_deps(nsk_csr_t A, const nsk_tiling &T, WaveDeps &out, const char **why);
