// Row tiles of the spatially sorted query points (svr_sort_points): a tile never straddles two groups of 2x2x2
// Morton-adjacent sort cells, so the voxels its rows touch on a coarse feature level form a small box.  Shared by the
// fused forward (tensor-core interpolation of the coarse levels) and the tensor-core scatter of the backward.
#pragma once
#include "common.cuh"

namespace svr {

constexpr int ST_TILE = 128;   // rows per tile (at most)
constexpr int ST_SUPER = 8;    // sort cells per tile group: 2x2x2 Morton-adjacent cells of the 16^3 sort grid

struct StTile {
    int row0, rows;
};

// tiles[0 .. *n_tiles) from cell_start (first sorted row of every (scene, cell), n_groups * ST_SUPER + 1 entries);
// tiles must hold ceil(total_rows / ST_TILE) + n_groups entries.  scatter_tc.cu
int launch_st_tiles(const int *cell_start, int n_groups, StTile *tiles, int *n_tiles, cudaStream_t st);

}  // namespace svr
