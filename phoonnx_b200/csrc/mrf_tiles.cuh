// Tile descriptors of the fused multi-receptive-field stage kernel (mrf3_tc.cuh): one thread per tile, so that the kernel's roles
// read ONE 16-byte record per tile instead of searching the utterance table.
#pragma once
#include "conv_tc.cuh"

#define MRF_MAX_RB 3

// {first row of the utterance, its rows, first stage-output row of the tile (o0), utterance}
__global__ void k_mrf_tiles(const int* __restrict__ cu, const int* __restrict__ tile_cu, int B, int rate, int ntiles,
                            int t_step, int post_halo, int4* __restrict__ out) {
    pdl_enter();
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= ntiles) return;
    const int b = find_segment(tile_cu, B, tile);
    const int cb0 = __ldg(cu + b), cb1 = __ldg(cu + b + 1);
    out[tile] = make_int4(cb0 * rate, (cb1 - cb0) * rate, (tile - __ldg(tile_cu + b)) * t_step - post_halo, b);
}
