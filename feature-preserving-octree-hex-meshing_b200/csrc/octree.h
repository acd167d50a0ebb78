// fpohm_octree: level-synchronous graded + paired octree, device resident.
#pragma once
#include "internal.h"
#include "mesh.h"

#define FPOHM_MAX_LEVELS 24

namespace fpohm {

// 21-bit 3-D Morton codes (x bit 0 -> code bit 0, y -> 1, z -> 2)
__host__ __device__ __forceinline__ uint64_t part1by2(uint64_t x) {
	x &= 0x1fffffull;
	x = (x | x << 32) & 0x1f00000000ffffull;
	x = (x | x << 16) & 0x1f0000ff0000ffull;
	x = (x | x << 8) & 0x100f00f00f00f00full;
	x = (x | x << 4) & 0x10c30c30c30c30c3ull;
	x = (x | x << 2) & 0x1249249249249249ull;
	return x;
}
__host__ __device__ __forceinline__ uint32_t compact1by2(uint64_t x) {
	x &= 0x1249249249249249ull;
	x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
	x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
	x = (x ^ (x >> 8)) & 0x1f0000ff0000ffull;
	x = (x ^ (x >> 16)) & 0x1f00000000ffffull;
	x = (x ^ (x >> 32)) & 0x1fffffull;
	return (uint32_t)x;
}
__host__ __device__ __forceinline__ uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) {
	return part1by2(x) | (part1by2(y) << 1) | (part1by2(z) << 2);
}

// all per-level sorted arrays of INTERNAL cell codes, concatenated; level l occupies [off[l], off[l+1])
struct LevelTable {
	const uint64_t *code;
	int64_t off[FPOHM_MAX_LEVELS + 2];
	int32_t n_levels;        // number of levels that have internal cells
	int32_t roots[3];
	int32_t n_roots;
	int32_t depth;           // m_MaxDepth: extent(level l) = 1 << (depth - l)
};

} // namespace fpohm

struct fpohm_octree {
	fpohm_ctx *ctx = nullptr;
	fpohm_octree_params prm{};
	int32_t roots[3] = {1, 1, 1};
	int32_t n_roots = 1;
	int32_t depth = 0;
	// internal cells per level (host copy of sizes, device codes concatenated)
	int32_t n_levels = 0;
	int64_t lvl_off[FPOHM_MAX_LEVELS + 2] = {0};
	fpohm::DevBuf<uint64_t> icode;

	// numbering (valid after finalize)
	int64_t n_cells = 0, n_nodes = 0, n_leaves = 0;
	int32_t node_shift = 0;  // node keys are Morton codes of (position >> node_shift)
	fpohm::DevBuf<uint8_t> cell_level;
	fpohm::DevBuf<uint64_t> cell_code;
	fpohm::DevBuf<int32_t> cell_first_child, cell_corner, cell_neigh;
	fpohm::DevBuf<uint64_t> node_key;
	fpohm::DevBuf<int32_t> node_pos, node_neigh;
	fpohm::DevBuf<int32_t> leaf_cell; // hex2Octree_map
	double dbg_close_ms = 0, dbg_number_ms = 0;   // FPOHM_OCTREE_TIMELINE
};
