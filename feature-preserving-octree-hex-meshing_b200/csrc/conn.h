// fpohm_conn: device-resident result of fpohm_hex_connectivity (connectivity.cu), shared with conforming.cu.
#pragma once
#include "internal.h"

struct fpohm_conn {
	fpohm_ctx *ctx = nullptr;
	int64_t H = 0, nV = 0, nF = 0, nE = 0;
	fpohm::DevBuf<uint32_t> hex;                      // 8/H, the list the tables were built from
	fpohm::DevBuf<uint32_t> F_vs, F_es, E_vs, H_fs;
	fpohm::DevBuf<uint8_t> F_boundary, E_boundary, V_boundary;
	// CSR relations: 0 F.neighbor_hs 1 E.neighbor_fs 2 E.neighbor_hs 3 V.neighbor_vs 4 V.neighbor_es 5 V.neighbor_fs 6 V.neighbor_hs
	fpohm::DevBuf<int64_t> off[7];
	fpohm::DevBuf<uint32_t> val[7];
	int64_t tot[7] = {0, 0, 0, 0, 0, 0, 0};
};

// result of fpohm_extract_surface: the oriented boundary surface with the tables of build_connectivity's Tri/Qua branch
struct fpohm_surface {
	fpohm_ctx *ctx = nullptr;
	int vn = 4;
	int64_t nV = 0, nF = 0, nE = 0, nV_hex = 0, nF_hex = 0;
	int64_t bfs_levels = 0;
	fpohm::DevBuf<double> V;
	fpohm::DevBuf<uint32_t> F_vs, F_es, E_vs;
	fpohm::DevBuf<uint8_t> E_boundary, V_boundary;
	fpohm::DevBuf<int32_t> V_map, V_rev, F_map, F_rev;
	// CSR relations: 0 E.neighbor_fs 1 V.neighbor_vs 2 V.neighbor_es 3 V.neighbor_fs
	fpohm::DevBuf<int64_t> off[4];
	fpohm::DevBuf<uint32_t> val[4];
	int64_t tot[4] = {0, 0, 0, 0};
};

namespace fpohm {
fpohm_conn *conn_build_dev(fpohm_ctx *ctx, DevBuf<uint32_t> &&hex, int64_t H, int64_t nV, bool full);
}
