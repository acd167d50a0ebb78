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

namespace fpohm {
fpohm_conn *conn_build_dev(fpohm_ctx *ctx, DevBuf<uint32_t> &&hex, int64_t H, int64_t nV, bool full);
}
