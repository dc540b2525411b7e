// "tiled" aggregation kernel (placeholder until the TMA-window kernel lands).
#pragma once
#include "common.cuh"

namespace wsage {
inline size_t tiled_workspace_bytes(const wsage_spmm_args*, bool) { return 0; }
inline bool tiled_supported(const wsage_spmm_args*, bool) { return false; }
inline bool tiled_profitable(const wsage_spmm_args*, bool) { return false; }
inline int launch_tiled(const wsage_spmm_args*, cudaStream_t) {
    return fail(WSAGE_EUNSUPPORTED, "%s: %s", "wsage_spmm", "tiled kernel not built");
}
}  // namespace wsage
