// tcgen05 edge kernel -- placeholder until the tensor-core path lands (next commit).
#pragma once
#include <cuda_fp16.h>
#include <vector>
#include "../../include/difflinker_b200.h"
#include "kernels_simt.cuh"

namespace dl { namespace tc {
constexpr int TN = 128;
constexpr int MAXR = 8;
constexpr bool AVAILABLE = false;
inline dl_status configure() { return DL_OK; }
inline size_t pack_w2(const std::vector<float>&, std::vector<__half>& blob) { return blob.size(); }
inline dl_status launch_edge_tc(const Geom&, const EdgeArgs&, bool, const void*, int, cudaStream_t) { return DL_ERR_UNSUPPORTED; }
inline dl_status selftest(int, float*, float*) { return DL_ERR_UNSUPPORTED; }
}}
