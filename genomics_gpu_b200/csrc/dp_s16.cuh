// dp_s16.cuh -- s16x2 DPX fast path (placeholder until the kernel lands).
#pragma once
#include "dp_common.cuh"
namespace gb200 {
struct FastScratch { void release() {} };
inline bool fast_path_eligible(const DpArgs&, int, int, bool, uint32_t, uint32_t) { return false; }
inline int fast_path_launch(const DpArgs&, int, int, FastScratch&, uint32_t, uint32_t, cudaStream_t) { return 0; }
}  // namespace gb200
