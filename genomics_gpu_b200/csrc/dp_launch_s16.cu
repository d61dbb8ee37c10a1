// dp_launch_s16.cu -- translation unit of the s16x2 DPX fast path (kernels and launcher in dp_s16.cuh).
#include <algorithm>
#include "dp_s16.cuh"
#include "nvb_strip.cuh"
