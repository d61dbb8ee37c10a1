// dp_launch_wave.cu -- translation unit of the intra-task wavefront kernels (kernels and launcher in dp_wave.cuh).
#include <algorithm>
#include "dp_wave.cuh"
