/* gasal.h -- source-compatibility header: same name as the reference's src/gasal.h.
 * All types, enums and the C entry points live in gasal_b200.h. */
#ifndef __GASAL_H__
#define __GASAL_H__
#ifndef GASAL_B200_WITH_CUDA_RUNTIME
#define GASAL_B200_WITH_CUDA_RUNTIME   /* the reference header pulls in cuda_runtime.h; keep cudaStream_t / uint4 the toolkit's */
#endif
#include "gasal_b200.h"
#endif
