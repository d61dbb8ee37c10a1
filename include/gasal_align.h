/* gasal_align.h -- reference signatures of src/gasal_align.h:114-120 as thin C++ overloads over the C ABI. */
#ifndef __GASAL_ALIGN_H__
#define __GASAL_ALIGN_H__
#include "args_parser.h"
inline void gasal_aln_async(gasal_gpu_storage_t* gpu_storage, const uint32_t actual_query_batch_bytes, const uint32_t actual_target_batch_bytes,
                            const uint32_t actual_n_alns, Parameters* params) {
  const gasal_params_t p = params->pod();
  gasal_aln_async(gpu_storage, actual_query_batch_bytes, actual_target_batch_bytes, actual_n_alns, &p);
}
#endif
