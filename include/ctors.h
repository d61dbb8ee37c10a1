/* ctors.h -- reference signatures of src/ctors.h:4-16 as thin C++ overloads over the C ABI. */
#ifndef __CTORS_H__
#define __CTORS_H__
#include "args_parser.h"
inline void gasal_init_streams(gasal_gpu_storage_v* gpu_storage_vec, int max_query_len, int max_target_len, int max_n_alns, Parameters* params) {
  const gasal_params_t p = params->pod();
  gasal_init_streams(gpu_storage_vec, max_query_len, max_target_len, max_n_alns, &p);
}
inline void gasal_destroy_streams(gasal_gpu_storage_v* gpu_storage_vec, Parameters* params) {
  const gasal_params_t p = params->pod();
  gasal_destroy_streams(gpu_storage_vec, &p);
}
/* deprecated in the reference (src/ctors.h:9,11; ctors.cpp:202-252) but still declared there: kept so old callers link */
inline void gasal_gpu_mem_alloc(gasal_gpu_storage_t* gpu_storage, int gpu_max_query_batch_bytes, int gpu_max_target_batch_bytes, int gpu_max_n_alns, Parameters* params) {
  const gasal_params_t p = params->pod();
  gasal_gpu_mem_alloc(gpu_storage, gpu_max_query_batch_bytes, gpu_max_target_batch_bytes, gpu_max_n_alns, &p);
}
inline void gasal_gpu_mem_free(gasal_gpu_storage_t* gpu_storage, Parameters* params) {
  const gasal_params_t p = params->pod();
  gasal_gpu_mem_free(gpu_storage, &p);
}
#endif
