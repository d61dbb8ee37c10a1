/* interfaces.h -- reference signatures of src/interfaces.h:8-14 as thin C++ overloads over the C ABI. */
#ifndef __GASAL_INTERFACES_H__
#define __GASAL_INTERFACES_H__
#include "args_parser.h"
inline void gasal_host_alns_resize(gasal_gpu_storage_t* gpu_storage, int new_max_alns, Parameters* params) {
  const gasal_params_t p = params->pod();
  gasal_host_alns_resize(gpu_storage, new_max_alns, &p);
}
inline void gasal_set_device(int gpu_select = 0, bool isPrintingProp = true) { gasal_set_device(gpu_select, isPrintingProp ? 1 : 0); }
#endif
