/* res.h -- reference signatures of src/res.h:4-8 as thin C++ overloads over the C ABI. */
#ifndef __RES_H__
#define __RES_H__
#include "args_parser.h"
inline gasal_res_t* gasal_res_new_host(uint32_t max_n_alns, Parameters* params) { const gasal_params_t p = params->pod(); return gasal_res_new_host(max_n_alns, &p); }
inline gasal_res_t* gasal_res_new_device_cpy(uint32_t max_n_alns, Parameters* params) { const gasal_params_t p = params->pod(); return gasal_res_new_device_cpy(max_n_alns, &p); }
#endif
