/* gasal_header.h -- umbrella header, same name and include order as the reference's src/gasal_header.h:1-15.
 * A GASAL2 caller switches to this engine by pointing its include path here and linking libgasal_b200.so. */
#ifndef __GASAL_HEADER_H__
#define __GASAL_HEADER_H__
#include "gasal.h"
#include "args_parser.h"
#include "gasal_align.h"
#include "host_batch.h"
#include "ctors.h"
#include "interfaces.h"
#include "res.h"
#endif
