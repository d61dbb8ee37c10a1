/* host_batch.h -- src/host_batch.h:9-17; the functions are declared (with C linkage) in gasal_b200.h. */
#ifndef __HOST_BACTH_H__
#define __HOST_BACTH_H__
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gasal.h"
#endif
