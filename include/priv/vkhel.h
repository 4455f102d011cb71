/*
 * Private umbrella header (reference: include/priv/vkhel.h:4-12).  Plain C,
 * no CUDA or Vulkan includes, so the reference's test/ntt.c compiles as is.
 */
#ifndef PRIV_VKHEL_H
#define PRIV_VKHEL_H

#include <vkhel.h>
#include "priv/vector.h"
#include "priv/device.h"

#define DIV_CEIL(a, b) (((a) + (b) - 1) / (b))

struct vkhel_ctx {
	struct device_ctx dev;
};

#endif
