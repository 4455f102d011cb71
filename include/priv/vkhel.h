/*
 * Private umbrella header: what a translation unit inside the library (or one
 * of the reference's white-box tests, which include "priv/vkhel.h") needs to
 * see of a context.  The reference's header of the same name
 * (include/priv/vkhel.h:4-12) pulls in its Vulkan layer; this one pulls in the
 * CUDA device layer's plain-C description instead, so it compiles with any C
 * compiler and without CUDA or Vulkan headers installed.
 */
#ifndef PRIV_VKHEL_H
#define PRIV_VKHEL_H

#include <vkhel.h>

#include "priv/device.h"
#include "priv/vector.h"

/* integer ceil(a / b); the reference's kernel wrappers size their dispatches
 * with a macro of this name */
#define DIV_CEIL(a, b) (((a) + (b) - 1) / (b))

/* A context is exactly one device-layer instance: a CUDA device, its streams,
 * its memory pool and its caches (see priv/device.h).  Vectors keep a pointer
 * to the context they were created in; NTT tables belong to no context. */
struct vkhel_ctx {
	struct device_ctx dev;
};

#endif
