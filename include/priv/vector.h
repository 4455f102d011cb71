/*
 * Device vector.  Mirrors the reference's struct vkhel_vector
 * (include/priv/vector.h:14-20): a back pointer to the context, the length in
 * elements, one device buffer and one host staging slot (one outstanding map
 * per vector).  VkBuffer/VmaAllocation are replaced by plain CUDA pointers.
 */
#ifndef PRIV_VECTOR_H
#define PRIV_VECTOR_H

#include <stdint.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

struct vkhel_ctx;

struct backing_memory {
	void *ptr;    /* device pointer (device) or pinned host pointer (host) */
	size_t bytes; /* capacity */
};

struct vkhel_vector {
	struct vkhel_ctx *ctx;

	size_t length;
	struct backing_memory device;
	struct backing_memory host; /* non-NULL ptr while mapped */

	/* asynchronous transfer tracking (vkhel_vector_upload/_download run on
	 * the context's copy streams): completion event of the last transfer
	 * that touched this vector, and whether the compute stream has yet to
	 * wait for it */
	void *xfer_event;   /* cudaEvent_t, created on first use */
	int xfer_pending;
	/* serial number (device_ctx.op_serial) of the last operation enqueued on
	 * the compute stream that touched this vector; `exposed` once the raw
	 * device pointer has been handed out (then every transfer orders itself
	 * after everything on the compute stream) */
	uint64_t last_op;
	int exposed;
	/* element range [map_offset, map_offset + host.bytes / 8) held by the
	 * current map (vkhel_vector_map_range; the reference's map: everything) */
	size_t map_offset;
	/* read-ahead for map() (vector.cu): a device -> host copy of the whole
	 * vector started before map() was called -- pinned buffer, the event
	 * that ends the copy, and last_op at the time (any later operation on the
	 * vector makes the copy stale) */
	void *ra_ptr;
	void *ra_event;     /* cudaEvent_t, created on first use */
	uint64_t ra_op;
};

void vkhel_vector_dbgprint(const struct vkhel_vector *);

#ifdef __cplusplus
}
#endif

#endif
