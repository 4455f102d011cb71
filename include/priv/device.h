/*
 * CUDA device layer: the replacement for the reference's Vulkan layer
 * (include/priv/vulkan.h:50-57, src/vulkan.c) and its VMA allocator.
 * Plain C so that test programs can include priv/vkhel.h without CUDA headers.
 */
#ifndef PRIV_DEVICE_H
#define PRIV_DEVICE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VKHEL_PINNED_SLOTS 16
#define VKHEL_FORK_EVENTS 4

struct pinned_slot {
	void *ptr;
	size_t bytes;
	int in_use;
	/* released while a copy out of the buffer was still in flight: `event`
	 * (cudaEvent_t, created on first use) marks the end of that copy and is
	 * waited for before the buffer is handed out again */
	int busy;
	void *event;
	uint64_t released;  /* sequence number of the release (oldest is reused first) */
};

struct device_ctx {
	int device;          /* CUDA ordinal */
	int sm_count;
	size_t smem_optin;   /* max dynamic shared memory per block */
	size_t l2_bytes;
	void *stream;        /* cudaStream_t: every kernel and the map()/copy path */
	void *stream_h2d;    /* vkhel_vector_upload: host -> device copy engine */
	void *stream_d2h;    /* vkhel_vector_download: device -> host copy engine */
	void *ev_scratch;    /* cudaEvent_t used to fork from the compute stream */
	void *stream_aux;    /* second compute stream for sliced transforms */
	void *ev_aux;        /* fork/join event of the sliced transforms */
	/* further compute streams for sliced transforms ($VKHEL_SLICE_STREAMS > 2),
	 * created on first use, with their join events */
	void *stream_more[2];
	void *ev_more[2];
	void *launch_stream; /* non-NULL while kernels go to stream_aux */
	/* Lazy join of a sliced transform (kernels_ntt.cu, run_fast_sliced): the
	 * odd slices of the last sliced transform are still un-joined on
	 * stream_aux.  The next sliced transform onto the same vector with the
	 * same partition continues slice by slice on the same two streams; anything
	 * else joins first (defer_flush).  split_hold: set by the batched transform
	 * entry points while they fetch their pointers, so that this one flush does
	 * not join. */
	int split_active, split_hold, split_by_limb;
	int in_slices;       /* kernels being launched belong to a sliced transform */
	int lazy_out;        /* the forward transform being launched stores lazy values */
	const void *split_dst;
	size_t split_bytes;
	uint64_t split_per, split_units;
	unsigned split_log2n;
	/* Every use of a vector on the compute stream takes the next serial
	 * number; fork_ev[i] was recorded on the compute stream when the serial
	 * stood at fork_serial[i], so it covers every use up to that number.  A
	 * transfer of a vector waits for the oldest of these that covers the
	 * vector's last use -- not for unrelated kernels enqueued after it. */
	uint64_t op_serial;
	void *fork_ev[VKHEL_FORK_EVENTS];
	uint64_t fork_serial[VKHEL_FORK_EVENTS];
	int fork_next;
	void *mem_pool;      /* cudaMemPool_t: stream-ordered allocator */
	struct pinned_slot pinned[VKHEL_PINNED_SLOTS]; /* map() staging cache */
	void *flush_buf;     /* L2 flush scratch, allocated on demand */
	size_t flush_bytes;
	void *plan_cache;    /* RNS plan cache (opaque, C++) */
	void *scratch;       /* transform scratch (two-pass out-of-place) */
	size_t scratch_bytes;
	uint64_t launches;   /* kernels launched so far */
	void *defer;         /* recorded single-vector transforms (opaque, C++) */
	uint64_t deferred_batches;    /* indirect batches launched so far */
	uint64_t deferred_transforms; /* transforms that went out in them */
	uint64_t fused_products;      /* elemmul calls fused into the inverse that followed */
	int stream_exposed;  /* vkhel_ctx_stream() handed the stream out: no more recording */
	uint64_t pinned_seq; /* counter behind pinned_slot.released */
	void *readahead;     /* results of the last recorded batch, for map() (opaque, C++) */
	uint64_t readahead_hits;   /* maps served from a copy started before they were called */
	uint64_t lazy_forwards;    /* held forward transforms launched with lazy stores */
};

void device_ctx_init(struct device_ctx *dev, int device);
void device_ctx_finish(struct device_ctx *dev);

#ifdef __cplusplus
}
#endif

#endif
