/*
 * CUDA device layer.
 *
 * Replaces the reference's Vulkan layer (src/vulkan.c:153-225: instance,
 * physical device 0, one compute queue, VMA allocator, command pool, eight
 * pipelines) and its context wrapper (src/vkhel.c:4-13) with:
 *   - one CUDA device + one non-blocking stream per context (the "queue");
 *   - a stream-ordered cudaMemPool for vector storage (the VMA role), with an
 *     unlimited release threshold so create/destroy cycles recycle blocks;
 *   - a small cache of pinned staging buffers for map()/unmap();
 *   - lazily uploaded device mirrors of the NTT tables, and a cache of RNS
 *     descriptor arrays.
 * Work is enqueued asynchronously; the reference waits on a fence after every
 * op (src/vector.c:327), here only map/dbgprint/destroy/sync wait.
 */
#include <string.h>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "vkhel_ext.h"

/* cudaSetDevice on the already-current device is a thread-local lookup;
 * calling it on every entry keeps contexts on different devices (one host
 * thread each, or interleaved on one thread) independent of each other and of
 * whatever else in the process moves the current device. */
static void enter_device(int device) {
	CUDA_CHECK(cudaSetDevice(device));
}

static void ctx_enter(const struct vkhel_ctx *ctx) {
	enter_device(ctx->dev.device);
}

/* ---- RNS plan cache ------------------------------------------------------- */
struct rns_plan {
	std::vector<uint64_t> serials;
	limb_desc *descs; /* device */
};

struct plan_cache {
	std::vector<rns_plan> plans;
};

/* ---- live contexts ----------------------------------------------------------
 * NTT tables belong to no context, but a context may hold recorded transforms
 * that use them (vector.cu): destroying tables first launches those. */
static std::mutex g_registry_lock;
static std::vector<struct vkhel_ctx *> g_registry;

static void registry_add(struct vkhel_ctx *ctx) {
	std::lock_guard<std::mutex> guard(g_registry_lock);
	g_registry.push_back(ctx);
}

static void registry_remove(struct vkhel_ctx *ctx) {
	std::lock_guard<std::mutex> guard(g_registry_lock);
	for (size_t i = 0; i < g_registry.size(); i++) {
		if (g_registry[i] == ctx) {
			g_registry.erase(g_registry.begin() + i);
			break;
		}
	}
}

/* ---- context -------------------------------------------------------------- */
extern "C" int vkhel_device_count(void) {
	int count = 0;
	cudaError_t err = cudaGetDeviceCount(&count);
	if (err != cudaSuccess) {
		(void) cudaGetLastError();
		return 0;
	}
	return count;
}

extern "C" void device_ctx_init(struct device_ctx *dev, int device) {
	int count = 0;
	cudaError_t err = cudaGetDeviceCount(&count);
	if (err != cudaSuccess || count == 0) {
		VK_DIE("no usable CUDA device (%s); vkhel has no CPU fallback",
				err != cudaSuccess ? cudaGetErrorString(err)
				: "device count is 0");
	}
	VK_REQUIRE(device >= 0 && device < count,
			"device %d out of range (have %d)", device, count);
	VK_REQUIRE(device < VKHEL_MAX_DEVICES,
			"device ordinal %d not supported", device);

	memset(dev, 0, sizeof(*dev));
	dev->device = device;
	enter_device(device);

	cudaDeviceProp prop;
	CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
	VK_REQUIRE(prop.major >= 10,
			"device %d (%s) is sm_%d%d; this build carries sm_100a code only",
			device, prop.name, prop.major, prop.minor);
	dev->sm_count = prop.multiProcessorCount;
	dev->smem_optin = prop.sharedMemPerBlockOptin;
	dev->l2_bytes = (size_t) prop.l2CacheSize;

	/* the reference prints its choice too (src/vulkan.c:171-172) */
	printf("using physical device %d: %s\n", device, prop.name);
	fflush(stdout);

	cudaStream_t stream;
	CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
	dev->stream = stream;
	cudaStream_t copy_stream;
	CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
	dev->stream_h2d = copy_stream;
	CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
	dev->stream_d2h = copy_stream;
	CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
	dev->stream_aux = copy_stream;
	cudaEvent_t fork_event;
	CUDA_CHECK(cudaEventCreateWithFlags(&fork_event, cudaEventDisableTiming));
	dev->ev_scratch = fork_event;
	CUDA_CHECK(cudaEventCreateWithFlags(&fork_event, cudaEventDisableTiming));
	dev->ev_aux = fork_event;
	for (int i = 0; i < VKHEL_FORK_EVENTS; i++) {
		CUDA_CHECK(cudaEventCreateWithFlags(&fork_event, cudaEventDisableTiming));
		dev->fork_ev[i] = fork_event;
		dev->fork_serial[i] = 0;
	}

	cudaMemPoolProps props;
	memset(&props, 0, sizeof(props));
	props.allocType = cudaMemAllocationTypePinned;
	props.handleTypes = cudaMemHandleTypeNone;
	props.location.type = cudaMemLocationTypeDevice;
	props.location.id = device;
	cudaMemPool_t pool;
	CUDA_CHECK(cudaMemPoolCreate(&pool, &props));
	uint64_t threshold = UINT64_MAX;
	CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold,
				&threshold));
	dev->mem_pool = pool;

	dev->plan_cache = new plan_cache();
}

extern "C" void device_ctx_finish(struct device_ctx *dev) {
	enter_device(dev->device);
	/* struct vkhel_ctx is exactly its device_ctx (priv/vkhel.h) */
	defer_destroy((struct vkhel_ctx *) dev);
	readahead_destroy((struct vkhel_ctx *) dev);
	registry_remove((struct vkhel_ctx *) dev);
	cudaStream_t stream = (cudaStream_t) dev->stream;
	CUDA_CHECK(cudaStreamSynchronize((cudaStream_t) dev->stream_h2d));
	CUDA_CHECK(cudaStreamSynchronize((cudaStream_t) dev->stream_d2h));
	CUDA_CHECK(cudaStreamSynchronize((cudaStream_t) dev->stream_aux));
	CUDA_CHECK(cudaStreamSynchronize(stream));
	CUDA_CHECK(cudaStreamDestroy((cudaStream_t) dev->stream_h2d));
	CUDA_CHECK(cudaStreamDestroy((cudaStream_t) dev->stream_d2h));
	CUDA_CHECK(cudaStreamDestroy((cudaStream_t) dev->stream_aux));
	for (int i = 0; i < 2; i++) {
		if (dev->stream_more[i]) {
			CUDA_CHECK(cudaStreamSynchronize((cudaStream_t) dev->stream_more[i]));
			CUDA_CHECK(cudaStreamDestroy((cudaStream_t) dev->stream_more[i]));
			CUDA_CHECK(cudaEventDestroy((cudaEvent_t) dev->ev_more[i]));
		}
	}
	CUDA_CHECK(cudaEventDestroy((cudaEvent_t) dev->ev_scratch));
	CUDA_CHECK(cudaEventDestroy((cudaEvent_t) dev->ev_aux));
	for (int i = 0; i < VKHEL_FORK_EVENTS; i++) {
		CUDA_CHECK(cudaEventDestroy((cudaEvent_t) dev->fork_ev[i]));
	}

	plan_cache *cache = (plan_cache *) dev->plan_cache;
	for (rns_plan &plan : cache->plans) {
		CUDA_CHECK(cudaFree(plan.descs));
	}
	delete cache;

	for (int i = 0; i < VKHEL_PINNED_SLOTS; i++) {
		if (dev->pinned[i].event) {
			CUDA_CHECK(cudaEventDestroy((cudaEvent_t) dev->pinned[i].event));
		}
		if (dev->pinned[i].ptr) {
			CUDA_CHECK(cudaFreeHost(dev->pinned[i].ptr));
		}
	}
	if (dev->flush_buf) {
		CUDA_CHECK(cudaFree(dev->flush_buf));
	}
	if (dev->scratch) {
		CUDA_CHECK(cudaFree(dev->scratch));
	}
	CUDA_CHECK(cudaMemPoolDestroy((cudaMemPool_t) dev->mem_pool));
	CUDA_CHECK(cudaStreamDestroy(stream));
	memset(dev, 0, sizeof(*dev));
}

extern "C" struct vkhel_ctx *vkhel_ctx_create_device(int device) {
	struct vkhel_ctx *ctx = (struct vkhel_ctx *) calloc(1, sizeof(*ctx));
	VK_REQUIRE(ctx, "out of host memory");
	device_ctx_init(&ctx->dev, device);
	registry_add(ctx);
	return ctx;
}

extern "C" struct vkhel_ctx *vkhel_ctx_create(void) {
	int device = 0;
	const char *env = getenv("VKHEL_DEVICE");
	if (env && *env) {
		device = atoi(env);
	}
	return vkhel_ctx_create_device(device);
}

extern "C" void vkhel_ctx_destroy(struct vkhel_ctx *ctx) {
	if (!ctx) {
		return;
	}
	device_ctx_finish(&ctx->dev);
	free(ctx);
}

extern "C" int vkhel_ctx_device(const struct vkhel_ctx *ctx) {
	return ctx->dev.device;
}

extern "C" void vkhel_ctx_sync(struct vkhel_ctx *ctx) {
	ctx_enter(ctx);
	defer_flush(ctx);
	CUDA_CHECK(cudaStreamSynchronize(ctx_stream(ctx)));
	CUDA_CHECK(cudaStreamSynchronize((cudaStream_t) ctx->dev.stream_h2d));
	CUDA_CHECK(cudaStreamSynchronize((cudaStream_t) ctx->dev.stream_d2h));
}

extern "C" void *vkhel_ctx_stream(struct vkhel_ctx *ctx) {
	/* the caller is about to order its own work against ours: launch what is
	 * recorded, and from now on record nothing (vector.cu, defer_transform) --
	 * a cached stream handle must see every later call already enqueued */
	ctx_enter(ctx);
	defer_flush(ctx);
	ctx->dev.stream_exposed = 1;
	return ctx->dev.stream;
}

extern "C" uint64_t vkhel_ctx_launch_count(const struct vkhel_ctx *ctx) {
	defer_flush((struct vkhel_ctx *) ctx);
	return ctx->dev.launches;
}

extern "C" uint64_t vkhel_ctx_launch_count_noflush(const struct vkhel_ctx *ctx) {
	return ctx->dev.launches;
}

extern "C" void vkhel_ctx_deferred_stats(const struct vkhel_ctx *ctx,
		uint64_t *batches, uint64_t *transforms) {
	if (batches) {
		*batches = ctx->dev.deferred_batches;
	}
	if (transforms) {
		*transforms = ctx->dev.deferred_transforms;
	}
}

extern "C" uint64_t vkhel_ctx_fused_products(const struct vkhel_ctx *ctx) {
	return ctx->dev.fused_products;
}

extern "C" uint64_t vkhel_ctx_readahead_hits(const struct vkhel_ctx *ctx) {
	return ctx->dev.readahead_hits;
}

extern "C" uint64_t vkhel_ctx_lazy_forwards(const struct vkhel_ctx *ctx) {
	return ctx->dev.lazy_forwards;
}

extern "C" void vkhel_ctx_flush(struct vkhel_ctx *ctx) {
	ctx_enter(ctx);
	defer_flush(ctx);
}

extern "C" void vkhel_ctx_flush_l2(struct vkhel_ctx *ctx) {
	ctx_enter(ctx);
	defer_flush(ctx);
	struct device_ctx *dev = &ctx->dev;
	if (!dev->flush_buf) {
		/* twice the L2 so that every line is displaced */
		dev->flush_bytes = dev->l2_bytes ? 2 * dev->l2_bytes : (256u << 20);
		CUDA_CHECK(cudaMalloc(&dev->flush_buf, dev->flush_bytes));
	}
	CUDA_CHECK(cudaMemsetAsync(dev->flush_buf, 0x5a, dev->flush_bytes,
				ctx_stream(ctx)));
}

/* ---- memory --------------------------------------------------------------- */
extern "C" void *device_alloc(struct vkhel_ctx *ctx, size_t bytes) {
	ctx_enter(ctx);
	void *ptr = NULL;
	if (bytes == 0) {
		bytes = 8;
	}
	CUDA_CHECK(cudaMallocFromPoolAsync(&ptr, bytes,
				(cudaMemPool_t) ctx->dev.mem_pool, ctx_stream(ctx)));
	return ptr;
}

extern "C" void device_free(struct vkhel_ctx *ctx, void *ptr) {
	ctx_enter(ctx);
	if (ptr) {
		CUDA_CHECK(cudaFreeAsync(ptr, ctx_stream(ctx)));
	}
}

/* grow-only scratch buffer owned by the context (stream-ordered use only) */
extern "C" void *device_scratch(struct vkhel_ctx *ctx, size_t bytes) {
	ctx_enter(ctx);
	struct device_ctx *dev = &ctx->dev;
	if (dev->scratch_bytes < bytes) {
		if (dev->scratch) {
			CUDA_CHECK(cudaStreamSynchronize(ctx_stream(ctx)));
			CUDA_CHECK(cudaFree(dev->scratch));
		}
		CUDA_CHECK(cudaMalloc(&dev->scratch, bytes));
		dev->scratch_bytes = bytes;
	}
	return dev->scratch;
}

/* the copy that was reading the slot's buffer when it was released is over */
static void pinned_slot_settle(struct pinned_slot *slot) {
	if (slot->busy) {
		CUDA_CHECK(cudaEventSynchronize((cudaEvent_t) slot->event));
		slot->busy = 0;
	}
}

extern "C" void *pinned_acquire(struct vkhel_ctx *ctx, size_t bytes) {
	ctx_enter(ctx);
	struct device_ctx *dev = &ctx->dev;
	/* small requests share one size class, so that the pieces of a staged
	 * upload and the buffers of single-polynomial maps can use each other's
	 * slots instead of replacing them */
	if (bytes < ((size_t) 1 << 20)) {
		bytes = (size_t) 1 << 20;
	}
	/* In order: a cached buffer that is large enough and idle; for small
	 * requests a fresh buffer in an empty slot (so that consecutive staged
	 * copies overlap instead of queueing on one buffer); a large-enough
	 * buffer whose last copy is still in flight (waited for); a slot whose
	 * too-small buffer is replaced; an uncached allocation. */
	const size_t small = (size_t) 16 << 20;
	int fit = -1, fit_busy = -1, empty = -1, victim = -1;
	for (int i = 0; i < VKHEL_PINNED_SLOTS; i++) {
		struct pinned_slot *slot = &dev->pinned[i];
		if (slot->in_use) {
			continue;
		}

		if (!slot->ptr) {
			empty = empty < 0 ? i : empty;
		} else if (slot->bytes >= bytes) {
			if (!slot->busy) {
				fit = fit < 0 ? i : fit;
			} else if (fit_busy < 0
					|| slot->released < dev->pinned[fit_busy].released) {
				/* the copy that was enqueued first ends first: waiting for
				 * the most recent one would serialise staging and DMA */
				fit_busy = i;
			}
		} else {
			victim = victim < 0 ? i : victim;
		}
	}
	if (fit < 0 && fit_busy >= 0 && (bytes > small || empty < 0
				|| cudaEventQuery((cudaEvent_t) dev->pinned[fit_busy].event)
					== cudaSuccess)) {
		/* (the oldest busy buffer is often idle by now: one query instead of
		 * a fresh allocation) */
		fit = fit_busy;
	}
	if (fit >= 0) {
		struct pinned_slot *slot = &dev->pinned[fit];
		pinned_slot_settle(slot);
		slot->in_use = 1;
		return slot->ptr;
	}
	void *ptr = NULL;
	const int target = empty >= 0 ? empty : victim;
	if (target < 0) {
		/* more simultaneous maps than cache slots: uncached allocation */
		CUDA_CHECK(cudaHostAlloc(&ptr, bytes, cudaHostAllocDefault));
		return ptr;
	}
	struct pinned_slot *slot = &dev->pinned[target];
	if (slot->ptr) {
		pinned_slot_settle(slot);
		CUDA_CHECK(cudaFreeHost(slot->ptr));
	}
	CUDA_CHECK(cudaHostAlloc(&ptr, bytes, cudaHostAllocDefault));
	slot->ptr = ptr;
	slot->bytes = bytes;
	slot->in_use = 1;
	return ptr;
}

extern "C" void pinned_release(struct vkhel_ctx *ctx, void *ptr) {
	ctx_enter(ctx);
	struct device_ctx *dev = &ctx->dev;
	for (int i = 0; i < VKHEL_PINNED_SLOTS; i++) {
		if (dev->pinned[i].ptr == ptr) {
			dev->pinned[i].in_use = 0;
			return;
		}
	}
	CUDA_CHECK(cudaFreeHost(ptr));
}

/* Release a staging buffer that a copy enqueued on `stream` is still reading:
 * the caller does not wait; the buffer is not reused before that copy ends. */
extern "C" void pinned_release_after(struct vkhel_ctx *ctx, void *ptr,
		void *stream) {
	ctx_enter(ctx);
	struct device_ctx *dev = &ctx->dev;
	for (int i = 0; i < VKHEL_PINNED_SLOTS; i++) {
		struct pinned_slot *slot = &dev->pinned[i];
		if (slot->ptr == ptr) {
			if (!slot->event) {
				cudaEvent_t ev;
				CUDA_CHECK(cudaEventCreateWithFlags(&ev,
							cudaEventDisableTiming));
				slot->event = ev;
			}
			CUDA_CHECK(cudaEventRecord((cudaEvent_t) slot->event,
						(cudaStream_t) stream));
			slot->busy = 1;
			slot->in_use = 0;
			slot->released = ++dev->pinned_seq;
			return;
		}
	}
	/* uncached buffer: it is freed here, so the copy has to end first */
	CUDA_CHECK(cudaStreamSynchronize((cudaStream_t) stream));
	CUDA_CHECK(cudaFreeHost(ptr));
}

extern "C" void *vkhel_host_alloc(size_t bytes) {
	void *ptr = NULL;
	CUDA_CHECK(cudaHostAlloc(&ptr, bytes ? bytes : 8, cudaHostAllocPortable));
	return ptr;
}

extern "C" void vkhel_host_free(void *ptr) {
	if (ptr) {
		CUDA_CHECK(cudaFreeHost(ptr));
	}
}

/* ---- NTT table device mirrors -------------------------------------------------
 * Layout of one mirror: [limb_desc (80 B)][2n pairs of (w, w')][the first
 * scaled_tw_pairs(n) inverse pairs times n^-1].  The tables object has no
 * context (reference src/ntt_tables.c:65-87), so the mirror is keyed by device
 * ordinal and freed in tables_destroy. */
/* inv_root[k] * n^-1 and its Shoup companion, k < scaled_tw_pairs(n) */
static void fill_scaled_pairs(const struct vkhel_ntt_tables *ntt,
		ulonglong2 *out) {
	const uint64_t count = scaled_tw_pairs(ntt->n);
	for (uint64_t k = 0; k < count; k++) {
		const uint64_t w = nt_multiply_mod(ntt->inv_roots_of_unity[k],
				ntt->inv_n, ntt->q, 0);
		out[k] = make_ulonglong2(w, nt_compute_barrett_factor(w, ntt->q, 64));
	}
}

static void fill_desc(const struct vkhel_ntt_tables *ntt, limb_desc *desc,
		const ulonglong2 *dev_pairs) {
	const uint64_t n = ntt->n;
	memset(desc, 0, sizeof(*desc));
	desc->tw = dev_pairs;
	desc->q = ntt->q;
	desc->inv_n = ntt->inv_n;
	desc->inv_n_shoup = ntt->inv_n_shoup;
	if (n >= 2) {
		desc->inv_w1n = nt_multiply_mod(ntt->inv_roots_of_unity[1],
				ntt->inv_n, ntt->q, 0);
		desc->inv_w1n_shoup = nt_compute_barrett_factor(desc->inv_w1n,
				ntt->q, 64);
	}
	const struct modulus m = make_modulus(ntt->q);
	desc->mm_d = m.d;
	desc->mm_v = m.v;
	desc->mm_s = m.s;
}

/* Mirrors outlive contexts (tables belong to none), so they cannot come from a
 * context's pool; plain cudaMalloc of a few MiB takes tens of milliseconds on
 * this platform (measured: 18-85 ms each, tools/tables_bench.py).  They are
 * taken from the device's default stream-ordered pool instead, which is kept
 * from trimming; ntt_tables_release_device returns them with cudaFree. */
/* Tables are context-free and may be shared by contexts driven from different
 * threads: creation, adoption and release of their device mirrors (and the
 * one-time pool configuration) are serialised by this lock.  Recursive,
 * because ensure_mirror allocates through ntt_tables_mirror_alloc. */
static std::recursive_mutex g_mirror_lock;

void *ntt_tables_mirror_alloc(struct vkhel_ctx *ctx, size_t bytes) {
	std::lock_guard<std::recursive_mutex> guard(g_mirror_lock);
	ctx_enter(ctx);
	static bool configured[VKHEL_MAX_DEVICES];
	const int device = ctx->dev.device;
	if (!configured[device]) {
		cudaMemPool_t pool;
		CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
		uint64_t threshold = UINT64_MAX;
		CUDA_CHECK(cudaMemPoolSetAttribute(pool,
					cudaMemPoolAttrReleaseThreshold, &threshold));
		configured[device] = true;
	}
	void *ptr = NULL;
	CUDA_CHECK(cudaMallocAsync(&ptr, bytes, ctx_stream(ctx)));
	return ptr;
}

static void *ensure_mirror(struct vkhel_ctx *ctx,
		struct vkhel_ntt_tables *ntt) {
	const int device = ctx->dev.device;
	std::lock_guard<std::recursive_mutex> guard(g_mirror_lock);
	if (ntt->dev_pairs[device]) {
		return ntt->dev_pairs[device];
	}
	ctx_enter(ctx);
	VK_REQUIRE(ntt->q < (1ull << 63),
			"NTT modulus must be below 2^63 (as in the reference, whose "
			"nt_inverse_mod works in int64_t)");
	const size_t bytes = sizeof(limb_desc) + mirror_pair_bytes(ntt->n);
	char *dev_buf = (char *) ntt_tables_mirror_alloc(ctx, bytes);
	char *host_buf = (char *) malloc(bytes);
	VK_REQUIRE(host_buf, "out of host memory");
	fill_desc(ntt, (limb_desc *) host_buf,
			(const ulonglong2 *) (dev_buf + sizeof(limb_desc)));
	ulonglong2 *pairs = (ulonglong2 *) (host_buf + sizeof(limb_desc));
	for (uint64_t k = 0; k < ntt->n; k++) {
		pairs[k] = make_ulonglong2(ntt->roots_of_unity[k],
				ntt->roots_barrett_factors[k]);
		pairs[ntt->n + k] = make_ulonglong2(ntt->inv_roots_of_unity[k],
				ntt->inv_roots_barrett_factors[k]);
	}
	fill_scaled_pairs(ntt, pairs + 2 * ntt->n);
	/* on the context's stream, after the allocation; completed here so that
	 * the host buffer can be freed right away and other contexts of this
	 * device can use the mirror */
	CUDA_CHECK(cudaMemcpyAsync(dev_buf, host_buf, bytes, cudaMemcpyHostToDevice,
				ctx_stream(ctx)));
	CUDA_CHECK(cudaStreamSynchronize(ctx_stream(ctx)));
	free(host_buf);
	ntt->dev_pairs[device] = dev_buf;
	return dev_buf;
}

/* tables_device.cu generated the pairs in place: write the descriptor in front
 * of them and register the buffer as this device's mirror */
void ntt_tables_adopt_mirror(struct vkhel_ctx *ctx,
		struct vkhel_ntt_tables *ntt, char *dev_buf) {
	std::lock_guard<std::recursive_mutex> guard(g_mirror_lock);
	ctx_enter(ctx);
	limb_desc desc;
	fill_desc(ntt, &desc, (const ulonglong2 *) (dev_buf + sizeof(limb_desc)));
	CUDA_CHECK(cudaMemcpyAsync(dev_buf, &desc, sizeof(desc),
				cudaMemcpyHostToDevice, ctx_stream(ctx)));
	/* the scaled top of the inverse heap: at most 1024 products, made here
	 * from the host arrays the generator has already copied back */
	ulonglong2 scaled[1 << SCALED_TW_MAX_LOG2];
	fill_scaled_pairs(ntt, scaled);
	CUDA_CHECK(cudaMemcpyAsync(dev_buf + sizeof(limb_desc)
				+ 2 * ntt->n * sizeof(ulonglong2), scaled,
				scaled_tw_pairs(ntt->n) * sizeof(ulonglong2),
				cudaMemcpyHostToDevice, ctx_stream(ctx)));
	CUDA_CHECK(cudaStreamSynchronize(ctx_stream(ctx)));
	ntt->dev_pairs[ctx->dev.device] = dev_buf;
}

const limb_desc *ntt_tables_device_desc(struct vkhel_ctx *ctx,
		struct vkhel_ntt_tables *ntt) {
	return (const limb_desc *) ensure_mirror(ctx, ntt);
}

extern "C" void ntt_tables_release_device(struct vkhel_ntt_tables *ntt) {
	{
		std::lock_guard<std::mutex> guard(g_registry_lock);
		for (struct vkhel_ctx *ctx : g_registry) {
			defer_flush_tables(ctx, ntt);
		}
	}
	std::lock_guard<std::recursive_mutex> guard(g_mirror_lock);
	for (int device = 0; device < VKHEL_MAX_DEVICES; device++) {
		if (!ntt->dev_pairs[device]) {
			continue;
		}
		enter_device(device);
		/* kernels of any context on this device may still read the mirror,
		 * and cudaFree does not wait for them when the block comes from a
		 * stream-ordered pool */
		CUDA_CHECK(cudaDeviceSynchronize());
		CUDA_CHECK(cudaFree(ntt->dev_pairs[device]));
		ntt->dev_pairs[device] = NULL;
	}
}

const limb_desc *rns_plan_device_descs(struct vkhel_ctx *ctx,
		struct vkhel_ntt_tables *const *ntt, uint64_t limbs) {
	plan_cache *cache = (plan_cache *) ctx->dev.plan_cache;
	for (rns_plan &plan : cache->plans) {
		if (plan.serials.size() != limbs) {
			continue;
		}
		bool same = true;
		for (uint64_t l = 0; l < limbs && same; l++) {
			same = plan.serials[l] == ntt[l]->serial
				&& ntt[l]->dev_pairs[ctx->dev.device] != NULL;
		}
		if (same) {
			return plan.descs;
		}
	}
	ctx_enter(ctx);
	defer_flush_held(ctx);   /* (it may use the plan evicted below) */
	std::vector<limb_desc> host(limbs);
	rns_plan plan;
	for (uint64_t l = 0; l < limbs; l++) {
		VK_REQUIRE(ntt[l]->n == ntt[0]->n,
				"all limbs of an RNS basis must share n");
		const char *mirror = (const char *) ensure_mirror(ctx, ntt[l]);
		/* the header of the mirror is the limb's descriptor: rebuild it on
		 * the host instead of reading it back */
		fill_desc(ntt[l], &host[l],
				(const ulonglong2 *) (mirror + sizeof(limb_desc)));
		plan.serials.push_back(ntt[l]->serial);
	}
	CUDA_CHECK(cudaMalloc(&plan.descs, limbs * sizeof(limb_desc)));
	CUDA_CHECK(cudaMemcpy(plan.descs, host.data(), limbs * sizeof(limb_desc),
				cudaMemcpyHostToDevice));
	/* bound the cache: drop the oldest plan beyond 64 entries */
	if (cache->plans.size() >= 64) {
		CUDA_CHECK(cudaStreamSynchronize(ctx_stream(ctx)));
		CUDA_CHECK(cudaFree(cache->plans.front().descs));
		cache->plans.erase(cache->plans.begin());
	}
	cache->plans.push_back(plan);
	return plan.descs;
}

/* ---- timers --------------------------------------------------------------- */
struct vkhel_timer {
	struct vkhel_ctx *ctx;
	cudaEvent_t start, stop;
};

extern "C" struct vkhel_timer *vkhel_timer_create(struct vkhel_ctx *ctx) {
	ctx_enter(ctx);
	struct vkhel_timer *timer =
		(struct vkhel_timer *) calloc(1, sizeof(*timer));
	VK_REQUIRE(timer, "out of host memory");
	timer->ctx = ctx;
	CUDA_CHECK(cudaEventCreate(&timer->start));
	CUDA_CHECK(cudaEventCreate(&timer->stop));
	return timer;
}

extern "C" void vkhel_timer_start(struct vkhel_timer *timer) {
	ctx_enter(timer->ctx);
	defer_flush(timer->ctx);
	CUDA_CHECK(cudaEventRecord(timer->start, ctx_stream(timer->ctx)));
}

extern "C" void vkhel_timer_stop(struct vkhel_timer *timer) {
	ctx_enter(timer->ctx);
	defer_flush(timer->ctx);
	/* join the copy streams first, so that the interval covers uploads and
	 * downloads enqueued since start as well as the kernels */
	struct device_ctx *dev = &timer->ctx->dev;
	cudaEvent_t join = (cudaEvent_t) dev->ev_scratch;
	CUDA_CHECK(cudaEventRecord(join, (cudaStream_t) dev->stream_h2d));
	CUDA_CHECK(cudaStreamWaitEvent(ctx_stream(timer->ctx), join, 0));
	CUDA_CHECK(cudaEventRecord(join, (cudaStream_t) dev->stream_d2h));
	CUDA_CHECK(cudaStreamWaitEvent(ctx_stream(timer->ctx), join, 0));
	CUDA_CHECK(cudaEventRecord(timer->stop, ctx_stream(timer->ctx)));
}

extern "C" double vkhel_timer_elapsed_ms(struct vkhel_timer *timer) {
	ctx_enter(timer->ctx);
	CUDA_CHECK(cudaEventSynchronize(timer->stop));
	float ms = 0.0f;
	CUDA_CHECK(cudaEventElapsedTime(&ms, timer->start, timer->stop));
	return (double) ms;
}

extern "C" void vkhel_timer_destroy(struct vkhel_timer *timer) {
	if (!timer) {
		return;
	}
	ctx_enter(timer->ctx);
	CUDA_CHECK(cudaEventDestroy(timer->start));
	CUDA_CHECK(cudaEventDestroy(timer->stop));
	free(timer);
}
