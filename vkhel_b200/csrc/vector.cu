/*
 * Vectors and the operator entry points.
 *
 * Host side of the reference's src/vector.c retargeted to the CUDA device
 * layer: buffer lifecycle (vector.c:206-260), map/unmap staging
 * (vector.c:262-296), the element-wise entry points (vector.c:298-511) and the
 * forward/inverse transforms (vector.c:513-657).  Where the reference records
 * one dispatch per butterfly group and waits on a fence per stage, each entry
 * point here enqueues one or two kernels on the context's stream and returns.
 */
#include <inttypes.h>
#include <string.h>
#include <unordered_set>
#include <vector>

#include "common.cuh"
#include "vkhel_ext.h"

static void enter(const struct vkhel_ctx *ctx) {
	CUDA_CHECK(cudaSetDevice(ctx->dev.device));
}

/* Device pointer of a vector for an operation on the compute stream.  If an
 * asynchronous upload/download of this vector is still in flight on a copy
 * stream, the compute stream is made to wait for it first (once). */
static inline u64 *dev_u64_nodefer(const struct vkhel_vector *v) {
	struct vkhel_vector *vec = (struct vkhel_vector *) v;
	if (v->xfer_pending) {
		CUDA_CHECK(cudaStreamWaitEvent(ctx_stream(vec->ctx),
					(cudaEvent_t) vec->xfer_event, 0));
		vec->xfer_pending = 0;
	}
	vec->last_op = ++vec->ctx->dev.op_serial;
	return (u64 *) v->device.ptr;
}

/* ---- deferred single-vector transforms ------------------------------------------
 * The reference API transforms one vector per call (src/vector.c:513-657), and
 * an application with many small polynomials calls it in a loop; launched one
 * by one those calls are bound by launch overhead (two kernels, 8 us), not by
 * the GPU.  vkhel_vector_forward_transform / _inverse_transform therefore only
 * RECORD the transform.  Consecutive transforms of the same direction and
 * size on unrelated vectors accumulate -- with any mix of tables, e.g. the
 * limbs of RNS polynomials held as one vector per limb -- and are launched as
 * ONE indirect batch (a table of per-polynomial pointers; one batch per table
 * when the tables are not used equally often) as soon as anything else needs
 * the context: another kind of operation, a vector of the batch being touched
 * again, a transfer, map, sync, a timer, destroy.  Nothing is observable
 * through the API before one of those, so the results are those of the
 * immediate launches.  A single recorded transform is launched exactly as
 * before.  $VKHEL_NO_DEFER=1 turns recording off. */
#define DEFER_MAX 4096
#define DEFER_TABLES 64

struct defer_item {
	ntt_ptrs ptrs;
	unsigned table;     /* index into defer_queue::tables */
	struct vkhel_vector *result;   /* for map()'s read-ahead */
};

/* A recorded vkhel_vector_elemmul or vkhel_vector_elemfma, see "recorded
 * product" below */
struct pending_product {
	bool active;
	bool fma;             /* a*multiplier + b instead of a*b */
	/* the two forward transforms that produced a and b are held back with it
	 * (defer_queue::triple): a candidate for a recorded whole product */
	bool with_forwards;
	/* the two forward transforms that produce a and b are still in the record
	 * (two-pass sizes): they have to be launched before this product, and the
	 * matching inverse transform makes it a recorded inverse-of-product */
	bool after_items;
	struct vkhel_ntt_tables *ntt;   /* after_items: the forward transforms' tables */
	const struct vkhel_vector *a, *b;
	struct vkhel_vector *result;
	uint64_t mod, multiplier;
};

/* A batched forward transform that has been held back, see "held forward
 * transform" below */
struct held_forward {
	bool active;
	/* the previous batched forward transform was followed at once by its
	 * in-place inverse: the next one is held back */
	bool predicted;
	const struct vkhel_vector *operand;
	struct vkhel_vector *result;
	const limb_desc *descs;
	uint64_t limbs, polys, q_max;
	unsigned log2n;
	/* the last batched forward transform that was launched at once */
	const void *last_dst;
	const limb_desc *last_descs;
	uint64_t last_polys, last_serial;
};

/* most whole products recorded for one launch (40 bytes of pointers each) */
#define PRODUCTS_MAX 1024

struct defer_queue {
	held_forward fwd;
	pending_product mul;
	/* ---- recorded whole products (small n, see "recorded whole products") ---- */
	struct {
		bool active;
		defer_item fa, fb;                 /* the forward transforms of a and b */
		struct vkhel_ntt_tables *ntt;
	} triple;
	std::vector<small_product> products;   /* complete four-call products */
	std::vector<struct vkhel_vector *> product_results;
	struct vkhel_ntt_tables *product_tables;
	/* recorded inverse transforms of a product (two-pass sizes): launched as
	 * one indirect batch after the recorded transforms, which produce their
	 * factors */
	std::vector<ntt_ptrs> inv_products;
	std::vector<struct vkhel_vector *> inv_product_results;
	struct vkhel_ntt_tables *inv_product_tables;
	bool inverse;
	uint64_t log2n;
	std::vector<struct vkhel_ntt_tables *> tables;   /* distinct, <= DEFER_TABLES */
	std::vector<defer_item> items;
	std::unordered_set<const void *> reads, writes;
	/* pinned staging for the pointer table: two halves used alternately,
	 * each guarded by an event recorded after the copy that reads it */
	ntt_ptrs *stage[2];
	cudaEvent_t staged[2];
	int half;
};

static defer_queue *defer_get(struct vkhel_ctx *ctx) {
	if (!ctx->dev.defer) {
		/* the staging events belong to the context's device, whatever device
		 * the calling thread last used */
		enter(ctx);
		defer_queue *dq = new defer_queue();
		memset(&dq->fwd, 0, sizeof(dq->fwd));
		dq->mul.active = false;
		dq->mul.with_forwards = false;
		dq->mul.after_items = false;
		dq->triple.active = false;
		dq->product_tables = NULL;
		dq->inv_product_tables = NULL;
		dq->inverse = false;
		dq->log2n = 0;
		dq->half = 0;
		for (int i = 0; i < 2; i++) {
			CUDA_CHECK(cudaHostAlloc((void **) &dq->stage[i],
						DEFER_MAX * sizeof(ntt_ptrs), cudaHostAllocDefault));
			CUDA_CHECK(cudaEventCreateWithFlags(&dq->staged[i],
						cudaEventDisableTiming));
		}
		ctx->dev.defer = dq;
	}
	return (defer_queue *) ctx->dev.defer;
}

/* one indirect launch: `count` pointer pairs already laid out in `host` as
 * [batch][limbs] (polynomial i uses descs[i % limbs]) */
static void defer_launch(struct vkhel_ctx *ctx, defer_queue *dq,
		const std::vector<ntt_ptrs> &host, const limb_desc *descs,
		uint64_t limbs, uint64_t q_max, bool inverse, unsigned log2n,
		bool product = false) {
	const size_t count = host.size();
	if (count <= NTT_INLINE_PTRS) {
		/* short record: the pointers travel in the kernel parameters */
		launch_ntt_indirect(ctx, inverse, NULL, descs, limbs, count,
				log2n, q_max, host.data(), product);
		ctx->dev.deferred_batches++;
		ctx->dev.deferred_transforms += count;
		return;
	}
	size_t done = 0;
	while (done < count) {
		/* whole batch entries per piece of at most DEFER_MAX pointers */
		size_t piece = count - done;
		if (piece > DEFER_MAX) {
			piece = DEFER_MAX / limbs * limbs;
		}
		const int h = dq->half;
		dq->half ^= 1;
		/* the copy that last read this half of the staging buffer */
		CUDA_CHECK(cudaEventSynchronize(dq->staged[h]));
		memcpy(dq->stage[h], host.data() + done, piece * sizeof(ntt_ptrs));
		ntt_ptrs *tab = (ntt_ptrs *) device_alloc(ctx, piece * sizeof(ntt_ptrs));
		CUDA_CHECK(cudaMemcpyAsync(tab, dq->stage[h], piece * sizeof(ntt_ptrs),
					cudaMemcpyHostToDevice, ctx_stream(ctx)));
		CUDA_CHECK(cudaEventRecord(dq->staged[h], ctx_stream(ctx)));
		launch_ntt_indirect(ctx, inverse, tab, descs, limbs, piece,
				log2n, q_max, NULL, product);
		device_free(ctx, tab);   /* stream-ordered: after the kernels */
		ctx->dev.deferred_batches++;
		ctx->dev.deferred_transforms += piece;
		done += piece;
	}
}

static void flush_product(struct vkhel_ctx *ctx, defer_queue *dq);

/* ---- recorded whole products -----------------------------------------------------------
 * For n <= 2^11 a transform is one small kernel, and the reference's product --
 * forward(a), forward(b), elemmul(a, b, c), inverse(c) in place
 * (examples/example.c:18-60) -- is bound by its launches even with the product
 * fused into the inverse transform (three launches).  The four calls are
 * therefore recognised as a unit: elemmul takes the two forward transforms
 * that produced its operands back out of the record (`triple`), and the
 * matching inverse transform turns the three into ONE recorded whole product.
 * Recorded products are independent of each other and of the recorded
 * transforms (the same read/write sets guard them), so a loop of products goes
 * out as one launch of ntt_small_product_kernel, one CTA per product
 * (kernels_ntt_small.cu); the forward results are stored as well, a and b stay
 * observable.  Anything that is not the matching inverse launches the held
 * calls one by one, in order. */
static void launch_recorded_products(struct vkhel_ctx *ctx, defer_queue *dq) {
	const size_t count = dq->products.size();
	if (!count) {
		return;
	}
	struct vkhel_ntt_tables *ntt = dq->product_tables;
	const limb_desc *desc = ntt_tables_device_desc(ctx, ntt);
	if (count <= 4) {
		launch_ntt_small_products(ctx, NULL, dq->products.data(),
				(unsigned) count, desc, (unsigned) ntt->log2n, ntt->q);
	} else {
		const size_t bytes = count * sizeof(small_product);
		const int h = dq->half;
		dq->half ^= 1;
		CUDA_CHECK(cudaEventSynchronize(dq->staged[h]));
		memcpy(dq->stage[h], dq->products.data(), bytes);
		small_product *tab = (small_product *) device_alloc(ctx, bytes);
		CUDA_CHECK(cudaMemcpyAsync(tab, dq->stage[h], bytes,
					cudaMemcpyHostToDevice, ctx_stream(ctx)));
		CUDA_CHECK(cudaEventRecord(dq->staged[h], ctx_stream(ctx)));
		launch_ntt_small_products(ctx, tab, NULL, (unsigned) count, desc,
				(unsigned) ntt->log2n, ntt->q);
		device_free(ctx, tab);   /* stream-ordered: after the kernel */
	}
	ctx->dev.deferred_batches++;
	dq->products.clear();
	dq->product_results.clear();
	dq->product_tables = NULL;
}

/* ---- read-ahead for map() ------------------------------------------------------------
 * The reference reads results by mapping one vector after the other
 * (src/vector.c:270-283: a device -> host copy and a fence wait per map).  A
 * loop of transforms over separate vectors is recorded and launched as one
 * batch (above); the loop of maps that follows would still pay one copy + one
 * wait each, back to back with the caller's own read of the data.  So the
 * context remembers the result vectors of the batch it launched last, and a
 * map() of one of them also starts the device -> host copies of the next
 * READAHEAD_WINDOW ones on the D2H stream; their maps then find the data on
 * its way (or there) instead of starting from nothing.  A copy is used only if
 * no operation has touched its vector since it was started (ra_op), and costs
 * at most READAHEAD_WINDOW vector copies when the caller stops mapping. */
#define READAHEAD_WINDOW 3
/* only vectors up to this size are copied ahead: beyond it a copy takes long
 * enough to hide the latency of starting it, and a speculative staging buffer
 * would be large */
#define READAHEAD_MAX_BYTES ((size_t) 4 << 20)

struct readahead_list {
	std::vector<struct vkhel_vector *> results;
};

static readahead_list *readahead_get(struct vkhel_ctx *ctx) {
	if (!ctx->dev.readahead) {
		ctx->dev.readahead = new readahead_list();
	}
	return (readahead_list *) ctx->dev.readahead;
}

/* give up a speculative copy (stale, or the vector goes away) */
static void readahead_drop(struct vkhel_vector *vec) {
	if (vec->ra_ptr) {
		/* the copy may still be writing the buffer */
		pinned_release_after(vec->ctx, vec->ra_ptr, vec->ctx->dev.stream_d2h);
		vec->ra_ptr = NULL;
	}
}

static void readahead_forget(struct vkhel_vector *vec) {
	readahead_drop(vec);
	readahead_list *ra = (readahead_list *) vec->ctx->dev.readahead;
	if (ra) {
		for (struct vkhel_vector *&v : ra->results) {
			if (v == vec) {
				v = NULL;
			}
		}
	}
}

void readahead_destroy(struct vkhel_ctx *ctx) {
	delete (readahead_list *) ctx->dev.readahead;
	ctx->dev.readahead = NULL;
}

static void xfer_begin(struct vkhel_vector *vec, cudaStream_t copy);
static void xfer_end(struct vkhel_vector *vec, cudaStream_t copy);

/* start the device -> host copy of the whole vector into a fresh staging
 * buffer; the event in ra_event marks its end */
static void readahead_start(struct vkhel_vector *vec) {
	struct vkhel_ctx *ctx = vec->ctx;
	cudaStream_t copy = (cudaStream_t) ctx->dev.stream_d2h;
	vec->ra_ptr = pinned_acquire(ctx, vec->device.bytes);
	vec->ra_op = vec->last_op;
	if (!vec->ra_event) {
		cudaEvent_t ev;
		CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
		vec->ra_event = ev;
	}
	xfer_begin(vec, copy);
	CUDA_CHECK(cudaMemcpyAsync(vec->ra_ptr, vec->device.ptr, vec->device.bytes,
				cudaMemcpyDeviceToHost, copy));
	xfer_end(vec, copy);
	CUDA_CHECK(cudaEventRecord((cudaEvent_t) vec->ra_event, copy));
}

/* map() of `vec` has just been served: start the copies of the vectors that
 * were recorded after it */
static void readahead_advance(struct vkhel_vector *vec) {
	static const bool off = getenv("VKHEL_NO_READAHEAD") != NULL;
	readahead_list *ra = (readahead_list *) vec->ctx->dev.readahead;
	if (off || !ra) {
		return;
	}
	size_t at = 0;
	while (at < ra->results.size() && ra->results[at] != vec) {
		at++;
	}
	struct vkhel_vector *todo[READAHEAD_WINDOW];
	int count = 0;
	for (size_t i = at + 1; i < ra->results.size()
			&& i <= at + READAHEAD_WINDOW; i++) {
		struct vkhel_vector *next = ra->results[i];
		if (next && !next->host.ptr && next->length
				&& next->device.bytes <= READAHEAD_MAX_BYTES
				&& !(next->ra_ptr && next->ra_op == next->last_op)) {
			todo[count++] = next;
		}
	}
	for (int i = 0; i < count; i++) {
		readahead_drop(todo[i]);
		readahead_start(todo[i]);
	}
}

/* launch the recorded transforms (regrouped by tables) and forget them */
static void launch_recorded_items(struct vkhel_ctx *ctx, defer_queue *dq) {
	const size_t count = dq->items.size();
	const size_t ntab = dq->tables.size();
	if (!count) {
		dq->tables.clear();
		return;
	}
	const unsigned log2n = (unsigned) dq->log2n;
	/* An indirect batch costs a pointer-table copy on top of its launches;
	 * a record that short is cheaper launched transform by transform:
	 * always a single one, and two where a transform is one launch. */
	if (count == 1 || (count == 2 && ntt_launches_per_transform(log2n) == 1)) {
		for (const defer_item &it : dq->items) {
			struct vkhel_ntt_tables *ntt = dq->tables[it.table];
			launch_ntt(ctx, dq->inverse, it.ptrs.src, it.ptrs.dst,
					ntt_tables_device_desc(ctx, ntt), 1, 1,
					(unsigned) ntt->log2n, ntt->q);
		}
	} else {
		/* the recorded transforms are independent of each other, so they
		 * may be regrouped: by tables (an RNS polynomial held as one vector
		 * per limb arrives as limb 0, limb 1, ... of polynomial after
		 * polynomial) */
		std::vector<std::vector<ntt_ptrs> > by_table(ntab);
		for (const defer_item &it : dq->items) {
			by_table[it.table].push_back(it.ptrs);
		}
		bool rectangular = !by_table[0].empty();
		for (size_t t = 1; t < ntab; t++) {
			rectangular = rectangular
				&& by_table[t].size() == by_table[0].size();
		}
		if (ntab > 1 && rectangular) {
			/* one launch for all tables, laid out [batch][limbs] like
			 * vkhel_vector_forward_transform_rns */
			uint64_t q_max = 0;
			for (struct vkhel_ntt_tables *t : dq->tables) {
				q_max = t->q > q_max ? t->q : q_max;
			}
			const size_t batch = by_table[0].size();
			std::vector<ntt_ptrs> host(count);
			for (size_t b = 0; b < batch; b++) {
				for (size_t t = 0; t < ntab; t++) {
					host[b * ntab + t] = by_table[t][b];
				}
			}
			defer_launch(ctx, dq, host,
					rns_plan_device_descs(ctx, dq->tables.data(), ntab), ntab,
					q_max, dq->inverse, log2n);
		} else {
			for (size_t t = 0; t < ntab; t++) {
				struct vkhel_ntt_tables *ntt = dq->tables[t];
				if (by_table[t].empty()) {
					continue;   /* its transforms became part of a product */
				}
				if (by_table[t].size() == 1) {
					launch_ntt(ctx, dq->inverse, by_table[t][0].src,
							by_table[t][0].dst,
							ntt_tables_device_desc(ctx, ntt), 1, 1,
							(unsigned) ntt->log2n, ntt->q);
				} else {
					defer_launch(ctx, dq, by_table[t],
							ntt_tables_device_desc(ctx, ntt), 1, ntt->q,
							dq->inverse, log2n);
				}
			}
		}
	}
	dq->items.clear();
	dq->tables.clear();
}

/* the recorded inverse transforms of products, one indirect batch */
static void launch_recorded_inverse_products(struct vkhel_ctx *ctx,
		defer_queue *dq) {
	if (dq->inv_products.empty()) {
		return;
	}
	struct vkhel_ntt_tables *ntt = dq->inv_product_tables;
	defer_launch(ctx, dq, dq->inv_products, ntt_tables_device_desc(ctx, ntt), 1,
			ntt->q, true, (unsigned) ntt->log2n, true);
	dq->inv_products.clear();
	dq->inv_product_results.clear();
	dq->inv_product_tables = NULL;
}

/* ---- held forward transform ---------------------------------------------------------
 * The reference stores canonical residues after every transform
 * (nttfwdbutterfly.comp:41-57), and so does every kernel here -- three
 * conditional subtractions per coefficient at the end of a forward transform
 * of the approximate-quotient family, 12 of the 171 us of the forward row pass
 * of the bench workload.  The inverse transform's butterflies, however, accept
 * [0,3q) as they are, so when a batched forward transform is followed at once
 * by the in-place inverse transform of its result with the same tables
 * (forward -> inverse is the reference's headline loop, BASELINE.json), the
 * forward may store after the first subtraction: the lazy values are
 * overwritten by the inverse before anything can observe them.
 *
 * Which call comes next is only known when it arrives, so the forward transform
 * is held back (like the recorded product above) and launched by the next use
 * of the context: lazily by the matching inverse, canonically by anything
 * else (every entry point passes through defer_flush).  Holding a launch back
 * delays the GPU if the application computes on the host before its next
 * call, so a forward transform is only held when the previous one was
 * followed by its inverse (`predicted`): a loop pays for one canonical
 * forward, a single call for nothing.  Callers that hold the stream or a
 * device pointer are never deferred.  $VKHEL_LAZY_FORWARD=0 turns it off. */
static void launch_held_forward(struct vkhel_ctx *ctx, defer_queue *dq,
		bool lazy) {
	held_forward &h = dq->fwd;
	h.active = false;
	enter(ctx);
	/* as sliced_operands below: slices left on the auxiliary streams are
	 * joined by launch_ntt unless this transform continues them -- which only
	 * the context's stream could not do for a pending transfer */
	if (h.operand->xfer_pending || h.result->xfer_pending) {
		ntt_split_join(ctx);
	}
	const u64 *src = dev_u64_nodefer(h.operand);
	u64 *dst = dev_u64_nodefer(h.result);
	launch_ntt(ctx, false, src, dst, h.descs, h.limbs, h.polys, h.log2n,
			h.q_max, lazy);
	ctx->dev.lazy_forwards += lazy;
}

/* forward entry points: hold the transform back (true) or let the caller
 * launch it now */
static bool hold_forward(const struct vkhel_vector *operand,
		struct vkhel_vector *result, const limb_desc *descs, uint64_t limbs,
		uint64_t polys, unsigned log2n, uint64_t q_max) {
	static const bool off = getenv("VKHEL_NO_DEFER") != NULL;
	struct vkhel_ctx *ctx = result->ctx;
	if (off || ctx->dev.stream_exposed || operand->exposed || result->exposed
			|| !ntt_lazy_forward_supported(log2n, q_max)) {
		return false;
	}
	defer_queue *dq = defer_get(ctx);
	/* whatever is recorded goes first (an earlier held forward included), the
	 * slices in flight stay as they are: this transform may continue them */
	ctx->dev.split_hold = 1;
	defer_flush(ctx);
	ctx->dev.split_hold = 0;
	held_forward &h = dq->fwd;
	if (!h.predicted) {
		/* launched by the caller; remembered for the inverse that may follow */
		h.last_dst = result->device.ptr;
		h.last_descs = descs;
		h.last_polys = polys;
		h.last_serial = ctx->dev.op_serial + 2;   /* after the caller's two fetches */
		return false;
	}
	h.active = true;
	h.operand = operand;
	h.result = result;
	h.descs = descs;
	h.limbs = limbs;
	h.polys = polys;
	h.q_max = q_max;
	h.log2n = log2n;
	return true;
}

/* inverse entry points, before anything else touches the context: launch a
 * held forward transform lazily if this is its inverse, and learn the pattern */
static void inverse_follows(const struct vkhel_vector *operand,
		struct vkhel_vector *result, const limb_desc *descs, uint64_t polys) {
	defer_queue *dq = (defer_queue *) result->ctx->dev.defer;
	if (!dq) {
		return;
	}
	held_forward &h = dq->fwd;
	const bool in_place = operand == result;
	if (h.active) {
		/* nothing may have been recorded since the forward transform was held:
		 * a recorded operation could read its result, and would be launched
		 * between the lazy forward and this inverse */
		const bool quiet = dq->items.empty() && dq->products.empty()
			&& dq->inv_products.empty() && !dq->mul.active && !dq->triple.active;
		if (quiet && in_place && h.result == result && h.descs == descs
				&& h.polys == polys) {
			launch_held_forward(result->ctx, dq, true);
		}
		return;   /* (anything else: defer_flush launches it canonically) */
	}
	if (in_place && h.last_dst == result->device.ptr && h.last_descs == descs
			&& h.last_polys == polys
			&& h.last_serial == result->ctx->dev.op_serial) {
		h.predicted = true;
	}
	h.last_dst = NULL;
}

void defer_flush_held(struct vkhel_ctx *ctx) {
	defer_queue *dq = (defer_queue *) ctx->dev.defer;
	if (dq && dq->fwd.active) {
		dq->fwd.predicted = false;
		launch_held_forward(ctx, dq, false);
	}
}

void defer_flush(struct vkhel_ctx *ctx) {
	/* everything that is about to use the context's stream comes through
	 * here: slices still on the auxiliary stream are joined first */
	defer_queue *dq = (defer_queue *) ctx->dev.defer;
	/* a held forward transform that was not followed by its inverse: launched
	 * as if it had never been held, and the next one is not held */
	defer_flush_held(ctx);
	const bool recorded = dq && (!dq->items.empty() || !dq->products.empty()
			|| !dq->inv_products.empty());
	/* (a batched transform that may continue the slices holds the join back
	 * while it fetches its pointers -- but not if recorded work is launched
	 * here, which goes to the context's stream and may touch the same vector) */
	if (!ctx->dev.split_hold || recorded || (dq && dq->mul.active)) {
		ntt_split_join(ctx);
	}
	if (!dq) {
		return;
	}
	if (recorded) {
		CUDA_CHECK(cudaSetDevice(ctx->dev.device));
		/* the vectors a loop of maps is likely to ask for next (read-ahead) */
		readahead_list *ra = readahead_get(ctx);
		ra->results.clear();
		if (dq->items.size() + dq->product_results.size()
				+ dq->inv_product_results.size() > 1) {
			for (const defer_item &it : dq->items) {
				ra->results.push_back(it.result);
			}
			for (struct vkhel_vector *v : dq->product_results) {
				ra->results.push_back(v);
			}
			for (struct vkhel_vector *v : dq->inv_product_results) {
				ra->results.push_back(v);
			}
		}
	}
	/* Order: recorded whole products are independent of everything else; the
	 * recorded transforms come before a held product and before the recorded
	 * inverse-of-products, whose factors they produce. */
	launch_recorded_products(ctx, dq);
	launch_recorded_items(ctx, dq);
	flush_product(ctx, dq);   /* (with its forward transforms, if it holds them) */
	launch_recorded_inverse_products(ctx, dq);
	dq->reads.clear();
	dq->writes.clear();
}

void defer_flush_tables(struct vkhel_ctx *ctx,
		const struct vkhel_ntt_tables *ntt) {
	defer_queue *dq = (defer_queue *) ctx->dev.defer;
	if (!dq) {
		return;
	}
	if (dq->fwd.active || dq->product_tables == ntt || dq->inv_product_tables == ntt
			|| (dq->triple.active && dq->triple.ntt == ntt)
			|| (dq->mul.active && dq->mul.after_items && dq->mul.ntt == ntt)) {
		defer_flush(ctx);
		return;
	}
	for (const struct vkhel_ntt_tables *t : dq->tables) {
		if (t == ntt) {
			defer_flush(ctx);
			return;
		}
	}
}

void defer_destroy(struct vkhel_ctx *ctx) {
	defer_queue *dq = (defer_queue *) ctx->dev.defer;
	if (!dq) {
		return;
	}
	defer_flush(ctx);
	for (int i = 0; i < 2; i++) {
		CUDA_CHECK(cudaEventSynchronize(dq->staged[i]));
		CUDA_CHECK(cudaEventDestroy(dq->staged[i]));
		CUDA_CHECK(cudaFreeHost(dq->stage[i]));
	}
	delete dq;
	ctx->dev.defer = NULL;
}

/* Record the transform if it can join (or start) a deferred batch; false when
 * the caller has to launch it itself. */
static bool defer_transform(bool inverse, const struct vkhel_vector *operand,
		struct vkhel_vector *result, struct vkhel_ntt_tables *ntt) {
	static const bool off = getenv("VKHEL_NO_DEFER") != NULL;
	struct vkhel_ctx *ctx = result->ctx;
	if (off || !ntt_indirect_supported((unsigned) ntt->log2n, ntt->q)) {
		return false;
	}
	/* A caller that holds the context's stream or a raw device pointer
	 * orders its own work against ours by stream position: everything must be
	 * enqueued when the call returns, so nothing is recorded any more. */
	if (ctx->dev.stream_exposed || operand->exposed || result->exposed) {
		return false;
	}
	defer_queue *dq = defer_get(ctx);
	flush_product(ctx, dq);
	const void *rd = operand->device.ptr, *wr = result->device.ptr;
	unsigned table = 0;
	if (!dq->items.empty() || !dq->products.empty()
			|| !dq->inv_products.empty()) {
		while (table < dq->tables.size() && dq->tables[table] != ntt) {
			table++;
		}
		const bool other_batch = !dq->items.empty()
			&& (dq->inverse != inverse || dq->log2n != ntt->log2n
					|| dq->items.size() >= DEFER_MAX);
		if (other_batch || table >= DEFER_TABLES
				|| dq->writes.count(rd) || dq->writes.count(wr)
				|| dq->reads.count(wr)) {
			/* different batch, or this transform depends on a recorded one
			 * (or overwrites what a recorded one still has to read) */
			defer_flush(ctx);
			table = 0;
		}
	}
	if (table == dq->tables.size()) {
		dq->tables.push_back(ntt);
	}
	defer_item item;
	item.ptrs.src = dev_u64_nodefer(operand);
	item.ptrs.dst = dev_u64_nodefer(result);
	item.table = table;
	item.result = result;
	dq->inverse = inverse;
	dq->log2n = ntt->log2n;
	dq->items.push_back(item);
	dq->reads.insert(rd);
	dq->writes.insert(wr);
	return true;
}

/* Device pointer of a vector for an operation that is launched now: recorded
 * transforms go first. */
static inline u64 *dev_u64(const struct vkhel_vector *v) {
	defer_flush(v->ctx);
	return dev_u64_nodefer(v);
}

/* ---- recorded product ---------------------------------------------------------------
 * The reference's polynomial product is the call sequence forward, forward,
 * elemmul, inverse (src/vector.c:388-427,513-657; examples/example.c:18-60).
 * When the inverse transform runs in place on the product, the product itself
 * is never observable, and the row pass of the inverse can multiply while it
 * loads (ntt_rows_kernel<.., MUL, ..>): one launch and one sweep over the
 * vector less.  vkhel_vector_elemmul therefore only records (a, b, result,
 * mod); the very next use of the context decides:
 *   - vkhel_vector_inverse_transform(result, result, tables) with the same
 *     modulus and n == result->length: one fused launch sequence;
 *   - anything else (every other entry point passes through defer_flush or
 *     defer_transform): the product is launched first, as if it had never
 *     been held back.
 * $VKHEL_NO_DEFER=1 turns this off together with the recorded transforms,
 * $VKHEL_NO_FUSED_PRODUCT=1 only this. */
static void launch_recorded_items(struct vkhel_ctx *ctx, defer_queue *dq);

static void flush_product(struct vkhel_ctx *ctx, defer_queue *dq) {
	if (!dq->mul.active) {
		return;
	}
	dq->mul.active = false;
	CUDA_CHECK(cudaSetDevice(ctx->dev.device));
	if (dq->mul.after_items) {
		/* its factors are results of recorded transforms */
		dq->mul.after_items = false;
		launch_recorded_items(ctx, dq);
	}
	if (dq->mul.with_forwards) {
		/* the product did not become a recorded whole product: its two
		 * forward transforms, taken out of the record, go first */
		dq->mul.with_forwards = false;
		dq->triple.active = false;
		struct vkhel_ntt_tables *ntt = dq->triple.ntt;
		const defer_item *fwd[2] = { &dq->triple.fa, &dq->triple.fb };
		for (const defer_item *it : fwd) {
			launch_ntt(ctx, false, it->ptrs.src, it->ptrs.dst,
					ntt_tables_device_desc(ctx, ntt), 1, 1,
					(unsigned) ntt->log2n, ntt->q);
		}
	}
	const pending_product &m = dq->mul;
	if (m.fma) {
		launch_elemfma(ctx, dev_u64_nodefer(m.a), dev_u64_nodefer(m.b),
				dev_u64_nodefer(m.result), m.result->length, m.multiplier,
				m.mod);
	} else {
		launch_elemmul(ctx, dev_u64_nodefer(m.a), dev_u64_nodefer(m.b),
				dev_u64_nodefer(m.result), m.result->length, m.mod);
	}
}

/* the inverse transform that follows a recorded product: true when the fused
 * kernels have been launched */
static bool fuse_product_into_inverse(const struct vkhel_vector *operand,
		struct vkhel_vector *result, struct vkhel_ntt_tables *ntt) {
	struct vkhel_ctx *ctx = result->ctx;
	defer_queue *dq = (defer_queue *) ctx->dev.defer;
	if (!dq || !dq->mul.active) {
		return false;
	}
	const pending_product m = dq->mul;
	if (m.with_forwards) {
		if (operand == result && result == m.result && ntt == dq->triple.ntt) {
			/* forward, forward, elemmul, inverse: one recorded whole product */
			small_product sp;
			sp.a_src = dq->triple.fa.ptrs.src;
			sp.a_dst = dq->triple.fa.ptrs.dst;
			sp.b_src = dq->triple.fb.ptrs.src;
			sp.b_dst = dq->triple.fb.ptrs.dst;
			sp.c = dev_u64_nodefer(result);
			dq->products.push_back(sp);
			dq->product_results.push_back(result);
			dq->product_tables = ntt;
			dq->mul.active = false;
			dq->mul.with_forwards = false;
			dq->triple.active = false;
			ctx->dev.fused_products++;
			ctx->dev.deferred_transforms += 3;
			if (dq->products.size() >= PRODUCTS_MAX) {
				defer_flush(ctx);
			}
			return true;
		}
		return false;   /* the caller's path launches the three held calls */
	}
	if (m.after_items) {
		if (operand == result && result == m.result && ntt == m.ntt) {
			/* forward, forward, elemmul, inverse at a two-pass size: the
			 * inverse transform of the product is recorded; it goes out with
			 * those of the other products of the loop, after the recorded
			 * forward transforms */
			ntt_ptrs ent;
			ent.src = dev_u64_nodefer(m.a);
			ent.src2 = dev_u64_nodefer(m.b);
			ent.dst = dev_u64_nodefer(result);
			dq->inv_products.push_back(ent);
			dq->inv_product_results.push_back(result);
			dq->inv_product_tables = ntt;
			dq->mul.active = false;
			dq->mul.after_items = false;
			ctx->dev.fused_products++;
			ctx->dev.deferred_transforms++;
			return true;
		}
		return false;   /* the caller's path launches transforms and product */
	}
	if (operand != result || result != m.result || m.mod != ntt->q
			|| result->length != ntt->n || !dq->items.empty()) {
		return false;   /* the caller's path launches the product first */
	}
	dq->mul.active = false;
	/* elemfma: the multiplier reduced mod q, with q standing for 0 (the
	 * kernel reads 0 as "product") */
	uint64_t fma_mult = 0;
	if (m.fma) {
		fma_mult = m.multiplier % m.mod;
		fma_mult = fma_mult ? fma_mult : m.mod;
	}
	if (launch_ntt_inverse_of_product(ctx, dev_u64_nodefer(m.a),
				dev_u64_nodefer(m.b), dev_u64_nodefer(result),
				ntt_tables_device_desc(ctx, ntt), 1, 1, (unsigned) ntt->log2n,
				ntt->q, fma_mult)) {
		ctx->dev.fused_products++;
		return true;
	}
	/* no fused kernel for this size or modulus */
	dq->mul.active = true;
	return false;
}

/* An event on the compute stream that covers every operation up to serial
 * number `serial`: the oldest recorded one that does, or a new one. */
static cudaEvent_t fork_event_covering(struct vkhel_ctx *ctx, uint64_t serial,
		bool fresh) {
	struct device_ctx *dev = &ctx->dev;
	int best = -1;
	for (int i = 0; i < VKHEL_FORK_EVENTS && !fresh; i++) {
		if (dev->fork_serial[i] >= serial && dev->fork_serial[i] != 0
				&& (best < 0 || dev->fork_serial[i] < dev->fork_serial[best])) {
			best = i;
		}
	}
	if (best < 0) {
		best = dev->fork_next;
		dev->fork_next = (best + 1) % VKHEL_FORK_EVENTS;
		CUDA_CHECK(cudaEventRecord((cudaEvent_t) dev->fork_ev[best],
					ctx_stream(ctx)));
		/* serial 0 marks an unused slot */
		dev->fork_serial[best] = dev->op_serial ? dev->op_serial
			: ++dev->op_serial;
	}
	return (cudaEvent_t) dev->fork_ev[best];
}

/* start a transfer of `vec` on copy stream `copy`: it must come after the
 * last operation on the compute stream that touched the vector (not after
 * unrelated kernels enqueued since: the upload of the next slice overlaps the
 * transform of this one) and after the previous transfer of this vector */
static void xfer_begin(struct vkhel_vector *vec, cudaStream_t copy) {
	struct vkhel_ctx *ctx = vec->ctx;
	defer_flush(ctx);
	cudaEvent_t fork = fork_event_covering(ctx, vec->last_op, vec->exposed);
	CUDA_CHECK(cudaStreamWaitEvent(copy, fork, 0));
	if (!vec->xfer_event) {
		cudaEvent_t ev;
		CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
		vec->xfer_event = ev;
	} else {
		CUDA_CHECK(cudaStreamWaitEvent(copy, (cudaEvent_t) vec->xfer_event, 0));
	}
}

static void xfer_end(struct vkhel_vector *vec, cudaStream_t copy) {
	CUDA_CHECK(cudaEventRecord((cudaEvent_t) vec->xfer_event, copy));
	vec->xfer_pending = 1;
}

/* ---- lifecycle ---------------------------------------------------------------- */
extern "C" struct vkhel_vector *vkhel_vector_create2(struct vkhel_ctx *ctx,
		uint64_t length, bool zero) {
	VK_REQUIRE(ctx, "vkhel_vector_create: NULL context");
	struct vkhel_vector *vec =
		(struct vkhel_vector *) calloc(1, sizeof(*vec));
	VK_REQUIRE(vec, "out of host memory");
	vec->ctx = ctx;
	vec->length = length;
	vec->device.bytes = length * sizeof(uint64_t);
	vec->device.ptr = device_alloc(ctx, vec->device.bytes);
	if (zero && length) {
		/* reference: vkCmdFillBuffer(0), src/vector.c:152-204 */
		CUDA_CHECK(cudaMemsetAsync(vec->device.ptr, 0, vec->device.bytes,
					ctx_stream(ctx)));
	}
	/* the block may come from a stream-ordered free that is still pending
	 * on the compute stream; count the allocation as a use */
	vec->last_op = ++ctx->dev.op_serial;
	return vec;
}

extern "C" struct vkhel_vector *vkhel_vector_create(struct vkhel_ctx *ctx,
		uint64_t length) {
	return vkhel_vector_create2(ctx, length, true);
}

extern "C" void vkhel_vector_destroy(struct vkhel_vector *vec) {
	if (!vec) {
		return;
	}
	struct vkhel_ctx *ctx = vec->ctx;
	enter(ctx);
	if (vec->host.ptr) {
		/* destroyed while mapped: drop the staging buffer, nothing is
		 * written back */
		pinned_release(ctx, vec->host.ptr);
	}
	readahead_forget(vec);
	if (vec->ra_event) {
		CUDA_CHECK(cudaEventDestroy((cudaEvent_t) vec->ra_event));
	}
	/* stream-ordered free: work already enqueued on the stream (and any
	 * transfer still in flight) completes before the block is reused */
	(void) dev_u64(vec);
	device_free(ctx, vec->device.ptr);
	if (vec->xfer_event) {
		CUDA_CHECK(cudaEventDestroy((cudaEvent_t) vec->xfer_event));
	}
	free(vec);
}

extern "C" struct vkhel_vector *vkhel_vector_dup(struct vkhel_vector *src) {
	/* no zero fill: about to be overwritten (reference vector.c:249-260) */
	struct vkhel_vector *dup = vkhel_vector_create2(src->ctx, src->length,
			false);
	if (src->length) {
		CUDA_CHECK(cudaMemcpyAsync(dev_u64(dup), dev_u64(src),
					src->device.bytes, cudaMemcpyDeviceToDevice,
					ctx_stream(src->ctx)));
	}
	return dup;
}

/* piece of the staged host -> device copies below.  Measured from C on the
 * B200 box (examples/api_e2e.c, 64 polynomials of 512 KiB per step): every
 * piece costs about 10 us of calls (copy, event, slot search) next to 27 us of
 * memcpy per 512 KiB, and in a loop over vectors the DMA of one vector overlaps
 * the memcpy of the next whatever the piece size -- so one piece per
 * polynomial of n = 2^16, several for anything larger. */
#define STAGE_CHUNK_BYTES ((size_t) 512 << 10)

extern "C" void vkhel_vector_copy_from_host(struct vkhel_vector *vec,
		const uint64_t *elements) {
	/* The reference maps (a device->host copy it does not need), memcpys and
	 * unmaps (vector.c:262-268).  Here: one host->device copy.  The source
	 * may be pageable and may be reused by the caller as soon as this
	 * returns, so it is copied into pinned staging buffers (pieces of 512 KiB
	 * from the context's cache, each released when its transfer has ended)
	 * and the transfers are left in flight: the call neither waits for them
	 * nor for the kernels enqueued before it. */
	if (!vec->length) {
		return;
	}
	struct vkhel_ctx *ctx = vec->ctx;
	enter(ctx);
	char *dst = (char *) dev_u64(vec);
	const char *src = (const char *) elements;
	size_t left = vec->device.bytes;
	while (left) {
		const size_t piece = left < STAGE_CHUNK_BYTES ? left : STAGE_CHUNK_BYTES;
		void *stage = pinned_acquire(ctx, piece);
		host_copy(stage, src, piece);
		CUDA_CHECK(cudaMemcpyAsync(dst, stage, piece, cudaMemcpyHostToDevice,
					ctx_stream(ctx)));
		pinned_release_after(ctx, stage, ctx_stream(ctx));
		dst += piece;
		src += piece;
		left -= piece;
	}
}

extern "C" void vkhel_vector_map_range(struct vkhel_vector *vec, void **mem,
		uint64_t offset, uint64_t count) {
	VK_REQUIRE(!vec->host.ptr, "vector is already mapped");
	VK_REQUIRE(offset + count <= vec->length && offset + count >= offset,
			"map_range out of range");
	struct vkhel_ctx *ctx = vec->ctx;
	enter(ctx);
	defer_flush(ctx);
	const bool whole = offset == 0 && count == vec->length;
	vec->map_offset = offset;
	vec->host.bytes = count * sizeof(uint64_t);
	if (whole && vec->ra_ptr && vec->ra_op == vec->last_op) {
		/* a read-ahead copy of exactly this state of the vector */
		CUDA_CHECK(cudaEventSynchronize((cudaEvent_t) vec->ra_event));
		vec->host.ptr = vec->ra_ptr;
		vec->ra_ptr = NULL;
		ctx->dev.readahead_hits++;
	} else {
		readahead_drop(vec);
		vec->host.ptr = pinned_acquire(ctx, vec->host.bytes);
		if (count) {
			/* on the D2H stream, behind the last operation that touched this
			 * vector: only that is waited for, not whatever else the compute
			 * stream holds */
			cudaStream_t copy = (cudaStream_t) ctx->dev.stream_d2h;
			xfer_begin(vec, copy);
			CUDA_CHECK(cudaMemcpyAsync(vec->host.ptr,
						(u64 *) vec->device.ptr + offset, vec->host.bytes,
						cudaMemcpyDeviceToHost, copy));
			xfer_end(vec, copy);
			CUDA_CHECK(cudaStreamSynchronize(copy));
		}
	}
	if (whole) {
		readahead_advance(vec);
	}
	*mem = vec->host.ptr;
}

extern "C" void vkhel_vector_map(struct vkhel_vector *vec, void **mem,
		size_t size) {
	/* `size` is in bytes in the reference's tests and in elements in its
	 * example; staging the whole vector serves both. */
	(void) size;
	vkhel_vector_map_range(vec, mem, 0, vec->length);
}

extern "C" void vkhel_vector_unmap(struct vkhel_vector *vec) {
	VK_REQUIRE(vec->host.ptr, "vector is not mapped");
	enter(vec->ctx);
	/* what was mapped is written back (reference vector.c:291: the whole
	 * vector; a sub-range map writes back its range) */
	if (vec->host.bytes) {
		CUDA_CHECK(cudaMemcpyAsync(dev_u64(vec) + vec->map_offset, vec->host.ptr,
					vec->host.bytes, cudaMemcpyHostToDevice,
					ctx_stream(vec->ctx)));
	}
	/* the staging buffer goes back to the cache and is handed out again only
	 * after this copy has ended; nothing waits here */
	pinned_release_after(vec->ctx, vec->host.ptr, ctx_stream(vec->ctx));
	vec->host.ptr = NULL;
	vec->host.bytes = 0;
	vec->map_offset = 0;
}

extern "C" void vkhel_vector_dbgprint(const struct vkhel_vector *vec) {
	uint64_t *mapped = NULL;
	vkhel_vector_map((struct vkhel_vector *) vec, (void **) &mapped,
			vec->length * sizeof(uint64_t));
	for (size_t i = 0; i < vec->length; i++) {
		printf(i + 1 == vec->length ? "%" PRIu64 : "%" PRIu64 ", ",
				mapped[i]);
	}
	printf("\n");
	vkhel_vector_unmap((struct vkhel_vector *) vec);
}

extern "C" uint64_t vkhel_vector_length(const struct vkhel_vector *vec) {
	return vec->length;
}

extern "C" void *vkhel_vector_device_ptr(struct vkhel_vector *vec) {
	/* the caller may enqueue its own kernels on vkhel_ctx_stream(): from now
	 * on transfers of this vector wait for the whole compute stream */
	vec->exposed = 1;
	return dev_u64(vec);
}

/* Asynchronous transfers run on the context's two copy streams, so that an
 * upload, kernels and a download of different vectors overlap (PCIe is full
 * duplex).  Ordering: a transfer starts after all compute enqueued before the
 * call and after the previous transfer of the same vector; the next operation
 * that touches the vector waits for the transfer (dev_u64). */
extern "C" void vkhel_vector_upload(struct vkhel_vector *vec,
		const uint64_t *src, uint64_t offset, uint64_t count) {
	VK_REQUIRE(offset + count <= vec->length, "upload out of range");
	enter(vec->ctx);
	if (count) {
		cudaStream_t copy = (cudaStream_t) vec->ctx->dev.stream_h2d;
		readahead_drop(vec);   /* the vector changes: a copy taken earlier is stale */
		xfer_begin(vec, copy);
		CUDA_CHECK(cudaMemcpyAsync((u64 *) vec->device.ptr + offset, src,
					count * sizeof(uint64_t), cudaMemcpyHostToDevice, copy));
		xfer_end(vec, copy);
	}
}

extern "C" void vkhel_vector_download(const struct vkhel_vector *cvec,
		uint64_t *dst, uint64_t offset, uint64_t count) {
	struct vkhel_vector *vec = (struct vkhel_vector *) cvec;
	VK_REQUIRE(offset + count <= vec->length, "download out of range");
	enter(vec->ctx);
	if (count) {
		cudaStream_t copy = (cudaStream_t) vec->ctx->dev.stream_d2h;
		xfer_begin(vec, copy);
		CUDA_CHECK(cudaMemcpyAsync(dst, (u64 *) vec->device.ptr + offset,
					count * sizeof(uint64_t), cudaMemcpyDeviceToHost, copy));
		xfer_end(vec, copy);
	}
}

extern "C" void vkhel_vector_copy_peer(struct vkhel_vector *dst,
		uint64_t dst_offset, const struct vkhel_vector *src,
		uint64_t src_offset, uint64_t count) {
	VK_REQUIRE(dst_offset + count <= dst->length
			&& src_offset + count <= src->length, "copy_peer out of range");
	if (!count) {
		return;
	}
	struct vkhel_ctx *sctx = src->ctx, *dctx = dst->ctx;
	/* source side: an event after everything enqueued on its stream */
	enter(sctx);
	const u64 *sp = dev_u64(src) + src_offset;
	cudaEvent_t ready = (cudaEvent_t) sctx->dev.ev_scratch;
	CUDA_CHECK(cudaEventRecord(ready, ctx_stream(sctx)));
	/* destination side: wait for it, then copy on the destination stream */
	enter(dctx);
	u64 *dp = dev_u64(dst) + dst_offset;
	CUDA_CHECK(cudaStreamWaitEvent(ctx_stream(dctx), ready, 0));
	if (sctx->dev.device == dctx->dev.device) {
		CUDA_CHECK(cudaMemcpyAsync(dp, sp, count * sizeof(u64),
					cudaMemcpyDeviceToDevice, ctx_stream(dctx)));
	} else {
		/* direct NVLink path when the devices can address each other;
		 * cudaMemcpyPeerAsync stages through the host otherwise */
		int can = 0;
		CUDA_CHECK(cudaDeviceCanAccessPeer(&can, dctx->dev.device,
					sctx->dev.device));
		if (can) {
			cudaError_t err = cudaDeviceEnablePeerAccess(sctx->dev.device, 0);
			if (err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled) {
				VK_DIE("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(err));
			}
			(void) cudaGetLastError();
		}
		CUDA_CHECK(cudaMemcpyPeerAsync(dp, dctx->dev.device, sp,
					sctx->dev.device, count * sizeof(u64), ctx_stream(dctx)));
	}
	/* the source must not be overwritten before the copy has read it */
	cudaEvent_t done = (cudaEvent_t) dctx->dev.ev_scratch;
	CUDA_CHECK(cudaEventRecord(done, ctx_stream(dctx)));
	enter(sctx);
	CUDA_CHECK(cudaStreamWaitEvent(ctx_stream(sctx), done, 0));
}

/* ---- debug tracing (reference: VKHEL_DEBUG blocks, e.g. vector.c:305-314) --- */
#ifdef VKHEL_DEBUG
#define DBG_VEC(label, v) do { printf("\t%s: ", label); \
		vkhel_vector_dbgprint(v); } while (0)
#define DBG(...) printf(__VA_ARGS__)
#else
#define DBG_VEC(label, v) do { } while (0)
#define DBG(...) do { } while (0)
#endif

/* ---- element-wise entry points ------------------------------------------------
 * As in the reference every op runs over result->length elements and operand
 * lengths are not checked (src/kernels/elemmul.c:164,173-174). */
/* Record a point-wise product or fma as the candidate for the fused
 * inverse-of-product (see "recorded product" above): true when it is held
 * back, false when the caller has to launch it. */
static bool record_pointwise(bool fma, const struct vkhel_vector *a,
		const struct vkhel_vector *b, struct vkhel_vector *result,
		uint64_t multiplier, uint64_t mod) {
	static const bool off = getenv("VKHEL_NO_DEFER") != NULL
		|| getenv("VKHEL_NO_FUSED_PRODUCT") != NULL;
	/* candidates: a power-of-two length the fast transform path covers */
	const uint64_t len = result->length;
	const bool exposed = result->ctx->dev.stream_exposed || a->exposed
		|| b->exposed || result->exposed;   /* see defer_transform */
	if (off || exposed || len < 8 || (len & (len - 1)) != 0
			|| mod >= (1ull << 62)) {
		return false;
	}
	defer_queue *dq = defer_get(result->ctx);
	/* the product of two vectors whose forward transforms are the last two
	 * recorded calls: hold all three for the inverse transform that would
	 * make them a recorded whole product */
	const size_t nitems = dq->items.size();
	const bool small = ntt_small_product_supported((unsigned) dq->log2n, mod);
	static const bool no_batch = getenv("VKHEL_NO_BATCHED_PRODUCT") != NULL;
	if (!fma && !dq->mul.active && nitems >= 2 && !dq->inverse
			&& (small ? dq->products.size() < PRODUCTS_MAX
				: !no_batch && ntt_indirect_product_supported(
					(unsigned) dq->log2n, mod)
				&& dq->inv_products.size() < DEFER_MAX)) {
		const defer_item &ia = dq->items[nitems - 2], &ib = dq->items[nitems - 1];
		struct vkhel_ntt_tables *ntt = dq->tables[ia.table];
		const void *rp = result->device.ptr;
		const bool operands = (ia.result == a && ib.result == b)
			|| (ia.result == b && ib.result == a);
		if (operands && a != b && dq->tables[ib.table] == ntt && ntt->q == mod
				&& ntt->n == len && a->length == len && b->length == len
				&& !dq->reads.count(rp) && !dq->writes.count(rp)
				&& !small) {
			/* two-pass sizes: the forward transforms stay in the record (a
			 * loop of products batches them), the product waits behind them */
			if (dq->inv_products.empty() || dq->inv_product_tables == ntt) {
				dq->writes.insert(rp);
				dq->mul.active = true;
				dq->mul.fma = false;
				dq->mul.with_forwards = false;
				dq->mul.after_items = true;
				dq->mul.ntt = ntt;
				dq->mul.a = a;
				dq->mul.b = b;
				dq->mul.result = result;
				dq->mul.mod = mod;
				dq->mul.multiplier = 0;
				return true;
			}
		} else if (operands && a != b && dq->tables[ib.table] == ntt && ntt->q == mod
				&& ntt->n == len && a->length == len && b->length == len
				&& !dq->reads.count(rp) && !dq->writes.count(rp)
				&& (dq->products.empty() || dq->product_tables == ntt)) {
			dq->triple.active = true;
			dq->triple.fa = ia.result == a ? ia : ib;
			dq->triple.fb = ia.result == a ? ib : ia;
			dq->triple.ntt = ntt;
			dq->items.pop_back();
			dq->items.pop_back();
			if (dq->items.empty()) {
				dq->tables.clear();
			}
			dq->writes.insert(rp);
			dq->mul.active = true;
			dq->mul.fma = false;
			dq->mul.with_forwards = true;
			dq->mul.a = a;
			dq->mul.b = b;
			dq->mul.result = result;
			dq->mul.mod = mod;
			dq->mul.multiplier = 0;
			return true;
		}
	}
	defer_flush(result->ctx);   /* what was recorded so far goes first */
	dq->mul.active = true;
	dq->mul.with_forwards = false;
	dq->mul.after_items = false;
	dq->mul.fma = fma;
	dq->mul.a = a;
	dq->mul.b = b;
	dq->mul.result = result;
	dq->mul.mod = mod;
	dq->mul.multiplier = multiplier;
	return true;
}

extern "C" void vkhel_vector_elemfma(
		const struct vkhel_vector *a, const struct vkhel_vector *b,
		struct vkhel_vector *result, uint64_t multiplier, uint64_t mod) {
	VK_REQUIRE(a->ctx == b->ctx && b->ctx == result->ctx,
			"elemfma: vectors belong to different contexts");
	VK_REQUIRE(mod >= 2, "elemfma: modulus must be at least 2");
	enter(result->ctx);
	DBG("elemfma (multiplier: %" PRIu64 " mod: %" PRIu64 ")\n",
			multiplier, mod);
	DBG_VEC("a", a);
	DBG_VEC("b", b);
	/* "modular add" (multiplier 1) or a scaled accumulation in the transform
	 * domain, followed by the in-place inverse transform: recorded like the
	 * product and folded into that transform's first load */
	if (record_pointwise(true, a, b, result, multiplier, mod)) {
		DBG_VEC("result", result);
		return;
	}
	launch_elemfma(result->ctx, dev_u64(a), dev_u64(b), dev_u64(result),
			result->length, multiplier, mod);
	DBG_VEC("result", result);
}

extern "C" void vkhel_vector_elemmul(
		const struct vkhel_vector *a, const struct vkhel_vector *b,
		struct vkhel_vector *result, uint64_t mod) {
	VK_REQUIRE(a->ctx == b->ctx && b->ctx == result->ctx,
			"elemmul: vectors belong to different contexts");
	VK_REQUIRE(mod >= 2, "elemmul: modulus must be at least 2");
	enter(result->ctx);
	DBG("elemmul mod: %" PRIu64 "\n", mod);
	DBG_VEC("a", a);
	DBG_VEC("b", b);
	if (record_pointwise(false, a, b, result, 0, mod)) {
		DBG_VEC("result", result);
		return;
	}
	launch_elemmul(result->ctx, dev_u64(a), dev_u64(b), dev_u64(result),
			result->length, mod);
	DBG_VEC("result", result);
}

extern "C" void vkhel_vector_elemmul_rns(
		const struct vkhel_vector *a, const struct vkhel_vector *b,
		struct vkhel_vector *result, const uint64_t *mods,
		uint64_t limbs, uint64_t n, uint64_t batch) {
	VK_REQUIRE(a->ctx == b->ctx && b->ctx == result->ctx,
			"elemmul_rns: vectors belong to different contexts");
	const uint64_t total = limbs * n * batch;
	VK_REQUIRE(a->length >= total && b->length >= total
			&& result->length >= total, "elemmul_rns: vector too short");
	VK_REQUIRE(limbs >= 1 && limbs <= 64, "elemmul_rns: 1..64 limbs");
	enter(result->ctx);
	launch_elemmul_rns(result->ctx, dev_u64(a), dev_u64(b), dev_u64(result),
			mods, limbs, n, batch);
}

extern "C" void vkhel_vector_elemgtadd(const struct vkhel_vector *operand,
		struct vkhel_vector *result, uint64_t bound, uint64_t diff) {
	VK_REQUIRE(operand->ctx == result->ctx,
			"elemgtadd: vectors belong to different contexts");
	enter(result->ctx);
	launch_elemgtadd(result->ctx, dev_u64(operand), dev_u64(result),
			result->length, bound, diff);
}

extern "C" void vkhel_vector_elemgtsub(const struct vkhel_vector *operand,
		struct vkhel_vector *result, uint64_t bound, uint64_t diff,
		uint64_t mod) {
	VK_REQUIRE(operand->ctx == result->ctx,
			"elemgtsub: vectors belong to different contexts");
	VK_REQUIRE(mod >= 2, "elemgtsub: modulus must be at least 2");
	enter(result->ctx);
	launch_elemgtsub(result->ctx, dev_u64(operand), dev_u64(result),
			result->length, bound, diff, mod);
}

extern "C" void vkhel_vector_elemmod(const struct vkhel_vector *operand,
		struct vkhel_vector *result, uint64_t mod, uint64_t q) {
	VK_REQUIRE(operand->ctx == result->ctx,
			"elemmod: vectors belong to different contexts");
	VK_REQUIRE(mod >= 2, "elemmod: modulus must be at least 2");
	enter(result->ctx);
	DBG("elemmod mod: %" PRIu64 ", q: %" PRIu64 "\n", mod, q);
	/* dispatch rule of the reference, src/vector.c:360-368 */
	if (mod == 2) {
		launch_elemmodbytwo(result->ctx, dev_u64(operand), dev_u64(result),
				result->length, q / 2);
	} else {
		launch_elemgtsub(result->ctx, dev_u64(operand), dev_u64(result),
				result->length, q / 2, q, mod);
	}
}

/* ---- transforms ------------------------------------------------------------------ */
static void check_ntt(const char *what, const struct vkhel_vector *operand,
		const struct vkhel_vector *result, const struct vkhel_ntt_tables *ntt,
		uint64_t count) {
	VK_REQUIRE(operand->ctx == result->ctx,
			"%s: vectors belong to different contexts", what);
	VK_REQUIRE(ntt, "%s: NULL tables", what);
	VK_REQUIRE(operand->length >= count && result->length >= count,
			"%s: vector shorter than the transform (%" PRIu64 ")", what,
			count);
}

extern "C" void vkhel_vector_forward_transform_batch(
		const struct vkhel_vector *operand, struct vkhel_vector *result,
		struct vkhel_ntt_tables *ntt, uint64_t batch) {
	check_ntt("forward_transform", operand, result, ntt, ntt->n * batch);
	struct vkhel_ctx *ctx = result->ctx;
	enter(ctx);
	if (ntt->n < 2 || batch == 0) {
		return; /* no stages: the reference leaves result untouched */
	}
	const limb_desc *desc = ntt_tables_device_desc(ctx, ntt);
	if (batch > 1 && hold_forward(operand, result, desc, 1, batch,
				(unsigned) ntt->log2n, ntt->q)) {
		return;
	}
	launch_ntt(ctx, false, dev_u64(operand), dev_u64(result), desc, 1, batch,
			(unsigned) ntt->log2n, ntt->q);
}

extern "C" void vkhel_vector_inverse_transform_batch(
		const struct vkhel_vector *operand, struct vkhel_vector *result,
		struct vkhel_ntt_tables *ntt, uint64_t batch) {
	check_ntt("inverse_transform", operand, result, ntt, ntt->n * batch);
	struct vkhel_ctx *ctx = result->ctx;
	enter(ctx);
	if (batch == 0) {
		return;
	}
	if (ntt->n < 2) {
		/* no butterfly stage; the reference still multiplies result (not
		 * operand) by n^-1 = 1, i.e. reduces it (vector.c:633-639) */
		launch_elemmulconst(ctx, dev_u64(result), dev_u64(result), batch,
				ntt->inv_n, ntt->q);
		return;
	}
	const limb_desc *desc = ntt_tables_device_desc(ctx, ntt);
	inverse_follows(operand, result, desc, batch);
	launch_ntt(ctx, true, dev_u64(operand), dev_u64(result), desc, 1, batch,
			(unsigned) ntt->log2n, ntt->q);
}

extern "C" void vkhel_vector_forward_transform(
		const struct vkhel_vector *operand, struct vkhel_vector *result,
		struct vkhel_ntt_tables *ntt) {
	DBG("forward transform (degree: %" PRIu64 " mod: %" PRIu64
			" omega: %" PRIu64 ")\n", ntt->n, ntt->q, ntt->w);
	DBG_VEC("operand", operand);
	check_ntt("forward_transform", operand, result, ntt, ntt->n);
	/* a single thread may interleave contexts on different devices: every
	 * CUDA call below (event creation, launches of a flushed record, the
	 * fused product) has to find this context's device current */
	enter(result->ctx);
	if (ntt->n >= 2 && defer_transform(false, operand, result, ntt)) {
		DBG_VEC("result", result);
		return;
	}
	vkhel_vector_forward_transform_batch(operand, result, ntt, 1);
	DBG_VEC("result", result);
}

extern "C" void vkhel_vector_inverse_transform(
		const struct vkhel_vector *operand, struct vkhel_vector *result,
		struct vkhel_ntt_tables *ntt) {
	DBG("inverse transform (degree: %" PRIu64 " mod: %" PRIu64
			" omega: %" PRIu64 ")\n", ntt->n, ntt->q, ntt->w);
	DBG_VEC("operand", operand);
	check_ntt("inverse_transform", operand, result, ntt, ntt->n);
	enter(result->ctx);
	if (fuse_product_into_inverse(operand, result, ntt)) {
		DBG_VEC("result", result);
		return;
	}
	if (ntt->n >= 2 && result->length == ntt->n
			&& defer_transform(true, operand, result, ntt)) {
		DBG_VEC("result", result);
		return;
	}
	vkhel_vector_inverse_transform_batch(operand, result, ntt, 1);
	/* The reference scales every element of result, not just the first n
	 * (vector.c:635-638 run elemmulconst over result->length).
	 * The n^-1 factor of the first n is folded into the last butterfly
	 * stage; the tail gets the same Shoup multiplication here. */
	if (result->length > ntt->n) {
		launch_elemmulconst(result->ctx, dev_u64(result) + ntt->n,
				dev_u64(result) + ntt->n, result->length - ntt->n,
				ntt->inv_n, ntt->q);
	}
	DBG_VEC("result", result);
}

/* Device pointers for a batched transform that may continue a sliced
 * transform slice by slice (device.h, split_*): fetching them does not join
 * the auxiliary stream -- unless a transfer of one of the vectors is pending,
 * which only the context's stream would wait for.  launch_ntt joins if the new
 * transform does not fit the slices that are in flight. */
static void sliced_operands(const struct vkhel_vector *operand,
		struct vkhel_vector *result, const u64 **src, u64 **dst) {
	struct vkhel_ctx *ctx = result->ctx;
	ctx->dev.split_hold = !operand->xfer_pending && !result->xfer_pending;
	*src = dev_u64(operand);
	*dst = dev_u64(result);
	ctx->dev.split_hold = 0;
}

static const limb_desc *rns_prepare(const char *what,
		const struct vkhel_vector *operand, const struct vkhel_vector *result,
		struct vkhel_ntt_tables *const *ntt, uint64_t limbs, uint64_t batch,
		uint64_t *q_max) {
	VK_REQUIRE(ntt && limbs >= 1, "%s: need at least one limb", what);
	check_ntt(what, operand, result, ntt[0], ntt[0]->n * limbs * batch);
	*q_max = 0;
	for (uint64_t l = 0; l < limbs; l++) {
		if (ntt[l]->q > *q_max) {
			*q_max = ntt[l]->q;
		}
	}
	return rns_plan_device_descs(result->ctx, ntt, limbs);
}

extern "C" void vkhel_vector_forward_transform_rns(
		const struct vkhel_vector *operand, struct vkhel_vector *result,
		struct vkhel_ntt_tables *const *ntt, uint64_t limbs, uint64_t batch) {
	uint64_t q_max;
	enter(result->ctx);
	const limb_desc *descs = rns_prepare("forward_transform_rns", operand,
			result, ntt, limbs, batch, &q_max);
	if (ntt[0]->n < 2 || batch == 0) {
		return;
	}
	if (hold_forward(operand, result, descs, limbs, limbs * batch,
				(unsigned) ntt[0]->log2n, q_max)) {
		return;
	}
	const u64 *src;
	u64 *dst;
	sliced_operands(operand, result, &src, &dst);
	launch_ntt(result->ctx, false, src, dst, descs,
			limbs, limbs * batch, (unsigned) ntt[0]->log2n, q_max);
}

extern "C" void vkhel_vector_inverse_transform_rns(
		const struct vkhel_vector *operand, struct vkhel_vector *result,
		struct vkhel_ntt_tables *const *ntt, uint64_t limbs, uint64_t batch) {
	uint64_t q_max;
	enter(result->ctx);
	const limb_desc *descs = rns_prepare("inverse_transform_rns", operand,
			result, ntt, limbs, batch, &q_max);
	VK_REQUIRE(ntt[0]->n >= 2, "inverse_transform_rns: n must be at least 2");
	if (batch == 0) {
		return;
	}
	inverse_follows(operand, result, descs, limbs * batch);
	const u64 *src;
	u64 *dst;
	sliced_operands(operand, result, &src, &dst);
	launch_ntt(result->ctx, true, src, dst, descs,
			limbs, limbs * batch, (unsigned) ntt[0]->log2n, q_max);
}

extern "C" void vkhel_vector_polymul_rns(
		const struct vkhel_vector *a, const struct vkhel_vector *b,
		struct vkhel_vector *result,
		struct vkhel_ntt_tables *const *ntt, uint64_t limbs, uint64_t batch) {
	uint64_t q_max;
	struct vkhel_ctx *ctx = result->ctx;
	enter(ctx);
	VK_REQUIRE(a->ctx == ctx && b->ctx == ctx,
			"polymul_rns: vectors belong to different contexts");
	const limb_desc *descs = rns_prepare("polymul_rns", a, result, ntt, limbs,
			batch, &q_max);
	const uint64_t n = ntt[0]->n;
	const uint64_t total = n * limbs * batch;
	VK_REQUIRE(b->length >= total, "polymul_rns: vector b too short");
	VK_REQUIRE(n >= 2 && limbs <= 64, "polymul_rns: n >= 2, at most 64 limbs");
	if (batch == 0) {
		return;
	}
	u64 *tmp = (u64 *) device_scratch(ctx, total * sizeof(u64));
	const unsigned log2n = (unsigned) ntt[0]->log2n;
	const u64 *pa = dev_u64(a), *pb = dev_u64(b);
	u64 *pr = dev_u64(result);
	/* fast path: strided passes of NTT(b) -> scratch and NTT(a) -> result,
	 * then one kernel for the row passes of both, the product and the row
	 * pass of the inverse, then the strided passes of the inverse */
	if (launch_ntt_polymul(ctx, pa, pb, tmp, pr, descs, limbs, limbs * batch,
				log2n, q_max)) {
		return;
	}
	/* otherwise the reference API sequence, NTT(b) first so that result may
	 * alias either operand: NTT(b) -> scratch, NTT(a) -> result, product
	 * (folded into the first pass of the inverse where that kernel applies),
	 * inverse */
	uint64_t mods[64];
	for (uint64_t l = 0; l < limbs; l++) {
		mods[l] = ntt[l]->q;
	}
	launch_ntt(ctx, false, pb, tmp, descs, limbs, limbs * batch, log2n, q_max);
	launch_ntt(ctx, false, pa, pr, descs, limbs, limbs * batch, log2n, q_max);
	if (!launch_ntt_inverse_of_product(ctx, pr, tmp, pr, descs, limbs,
				limbs * batch, log2n, q_max)) {
		launch_elemmul_rns(ctx, pr, tmp, pr, mods, limbs, n, batch);
		launch_ntt(ctx, true, pr, pr, descs, limbs, limbs * batch, log2n,
				q_max);
	}
}
