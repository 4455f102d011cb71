/*
 * Host-side copy into staging memory with a few helper threads.
 *
 * vkhel_vector_copy_from_host (reference src/vector.c:262-268) takes pageable
 * memory that the caller may reuse as soon as the call returns, so the data is
 * copied into a pinned staging buffer before the DMA starts.  From C on the
 * B200 box that memcpy is what a loop of uploads costs (examples/api_e2e.c:
 * 44 us per 512 KiB polynomial on one core, the DMA itself 11 us), so copies of
 * 128 KiB and more are cut into 64 KiB parts that the caller and two helper
 * threads take from a shared counter (more helpers stop paying in that loop:
 * 25 k NTT/s with 3 threads per copy, 22 k with 4, 17 k with 1 --
 * profiles/r02_hostcopy.txt).
 *
 * The helpers are started on the first large copy, spin for a short while after
 * a job (a loop of uploads finds them awake) and then sleep on a condition
 * variable.  One job at a time: a caller that finds the pool taken (another
 * thread of the application is inside a copy) copies on its own.
 * $VKHEL_COPY_THREADS sets the number of threads per copy including the
 * caller (default 3, 1 = plain memcpy, at most 8; never more than the CPUs the
 * process is allowed on).  The helpers block every signal, and libvkhel.so is
 * linked with -z nodelete: a dlclose() must not unmap code they are running.
 */
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

#include <pthread.h>
#include <sched.h>
#include <signal.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "common.cuh"

#define COPY_PART_BYTES ((size_t) 64 << 10)
#define COPY_MIN_BYTES ((size_t) 128 << 10)
#define COPY_MAX_THREADS 8
#define COPY_SPIN_NS 200000   /* helpers stay awake this long after a job */

namespace {

/* `state` = generation << 32 | next part.  The generation is odd while a job
 * is open; parts are claimed by compare-and-swap on the whole word, so a
 * helper that is late for job G can never claim a part of a later job, and the
 * job's fields (written only while the generation is even) are stable for
 * whoever holds a claim. */
struct copy_pool {
	std::mutex owner;                 /* one job at a time */
	std::mutex m;                     /* sleepers and the wake-up */
	std::condition_variable cv;
	int sleepers = 0;
	std::atomic<uint64_t> state{0};
	char *dst = NULL;
	const char *src = NULL;
	size_t bytes = 0;
	std::atomic<size_t> parts{0};
	std::atomic<size_t> done{0};
};

copy_pool *g_pool;   /* never freed: the helpers outlive static destructors */
std::once_flag g_pool_once;

uint64_t now_ns() {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (uint64_t) ts.tv_sec * 1000000000ull + (uint64_t) ts.tv_nsec;
}

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
	__builtin_ia32_pause();
#elif defined(__aarch64__)
	asm volatile("yield");
#endif
}

inline bool open_gen(uint64_t state) {
	return (state >> 32) & 1;
}

/* take parts of job `gen` until none are left or the job has been closed */
void take_parts(copy_pool *p, uint64_t gen) {
	uint64_t cur = p->state.load(std::memory_order_acquire);
	for (;;) {
		if (cur >> 32 != gen || (cur & 0xffffffffu)
				>= p->parts.load(std::memory_order_relaxed)) {
			return;
		}
		if (!p->state.compare_exchange_weak(cur, cur + 1,
					std::memory_order_acq_rel, std::memory_order_acquire)) {
			continue;
		}
		/* the claim succeeded while job `gen` was open, and the job stays open
		 * until this part is counted: the fields are this job's */
		const size_t off = (size_t) (cur & 0xffffffffu) * COPY_PART_BYTES;
		const size_t len = p->bytes - off < COPY_PART_BYTES ? p->bytes - off
			: COPY_PART_BYTES;
		memcpy(p->dst + off, p->src + off, len);
		p->done.fetch_add(1, std::memory_order_release);
		cur = p->state.load(std::memory_order_acquire);
	}
}

void helper_main(copy_pool *p) {
	/* signals are the application's business: none is handled here */
	sigset_t all;
	sigfillset(&all);
	pthread_sigmask(SIG_BLOCK, &all, NULL);
	uint64_t seen = 0;   /* generation of the last job worked on */
	for (;;) {
		const uint64_t idle_since = now_ns();
		uint64_t cur;
		unsigned spins = 0;
		while (!open_gen(cur = p->state.load(std::memory_order_acquire))
				|| cur >> 32 == seen) {
			cpu_relax();
			if ((++spins & 255) == 0 && now_ns() - idle_since > COPY_SPIN_NS) {
				std::unique_lock<std::mutex> lock(p->m);
				p->sleepers++;
				p->cv.wait(lock, [&] {
					const uint64_t s = p->state.load(std::memory_order_acquire);
					return open_gen(s) && s >> 32 != seen;
				});
				p->sleepers--;
			}
		}
		seen = cur >> 32;
		take_parts(p, seen);
	}
}

/* threads per copy, the caller included: read once, whichever application
 * thread comes first */
int copy_threads() {
	static const int threads = [] {
		const char *env = getenv("VKHEL_COPY_THREADS");
		int n = env && *env ? atoi(env) : 3;
		/* no more threads than CPUs this process may run on (a caller pinned
		 * to one core gets the plain memcpy) */
		cpu_set_t cpus;
		int allowed = (int) std::thread::hardware_concurrency();
		if (sched_getaffinity(0, sizeof(cpus), &cpus) == 0) {
			allowed = CPU_COUNT(&cpus);
		}
		if (allowed > 0 && n > allowed) {
			n = allowed;
		}
		return n < 1 ? 1 : n > COPY_MAX_THREADS ? COPY_MAX_THREADS : n;
	}();
	return threads;
}

void pool_start() {
	g_pool = new copy_pool;
	for (int i = 0; i + 1 < copy_threads(); i++) {
		std::thread(helper_main, g_pool).detach();
	}
}

}  // namespace

extern "C" void host_copy(void *dst, const void *src, size_t bytes) {
	const int threads = copy_threads();
	if (threads == 1 || bytes < COPY_MIN_BYTES) {
		memcpy(dst, src, bytes);
		return;
	}
	std::call_once(g_pool_once, pool_start);
	copy_pool *p = g_pool;
	if (!p->owner.try_lock()) {
		memcpy(dst, src, bytes);
		return;
	}
	p->dst = (char *) dst;
	p->src = (const char *) src;
	p->bytes = bytes;
	const size_t parts = (bytes + COPY_PART_BYTES - 1) / COPY_PART_BYTES;
	p->parts.store(parts, std::memory_order_relaxed);
	p->done.store(0, std::memory_order_relaxed);
	/* open the job: the next (odd) generation, part 0 */
	const uint64_t gen =
		((p->state.load(std::memory_order_relaxed) >> 32) + 1) & 0xffffffffu;
	p->state.store(gen << 32, std::memory_order_release);
	{
		std::lock_guard<std::mutex> lock(p->m);
		if (p->sleepers) {
			p->cv.notify_all();
		}
	}
	take_parts(p, gen);
	while (p->done.load(std::memory_order_acquire) < parts) {
		cpu_relax();
	}
	/* close it (even generation) before the fields may change again: a helper
	 * that wakes up late finds nothing to claim */
	p->state.store(((gen + 1) & 0xffffffffu) << 32, std::memory_order_release);
	p->owner.unlock();
}
