// Exercises host_copy (vkhel_b200/csrc/hostcopy.cu) without a GPU: copies of
// random sizes from several application threads at once, each compared with the
// source.  Exit status 0 = every copy exact.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

extern "C" void host_copy(void *dst, const void *src, size_t bytes);

static int run(unsigned seed, int copies, size_t max_bytes) {
	std::vector<unsigned char> src(max_bytes + 64), dst(max_bytes + 64);
	uint64_t s = 0x9E3779B97F4A7C15ull * (seed + 1);
	for (int c = 0; c < copies; c++) {
		s ^= s << 13; s ^= s >> 7; s ^= s << 17;
		const size_t bytes = c % 4 == 0 ? max_bytes : (size_t) (s % (max_bytes + 1));
		const size_t so = s >> 40 & 31, dof = s >> 50 & 31;
		for (size_t i = 0; i < bytes; i += 61) {
			src[so + i] = (unsigned char) (s >> (i % 56));
		}
		if (bytes) {
			src[so + bytes - 1] = (unsigned char) c;
		}
		memset(dst.data(), 0xA5, dst.size());
		host_copy(dst.data() + dof, src.data() + so, bytes);
		if (memcmp(dst.data() + dof, src.data() + so, bytes) != 0) {
			fprintf(stderr, "thread %u copy %d: %zu bytes differ\n", seed, c, bytes);
			return 1;
		}
		// nothing written outside the destination
		if ((dof && dst[dof - 1] != 0xA5) || dst[dof + bytes] != 0xA5) {
			fprintf(stderr, "thread %u copy %d: wrote out of range\n", seed, c);
			return 1;
		}
	}
	return 0;
}

int main(int argc, char **argv) {
	const int threads = argc > 1 ? atoi(argv[1]) : 4;
	const int copies = argc > 2 ? atoi(argv[2]) : 200;
	std::vector<int> status(threads, 0);
	std::vector<std::thread> pool;
	for (int t = 0; t < threads; t++) {
		pool.emplace_back([&, t] { status[t] = run(t, copies, (size_t) 3 << 20); });
	}
	int bad = 0;
	for (int t = 0; t < threads; t++) {
		pool[t].join();
		bad |= status[t];
	}
	// and a quiet spell long enough for the helpers to fall asleep, then more
	struct timespec ts = { 0, 5000000 };
	nanosleep(&ts, NULL);
	bad |= run(99, 20, (size_t) 1 << 20);
	puts(bad ? "FAILED" : "ok");
	return bad;
}
