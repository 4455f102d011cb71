#!/usr/bin/env python3
"""Shared-memory bank model of the row pass's exchanges (K = 8, one tile per
warp): wavefronts per warp instruction for every access pattern of
ntt_rows_kernel under the shipped padding xpad(i) = i + 4*(i >> 5) and under
the alternative i + 2*(i >> 4) (same buffer size).  A wavefront serves 128
bytes, one 4-byte word per bank; a 64-bit access of 32 lanes needs at least 2,
a 128-bit access at least 4.

Result: the 128-bit accesses of the deepest round's layout (a lane's four
adjacent coefficients, i.e. 16 bytes at a 32-byte lane stride) take 8
wavefronts with the shipped padding and 4 with the alternative; everything
else is already minimal.  ncu counts 6.0 M conflict wavefronts out of 21.9 M in
the row kernels (profiles/r01_ntt_ncu_full.txt); the shared-memory pipe is at
18-20 %, so this is latency, not throughput: built with -DROWS_XPAD_HALF=1 the
bench step is 0.6630 ms against 0.6638 ms (forward 0.5 % slower, inverse 0.3 %
faster) -- no gain, the shipped padding stays."""
K = 8


def tile_index(r, t, e):
    """tile_geom<8>::index"""
    if r < 2:
        p = K - 3 * (r + 1)
        return ((t >> p) << (p + 3)) | (e << p) | (t & ((1 << p) - 1))
    return ((e >> 2) << 7) | (t << 2) | (e & 3)


def wavefronts(word_addrs, words_per_lane):
    """minimum number of 128-byte wavefronts for one warp instruction: lanes are
    served in groups as large as the access width allows (32, 16 or 8 lanes),
    and a group needs as many wavefronts as its most loaded bank"""
    lanes_per_group = 32 // (words_per_lane * 2)
    total = 0
    for g in range(0, 32, lanes_per_group):
        load = {}
        for lane in range(g, g + lanes_per_group):
            for w in range(words_per_lane):
                for half in range(2):          # a 64-bit word is two banks
                    bank = (2 * (word_addrs[lane] + w) + half) % 32
                    load.setdefault(bank, set()).add(word_addrs[lane] + w)
        total += max(len(v) for v in load.values())
    return total


def report(name, pad):
    print(name)
    for r, width in ((0, 1), (1, 1), (2, 2)):
        regs = range(0, 8, width)
        counts = [wavefronts([pad(tile_index(r, t, e)) for t in range(32)], width)
                  for e in regs]
        print("  round-%d layout, %3d-bit accesses: %s wavefronts per "
              "instruction (minimum %d)" % (r, 64 * width, sorted(set(counts)),
                                            2 * width))


def main():
    report("shipped      xpad(i) = i + 4*(i >> 5)", lambda i: i + ((i >> 5) << 2))
    report("alternative  xpad(i) = i + 2*(i >> 4)", lambda i: i + ((i >> 4) << 1))


if __name__ == "__main__":
    main()
