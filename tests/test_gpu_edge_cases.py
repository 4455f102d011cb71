"""GPU suite: edge cases of the C API beyond the reference's own tests --
degenerate sizes, aliasing, tails, asynchronous ordering, several contexts,
table lifetime, ragged RNS shapes."""
import ctypes

import numpy as np
import pytest

import oracle
import vkhel_b200 as vk
from vkhel_b200 import params
from conftest import u64, rand_mod, rand_u64
from test_gpu_parity import TablePair, run_forward, run_inverse

pytestmark = pytest.mark.gpu


def test_degenerate_transform_sizes(ctx):
    """n = 1: no butterfly stage.  The reference's forward leaves `result`
    untouched and its inverse multiplies `result` (not operand) by 1 mod q
    (src/vector.c:536-537, 633-639); n = 2: a single stage."""
    q = params.P0
    t1 = vk.NttTables(1, q, 1)
    a = ctx.from_host(u64([5, 6, 7]))
    b = ctx.from_host(u64([q + 3, 9, (1 << 64) - 1]))
    ctx.forward_transform(a, b, t1)
    assert b.to_host().tolist() == [q + 3, 9, (1 << 64) - 1]
    ctx.inverse_transform(a, b, t1)
    assert b.to_host().tolist() == [3, 9, ((1 << 64) - 1) % q]
    a.destroy(), b.destroy(), t1.destroy()

    tp = TablePair(2, q)
    x = u64([q - 1, 12345])
    assert np.array_equal(run_forward(ctx, x, tp), oracle.forward(x, tp.ora))
    assert np.array_equal(run_inverse(ctx, x, tp), oracle.inverse(x, tp.ora))
    tp.destroy()


def test_zero_batch_is_a_no_op(ctx):
    tp = TablePair(64, params.P0)
    x = np.arange(64, dtype=np.uint64)
    a, b = ctx.from_host(x), ctx.from_host(x)
    ctx.forward_transform_batch(a, b, tp.lib, 0)
    ctx.inverse_transform_batch(a, b, tp.lib, 0)
    ctx.forward_transform_rns(a, b, [tp.lib], 0)
    ctx.polymul_rns(a, a, b, [tp.lib], 0)
    assert np.array_equal(b.to_host(), x)
    a.destroy(), b.destroy(), tp.destroy()


@pytest.mark.parametrize("n,q", [(2, 113), (4, 113), (8, 113), (16, 769),
                                 (128, 769), (32, params.Q_KAT_52)])
def test_small_kat_moduli_at_their_largest_sizes(ctx, n, q):
    """the reference's tiny KAT moduli (7-, 10-, 52-bit) up to the largest n
    their 2-adicity allows"""
    tp = TablePair(n, q)
    rng = np.random.default_rng(n)
    x = rand_mod(rng, n, q)
    assert np.array_equal(run_forward(ctx, x, tp, True), oracle.forward(x, tp.ora))
    assert np.array_equal(run_inverse(ctx, x, tp, True), oracle.inverse(x, tp.ora))
    tp.destroy()


def test_forward_leaves_tail_and_operand_alone(ctx):
    n, q = 512, params.Q61
    tp = TablePair(n, q)
    rng = np.random.default_rng(3)
    op = np.concatenate([rand_mod(rng, n, q), rand_u64(rng, 100)])
    res0 = rand_u64(rng, n + 300)
    a, b = ctx.from_host(op), ctx.from_host(res0)
    ctx.forward_transform(a, b, tp.lib)
    got = b.to_host()
    assert np.array_equal(got[:n], oracle.forward(op[:n], tp.ora))
    assert np.array_equal(got[n:], res0[n:])
    assert np.array_equal(a.to_host(), op)
    a.destroy(), b.destroy(), tp.destroy()


def test_elementwise_in_place_and_odd_lengths(ctx):
    q = params.P0
    rng = np.random.default_rng(4)
    for length in (1, 2, 3, 255, 1025):
        a, b = rand_u64(rng, length), rand_u64(rng, length)
        va, vb = ctx.from_host(a), ctx.from_host(b)
        ctx.elemfma(va, vb, vb, 77, q)          # result aliases b
        want = oracle.elemfma(a, b, 77, q)
        assert np.array_equal(vb.to_host(), want)
        ctx.elemmul(va, va, va, q)              # everything aliases
        assert np.array_equal(va.to_host(), oracle.elemmul(a, a, q))
        ctx.elemgtadd(vb, vb, q // 2, 5)
        assert np.array_equal(vb.to_host(), oracle.elemgtadd(want, q // 2, 5))
        ctx.elemgtsub(va, va, 3, 4, 1000003)
        ctx.elemmod(vb, vb, 2, q)
        va.destroy(), vb.destroy()
    empty = ctx.vector(0)
    ctx.elemmul(empty, empty, empty, q)
    assert empty.to_host().size == 0
    empty.destroy()


def test_asynchronous_ordering_without_syncs(ctx):
    """many dependent operations enqueued back to back; only the final map
    synchronises (the reference waits after every single op)"""
    n, q = 1024, params.P0
    tp = TablePair(n, q)
    rng = np.random.default_rng(5)
    x = rand_mod(rng, n, q)
    v = ctx.from_host(x)
    w = ctx.vector(n)
    for _ in range(25):
        ctx.forward_transform(v, w, tp.lib)
        ctx.elemmul(w, w, w, q)
        ctx.inverse_transform(w, v, tp.lib)
    want = x
    for _ in range(25):
        f = oracle.forward(want, tp.ora)
        want = oracle.inverse(oracle.elemmul(f, f, q), tp.ora)
    assert np.array_equal(v.to_host(), want)
    v.destroy(), w.destroy(), tp.destroy()


def test_two_contexts_share_tables(ctx):
    """tables belong to no context (reference src/ntt_tables.c:65-87); two
    contexts on the same device use the same tables object"""
    other = vk.Context(0)
    tp = TablePair(2048, params.P0)
    rng = np.random.default_rng(6)
    x, y = rand_mod(rng, 2048, tp.q), rand_mod(rng, 2048, tp.q)
    a, b = ctx.from_host(x), other.from_host(y)
    ctx.forward_transform(a, a, tp.lib)
    other.forward_transform(b, b, tp.lib)
    assert np.array_equal(a.to_host(), oracle.forward(x, tp.ora))
    assert np.array_equal(b.to_host(), oracle.forward(y, tp.ora))
    a.destroy(), b.destroy()
    other.destroy()
    tp.destroy()


def _second_device():
    """another GPU when the box has one, else the same device: the calls and
    the expected results are the same either way"""
    return 1 if vk.device_count() > 1 else 0


def test_interleaved_contexts_on_one_thread(ctx):
    """One host thread drives two contexts (two GPUs where there are two)
    call by call: every entry point has to make its own context's device
    current -- including the recorded transforms, the staging events of a
    record longer than the inline pointer table (> 8 transforms) and the
    product fused into the inverse transform."""
    other = vk.Context(_second_device())
    n, q = 1 << 12, params.P0
    tp = TablePair(n, q)
    rng = np.random.default_rng(61)
    x, y, z = (rand_mod(rng, n, q) for _ in range(3))
    a0, b0, c0 = ctx.from_host(x), ctx.from_host(y), ctx.vector(n)
    a1 = other.from_host(z)
    # elemmul recorded on ctx, an operation on the other context in between,
    # then the inverse that fuses with the product
    ctx.elemmul(a0, b0, c0, q)
    other.elemmul(a1, a1, a1, q)
    ctx.inverse_transform(c0, c0, tp.lib)
    other.forward_transform(a1, a1, tp.lib)
    want_c = oracle.inverse(oracle.elemmul(x, y, q), tp.ora)
    want_a1 = oracle.forward(oracle.elemmul(z, z, q), tp.ora)
    assert np.array_equal(c0.to_host(), want_c)
    assert np.array_equal(a1.to_host(), want_a1)
    # a record of more than 8 transforms per context, the two interleaved
    xs = [rand_mod(rng, n, q) for _ in range(24)]
    v0 = [ctx.from_host(v) for v in xs[:12]]
    v1 = [other.from_host(v) for v in xs[12:]]
    for p, r in zip(v0, v1):
        ctx.forward_transform(p, p, tp.lib)
        other.inverse_transform(r, r, tp.lib)
    for v, data in zip(v0, xs[:12]):
        assert np.array_equal(v.to_host(), oracle.forward(data, tp.ora))
    for v, data in zip(v1, xs[12:]):
        assert np.array_equal(v.to_host(), oracle.inverse(data, tp.ora))
    for v in [a0, b0, c0, a1] + v0 + v1:
        v.destroy()
    other.destroy()
    tp.destroy()


def test_copy_peer_between_contexts(ctx):
    """vkhel_vector_copy_peer gathers shards that live in another context
    (over NVLink when the contexts sit on different GPUs): ordered after the
    source's pending work and before the destination's next use."""
    other = vk.Context(_second_device())
    n, q = 1 << 13, params.P0
    tp = TablePair(n, q)
    rng = np.random.default_rng(62)
    x, y = rand_mod(rng, n, q), rand_mod(rng, n, q)
    local, remote = ctx.from_host(x), other.from_host(y)
    ctx.forward_transform(local, local, tp.lib)
    other.forward_transform(remote, remote, tp.lib)     # still only recorded
    gathered = ctx.vector(2 * n + 5)
    gathered.copy_peer(local, dst_offset=0)
    gathered.copy_peer(remote, dst_offset=n)
    gathered.copy_peer(remote, dst_offset=2 * n, src_offset=7, count=5)
    # the source may be overwritten right away: the copy has been ordered
    other.inverse_transform(remote, remote, tp.lib)
    fy = oracle.forward(y, tp.ora)
    got = gathered.to_host()
    assert np.array_equal(got[:n], oracle.forward(x, tp.ora))
    assert np.array_equal(got[n:2 * n], fy)
    assert np.array_equal(got[2 * n:], fy[7:12])
    assert np.array_equal(remote.to_host(), y)
    # and the gathered vector is an ordinary vector of its context
    ctx.inverse_transform_batch(gathered, gathered, tp.lib, 2)
    assert np.array_equal(gathered.to_host()[:2 * n], np.concatenate([x, y]))
    for v in (local, remote, gathered):
        v.destroy()
    other.destroy()
    tp.destroy()


def test_exposed_stream_stops_recording():
    """Once vkhel_ctx_stream() or vkhel_vector_device_ptr() has handed out a
    handle, the caller orders its own work by stream position, so every later
    call must be enqueued when it returns: nothing is recorded any more."""
    own = vk.Context(0)
    n, q = 1 << 10, params.P0
    tp = TablePair(n, q)
    rng = np.random.default_rng(63)
    xs = [rand_mod(rng, n, q) for _ in range(6)]
    vs = [own.from_host(v) for v in xs]
    # before: a loop of transforms goes out as one recorded batch
    for v in vs:
        own.forward_transform(v, v, tp.lib)
    own.sync()
    assert own.deferred_stats == (1, 6)
    # a vector whose pointer is out: transforms that touch it launch at once
    assert vs[0].device_ptr != 0
    l0 = own.launch_count
    own.inverse_transform(vs[0], vs[0], tp.lib)
    assert own.launch_count_noflush > l0
    # the stream is out: nothing at all is recorded from now on
    assert own.stream != 0
    l0 = own.launch_count
    for v in vs[1:]:
        own.inverse_transform(v, v, tp.lib)
        assert own.launch_count_noflush > l0
        l0 = own.launch_count_noflush
    own.elemmul(vs[0], vs[1], vs[2], q)
    assert own.launch_count_noflush > l0
    assert own.deferred_stats == (1, 6) and own.fused_products == 0
    # ... and a sliced transform joins its two streams before it returns: a
    # caller's work enqueued on the exposed stream right behind it (here a
    # plain cudaMemcpyAsync through cuda-python) sees the complete result
    from cuda.bindings import runtime as rt
    n2, limbs, batch = 1 << 14, 4, 16
    primes = params.ntt_primes(limbs)
    tps = [TablePair(n2, p) for p in primes]
    x = np.concatenate([rand_mod(rng, n2, primes[i % limbs])
                        for i in range(limbs * batch)])
    big = own.from_host(x)
    ptr = big.device_ptr
    host = np.empty_like(x)
    own.forward_transform_rns(big, big, [t.lib for t in tps], batch)
    err, = rt.cudaMemcpyAsync(host.ctypes.data, ptr, x.nbytes,
                              rt.cudaMemcpyKind.cudaMemcpyDeviceToHost,
                              own.stream)
    assert int(err) == 0
    err, = rt.cudaStreamSynchronize(own.stream)
    assert int(err) == 0
    assert np.array_equal(host, oracle.forward_batch(
        x, [t.ora for t in tps], threads=8))
    big.destroy()
    for t in tps:
        t.destroy()
    assert np.array_equal(vs[2].to_host(), oracle.elemmul(xs[0], xs[1], q))
    assert np.array_equal(vs[3].to_host(), xs[3])
    for v in vs:
        v.destroy()
    own.destroy()
    tp.destroy()


def test_table_lifetime_and_rns_plan_cache(ctx):
    """destroying and re-creating tables (possibly at the same address) must
    not resurrect a cached RNS plan: plans are keyed by table serials"""
    n = 256
    rng = np.random.default_rng(7)
    for round_ in range(4):
        primes = params.ntt_primes(3 + round_ % 2)[round_ % 2:]
        tps = [TablePair(n, q) for q in primes]
        limbs = len(primes)
        x = np.concatenate([rand_mod(rng, n, primes[p % limbs])
                            for p in range(limbs * 2)])
        v = ctx.from_host(x)
        ctx.forward_transform_rns(v, v, [t.lib for t in tps], 2)
        assert np.array_equal(
            v.to_host(), oracle.forward_batch(x, [t.ora for t in tps], 4))
        v.destroy()
        for t in tps:
            t.destroy()


@pytest.mark.parametrize("log2n,limbs,batch", [(3, 5, 7), (7, 3, 11),
                                                (9, 2, 9), (10, 7, 3),
                                                (15, 2, 3)])
def test_ragged_rns_shapes(ctx, log2n, limbs, batch):
    n = 1 << log2n
    primes = params.ntt_primes(limbs)
    tps = [TablePair(n, q) for q in primes]
    rng = np.random.default_rng(limbs * batch)
    x = np.concatenate([rand_mod(rng, n, primes[p % limbs])
                        for p in range(limbs * batch)])
    a, b = ctx.from_host(x), ctx.vector(x.size + 13)
    ctx.forward_transform_rns(a, b, [t.lib for t in tps], batch)
    want = oracle.forward_batch(x, [t.ora for t in tps], threads=8)
    assert np.array_equal(b.to_host()[:x.size], want)
    ctx.inverse_transform_rns(b, b, [t.lib for t in tps], batch)
    assert np.array_equal(b.to_host()[:x.size], x)
    a.destroy(), b.destroy()
    for t in tps:
        t.destroy()


def test_map_range_moves_only_its_range(ctx):
    """vkhel_vector_map_range stages [offset, offset + count) and unmap
    writes exactly that range back; the reference's map (the whole vector)
    is the special case"""
    data = np.arange(1000, dtype=np.uint64) * np.uint64(3)
    v = ctx.from_host(data)
    view = v.map_range(100, 50)
    assert np.array_equal(view, data[100:150])
    view[:] = np.arange(50, dtype=np.uint64) + np.uint64(7000)
    v.unmap()
    want = data.copy()
    want[100:150] = np.arange(50, dtype=np.uint64) + np.uint64(7000)
    assert np.array_equal(v.to_host(), want)
    assert v.map_range(1000, 0).size == 0      # empty range at the end
    v.unmap()
    assert np.array_equal(v.map_range(0, 1000), want)
    v.unmap()
    v.destroy()


def test_map_readahead_follows_a_recorded_batch():
    """A loop of maps after a loop of (recorded) transforms: mapping one
    result starts the device -> host copies of the next ones; every map still
    returns the current contents -- a copy started before the vector was
    modified again is not used."""
    own = vk.Context(0)
    n, q = 1 << 12, params.P0
    tp = TablePair(n, q)
    rng = np.random.default_rng(64)
    xs = [rand_mod(rng, n, q) for _ in range(10)]
    vs = [own.from_host(x) for x in xs]
    for v in vs:
        own.forward_transform(v, v, tp.lib)
    want = [oracle.forward(x, tp.ora) for x in xs]
    hits0 = own.readahead_hits
    assert np.array_equal(vs[0].to_host(), want[0])     # starts copies of 1, 2, 3
    # vector 2 changes after its copy was started: the copy is stale
    own.elemmul(vs[2], vs[2], vs[2], q)
    want[2] = oracle.elemmul(want[2], want[2], q)
    # vector 3 is overwritten from the host through the async upload path
    pinned = vk.host_alloc(n)
    pinned.array[:] = xs[3]
    vs[3].upload(pinned)
    want[3] = xs[3]
    for v, w in zip(vs[1:], want[1:]):
        assert np.array_equal(v.to_host(), w)
    assert own.readahead_hits - hits0 >= 5
    # mapping again (nothing recorded since) and in reverse order still works
    for v, w in zip(reversed(vs), reversed(want)):
        assert np.array_equal(v.to_host(), w)
    # a vector destroyed while its copy is in flight
    for v in vs:
        own.inverse_transform(v, v, tp.lib)
    got = vs[0].to_host()
    vs[1].destroy()
    assert np.array_equal(vs[2].to_host(), oracle.inverse(want[2], tp.ora))
    assert np.array_equal(got, xs[0])
    for v in [vs[0]] + vs[2:]:
        v.destroy()
    pinned.free()
    own.destroy()
    tp.destroy()


def test_map_sees_pending_upload_and_kernels(ctx):
    n = 4096
    tp = TablePair(n, params.P0)
    host = vk.host_alloc(n)
    host.array[:] = np.arange(n, dtype=np.uint64)
    v = ctx.vector(n, zero=False)
    v.upload(host)                              # copy stream
    ctx.forward_transform(v, v, tp.lib)         # compute stream, waits for it
    got = v.to_host()                           # map: waits for everything
    assert np.array_equal(got, oracle.forward(host.array, tp.ora))
    v.destroy(), tp.destroy()
    host.free()


def test_sliced_pipeline_orders_transfers_per_vector(ctx):
    """the end-to-end pattern of bench.py: slices rotate through upload ->
    transforms -> download with no synchronisation in between.  A transfer
    waits for the last kernel that touched ITS vector (and the previous
    transfer of it), not for the whole compute stream -- the result must
    still be exact when buffers are reused many times."""
    n, batch, nslices, rounds = 1 << 12, 8, 3, 6
    tp = TablePair(n, params.P0)
    count = n * batch
    rng = np.random.default_rng(77)
    hin = vk.host_alloc(count * nslices * rounds)
    hout = vk.host_alloc(count * nslices * rounds)
    hin.array[:] = rand_mod(rng, hin.array.size, tp.q)
    hout.array[:] = 0
    slices = [ctx.vector(count, zero=False) for _ in range(nslices)]
    big = ctx.vector(1 << 22, zero=False)       # unrelated long-running work
    tbig = TablePair(1 << 16, params.P0)
    for r in range(rounds):
        for c, v in enumerate(slices):
            off = (r * nslices + c) * count
            v.upload(hin, count=count, host_offset=off)
            ctx.forward_transform_batch(v, v, tp.lib, batch)
            ctx.forward_transform_batch(big, big, tbig.lib, 64)
            v.download(hout, count=count, host_offset=off)
    ctx.sync()
    want = oracle.forward_batch(hin.array, [tp.ora], threads=8)
    assert np.array_equal(hout.array, want)
    for v in slices + [big]:
        v.destroy()
    tp.destroy(), tbig.destroy()
    hin.free(), hout.free()


def test_dbgprint_symbols(ctx, capfd):
    lib = vk.lib()
    v = ctx.from_host(u64([1, 2, 3]))
    lib.vkhel_vector_dbgprint(v.handle)
    t = vk.NttTables(4, 113, 18)
    lib.vkhel_ntt_tables_dbgprint(t.handle)
    out = capfd.readouterr().out
    assert "1, 2, 3" in out
    assert "ntt_tables: (n=4 q=113 w=18)" in out and "1, 98, 18, 69" in out
    v.destroy(), t.destroy()


def test_sliced_transforms_chain_without_joins_and_join_when_needed(ctx):
    """An RNS batch of at least 8 MiB is transformed as limb slices on two
    streams, and the join of the two streams is left to whoever needs the
    context next (lazy join, kernels_ntt.cu): a chain of transforms over the
    same vector continues slice by slice, everything else -- another operation
    on the vector, a transfer, a recorded single-vector transform, another
    partition, another destination -- has to see all slices finished."""
    n, limbs, batch = 1 << 14, 4, 16                       # 8 MiB
    primes = params.ntt_primes(limbs)
    tps = [TablePair(n, q) for q in primes]
    libs, oras = [t.lib for t in tps], [t.ora for t in tps]
    rng = np.random.default_rng(88)
    x = np.concatenate([rand_mod(rng, n, primes[p % limbs])
                        for p in range(limbs * batch)])
    fx = oracle.forward_batch(x, oras, threads=8)
    data, work = ctx.from_host(x), ctx.vector(x.size, zero=False)
    # a chain: forward out of place, inverse / forward in place, several times
    for _ in range(3):
        ctx.forward_transform_rns(data, work, libs, batch)
        ctx.inverse_transform_rns(work, work, libs, batch)
    ctx.forward_transform_rns(work, work, libs, batch)
    assert np.array_equal(work.to_host(), fx)              # map joins
    # an element-wise operation right behind the slices
    ctx.inverse_transform_rns(work, work, libs, batch)
    ctx.elemgtadd(work, work, 0, 0)                        # identity, must join
    assert np.array_equal(work.to_host(), x)
    # another destination, then another partition of the same vector
    other = ctx.vector(x.size, zero=False)
    ctx.forward_transform_rns(work, work, libs, batch)
    ctx.inverse_transform_rns(work, other, libs, batch)
    assert np.array_equal(other.to_host(), x)
    ctx.forward_transform_rns(data, work, libs, batch)
    ctx.inverse_transform_rns(work, work, libs[:2] * 2, batch)   # not the tables it was made with
    ctx.forward_transform_rns(work, work, libs[:2] * 2, batch)
    assert np.array_equal(work.to_host(), fx)
    ctx.inverse_transform_rns(work, work, libs, batch)
    ctx.forward_transform_rns(work, work, [libs[0]], 1)    # one polynomial of it
    got = work.to_host()
    assert np.array_equal(got[:n], oracle.forward(x[:n], oras[0]))
    assert np.array_equal(got[n:], x[n:])
    # a RECORDED operation on the same vector between two sliced transforms:
    # it is launched (on the context's stream) when the next call fetches its
    # pointers, and must find every slice finished
    ctx.inverse_transform_rns(work, work, [libs[0]], 1)
    ones = ctx.from_host(np.ones(x.size, np.uint64))
    f0 = ctx.fused_products
    ctx.forward_transform_rns(data, work, libs, batch)     # slices in flight
    ctx.elemmul(work, ones, work, params.Q61)              # recorded; identity
    ctx.inverse_transform_rns(work, work, libs, batch)     # launches it first
    assert np.array_equal(work.to_host(), x)
    assert ctx.fused_products == f0
    ones.destroy()
    # transfers: a download behind the slices, an upload in front of them
    pinned, back = vk.host_alloc(x.size), vk.host_alloc(x.size)
    pinned.array[:] = x
    ctx.forward_transform_rns(data, work, libs, batch)
    work.download(back)
    ctx.sync()
    assert np.array_equal(back.array, fx)
    work.upload(pinned)
    ctx.forward_transform_rns(work, work, libs, batch)
    ctx.inverse_transform_rns(work, work, libs, batch)
    work.download(back)
    ctx.sync()
    assert np.array_equal(back.array, x)
    for v in (data, work, other):
        v.destroy()
    for t in tps:
        t.destroy()
    pinned.free(), back.free()
