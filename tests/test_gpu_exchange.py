"""Dot / reduce over several ranks with the exchange fused into the reduction launch (b200_dot_allranks_* /
b200_reduce_sum_allranks_*): every rank's last block stores its scalar into all ranks' slot arrays through peer pointers,
publishes the call number, waits for the others and folds the slots in rank order.

Here all ranks live on ONE device (one queue per rank, plain device pointers as "peer" pointers); the multi-process form
over CUDA IPC is tests/mp_worker.py. Parity definition (multi-GPU sharding): the sharded result
equals the rank-ordered combination of the per-slab results bit for bit on EVERY rank, and the unsharded oracle within
the path's tolerance (1e-12 double; exact for integers)."""
import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import P

pytestmark = pytest.mark.gpu


def _ranks(ab, dev, world):
    from alpaka_b200 import multi

    queues = [ab.Queue(dev) for _ in range(world)]
    exs = [multi.ScalarExchange(q, r, world) for r, q in enumerate(queues)]
    multi.connect_exchange_in_process(exs)
    return queues, exs


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_dot_all_ranks_is_the_rank_ordered_combination(gpu, world):
    ab, dev, _ = gpu
    from alpaka_b200 import decomp

    queues, exs = _ranks(ab, dev, world)
    n = (1 << 20) + 77
    outs = [ab.alloc_buf(dev, np.float64, 1, q) for q in queues]
    for call in range(5):  # several collective calls in a row: slots are double-buffered by the parity of the call number
        a, b = ol.fill("uniform_f64", n, seed=50 + call), ol.fill("uniform_f64", n, seed=70 + call)
        slabs, local = [], []
        for r, q in enumerate(queues):
            lo, hi = decomp.slab_bounds(n, world, r, align=4)
            da, db = ab.alloc_buf(dev, np.float64, hi - lo, q), ab.alloc_buf(dev, np.float64, hi - lo, q)
            ab.memcpy(q, da, a[lo:hi])
            ab.memcpy(q, db, b[lo:hi])
            local.append(float(ab.babelstream.dot(q, da, db)))  # the rank's own scalar, no exchange
            slabs.append((da, db))
        for r, q in enumerate(queues):  # enqueue on every rank BEFORE waiting on any: the launches wait for each other
            exs[r].dot_async(q, slabs[r][0], slabs[r][1], outs[r])
        got = []
        for r, q in enumerate(queues):
            h = np.empty(1)
            ab.memcpy(q, h, outs[r])
            q.wait()
            got.append(float(h[0]))
        want = decomp.combine_in_rank_order(local)
        assert all(g == want for g in got), (got, want)
        d_orc = float(ol.oracle().orc_dot_f64(P(a), P(b), n, 256, 1, None))
        assert abs(got[0] - d_orc) <= 1e-12 * float(np.sum(np.abs(a * b)))
        for da, db in slabs:
            da.free()
            db.free()
    for e in exs:
        assert e.status() == 0
        e.close()


@pytest.mark.parametrize("dtype", [np.uint32, np.float32, np.float64])
def test_reduce_all_ranks(gpu, dtype):
    ab, dev, _ = gpu
    from alpaka_b200 import decomp

    world = 4
    queues, exs = _ranks(ab, dev, world)
    n = (1 << 21) + 13
    x = ol.fill("hash_u32", n, seed=5) if dtype == np.uint32 else ol.fill("bernoulli_f32", n, seed=6).astype(dtype)
    bufs, outs = [], []
    for r, q in enumerate(queues):
        lo, hi = decomp.slab_bounds(n, world, r, align=8)
        d = ab.alloc_buf(dev, dtype, hi - lo, q)
        ab.memcpy(q, d, x[lo:hi])
        bufs.append(d)
        outs.append(ab.alloc_buf(dev, dtype, 1, q))
    for r, q in enumerate(queues):
        exs[r].reduce_sum_async(q, bufs[r], outs[r])
    got = []
    for r, q in enumerate(queues):
        h = np.empty(1, dtype=dtype)
        ab.memcpy(q, h, outs[r])
        q.wait()
        got.append(h[0])
    if dtype == np.uint32:
        assert all(int(g) == int(x.astype(np.uint64).sum()) % 2**32 for g in got)  # wrap-add is order-free: bit-exact
    else:
        assert all(float(g) == float(x.astype(np.float64).sum()) for g in got)  # {0,1} data: every partial sum is exact
    for e in exs:
        assert e.status() == 0
        e.close()


def test_exchange_argument_errors(gpu):
    ab, dev, queue = gpu
    from alpaka_b200 import multi

    ex = multi.ScalarExchange(queue, 0, 1)
    a = ab.alloc_buf(dev, np.float64, 16, queue)
    out = ab.alloc_buf(dev, np.float64, 1, queue)
    with pytest.raises(ab.B200Error):  # before connect()
        ex.dot_async(queue, a, a, out)
    with pytest.raises(ab.B200Error):
        ex.connect([0])  # not the own buffer
    multi.connect_exchange_in_process([ex])
    with pytest.raises(ab.B200Error):  # unsupported element type
        ex.reduce_sum_async(queue, ab.alloc_buf(dev, np.int16, 8, queue), out)
    ab.memset(queue, a, 0)
    ex.dot_async(queue, a, a, out)  # world 1: plain Dot
    queue.wait()
    ex.close()
