"""tests/golden/multi_gpu_vectors.npz (unsharded inputs + outputs of the UNMODIFIED reference, made by
tests/golden/make_golden_multi.py): the C oracle is pinned to it on CPU; on the GPU the SHARDED paths are checked against
it with every rank on one device (tests/golden_multi.py; the same checks run across processes in tests/mp_worker.py and
inside bench.py's N > 1 arm)."""
import numpy as np
import pytest

import golden_multi as gm
import oracle_lib as ol
from oracle_lib import P

G = np.load(gm.GOLDEN)


def test_oracle_streams_match_the_multi_gpu_fixture():
    for k, pick in (("triad", 2), ("nstream", 0)):
        a, b, c = (x.copy() for x in G["stream_in"])
        ol.orc_stream(k, a, b, c, scalar=2.0)
        assert (a, b, c)[pick].tobytes() == G[f"stream_{k}"].tobytes()


def test_oracle_dot_and_reduce_match_the_multi_gpu_fixture():
    a, b, _ = (np.ascontiguousarray(x) for x in G["stream_in"])
    assert ol.oracle().orc_dot_f64(P(a), P(b), a.size, 256, 1, None) == float(G["dot_uniform"][0])
    ia, ib = (np.ascontiguousarray(x) for x in G["dot_int_in"])
    assert ol.oracle().orc_dot_f64(P(ia), P(ib), ia.size, 256, 1, None) == float(G["dot_int"][0])
    for tag in ("u32", "f32"):
        x = np.ascontiguousarray(G[f"reduce_{tag}_in"])
        got = ol.orc_reduce(x, ol.oracle().orc_reduce_block_count(x.size, 1, 1), 1, iterator=0)
        assert np.array([got], dtype=x.dtype).tobytes() == G[f"reduce_{tag}"].tobytes()


def test_oracle_heat_matches_the_multi_gpu_fixture():
    dx, dy, dt, steps = G["heat_params"]
    u0 = G["heat_init"]
    ny, nx = u0.shape[0] - 2, u0.shape[1] - 2
    assert (dx, dy, dt) == ol.heat_params(ny, nx)
    mine = np.empty_like(u0)
    ol.oracle().orc_heat2d_init(P(mine), ny, nx, nx + 2, dx, dy)
    assert mine.tobytes() == u0.tobytes()
    assert ol.orc_heat_run(u0, 1, int(steps), dx, dy, dt).tobytes() == G["heat_final"].tobytes()


def test_fixture_sizes_split_over_2_4_8_ranks():
    from alpaka_b200 import decomp

    ny, nx = G["heat_final"].shape[0] - 2, G["heat_final"].shape[1] - 2
    steps = int(G["heat_params"][3])
    for world in (2, 4, 8):
        decomp.tile_for(world - 1, world, ny, nx)
        for levels in (2, 3, 4, 6, 8):
            decomp.slab_for(world - 1, world, ny, nx, levels)
            sched = decomp.launch_schedule(steps, levels, min_depth=2)
            assert sched.count(levels) >= 2, "the fixture must exercise at least two launches of every depth"
    n = G["stream_in"].shape[1]
    last = [decomp.slab_bounds(n, w, w - 1, align=4) for w in (2, 4, 8)]
    first = [decomp.slab_bounds(n, w, 0, align=4) for w in (2, 4, 8)]
    assert all(0 < hi - lo <= f[1] for (lo, hi), f in zip(last, first))
    assert any(hi - lo < f[1] for (lo, hi), f in zip(last, first)), "at least one world size must have a ragged last slab"


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_sharded_paths_match_the_reference_fixture_all_ranks_on_one_device(gpu, world):
    ab, dev, _ = gpu
    R = gm.Ranks(ab, world, {r: ab.Queue(dev) for r in range(world)})
    verdict = gm.check_all(R)
    assert verdict["triad"] == "bit-exact" and verdict["heat_slabs_4_levels"] == "bit-exact"
