"""Work-division selection pinned to the reference's own device-independent known-answer test
(test/unit/workDiv/src/WorkDivHelpersTest.cpp:34-180, "subDivideGridElems.2D.examples": device properties overridden,
so the expected work divisions do not depend on the machine). CPU only: the logic lives behind the C ABI
(b200_subdivide_grid_elems / b200_is_valid_work_div, alpaka_b200/csrc/b200_workdiv.cpp) and needs no device."""
import pytest

from alpaka_b200 import workdiv as wd

E, C, U = wd.EQUAL_EXTENT, wd.CLOSE_TO_EQUAL_EXTENT, wd.UNRESTRICTED


def props_2d():
    # WorkDivHelpersTest.cpp:27-32
    return wd.make_props(2, grid_block_extent_max=(1024, 1024), grid_block_count_max=1024 * 1024,
                         block_thread_extent_max=(256, 128), block_thread_count_max=512, thread_elem_extent_max=(8, 8),
                         thread_elem_count_max=16)


# (gridElemExtent, kernelBlockThreadCountMax, mustDivide, restriction) -> (gridBlockExtent, blockThreadExtent)
GOLDEN = [
    ((300, 600), 0, False, E, (14, 28), (22, 22)),  # :34-45
    ((300, 600), 0, False, C, (19, 19), (16, 32)),  # :46-54
    ((300, 600), 0, False, U, (75, 5), (4, 128)),  # :55-63
    ((300, 600), 0, True, E, (1, 2), (256, 128)),  # :65-74
    ((300, 600), 0, True, C, (20, 20), (15, 30)),  # :75-83
    ((300, 600), 0, True, U, (75, 5), (4, 120)),  # :84-92
    ((300, 600), 256, True, U, (150, 5), (2, 120)),  # :99-110
    ((300, 600), 256, True, C, (20, 40), (15, 15)),  # :111-119
    ((300, 600), 256, False, E, (19, 38), (16, 16)),  # :123-131
    ((300, 600), 256, False, U, (150, 5), (2, 128)),  # :132-140
    ((300, 600), 256, False, C, (19, 38), (16, 16)),  # :141-149
    ((1000, 600), 256, False, E, (63, 38), (16, 16)),  # :152-160
    ((1000, 600), 256, False, U, (500, 5), (2, 128)),  # :161-169
    ((1000, 600), 256, False, C, (63, 38), (16, 16)),  # :170-178
]


@pytest.mark.parametrize("extent,kmax,must_divide,restriction,grid,block", GOLDEN)
def test_subdivide_grid_elems_reference_golden_vectors(extent, kmax, must_divide, restriction, grid, block):
    got = wd.sub_divide_grid_elems(extent, (1, 1), props_2d(), kmax, must_divide, restriction)
    assert got.grid_block_extent == grid
    assert got.block_thread_extent == block
    assert got.thread_elem_extent == (1, 1)
    # every answer is itself a valid work division for those properties -- except the one case where the reference
    # returns the per-axis maxima (256 x 128 = 32768 threads > 512; its own comment at :73 says so)
    if block != (256, 128):
        assert wd.is_valid_work_div(got, props_2d(), kmax)


def test_babelstream_shape_on_cuda_like_limits():
    """1-D power-of-two extents on CUDA limits give {N/1024, 1024, 1} (benchmarks/babelstream/src/README.md:47)."""
    p = wd.make_props(1, grid_block_extent_max=(2**31 - 1,), grid_block_count_max=2**31 - 1, block_thread_extent_max=(1024,),
                      block_thread_count_max=1024, thread_elem_extent_max=(2**31 - 1,), thread_elem_count_max=2**31 - 1)
    for n in (1 << 25, 1 << 30):
        got = wd.sub_divide_grid_elems((n,), (1,), p, 1024, True, U)
        assert got.grid_block_extent == (n // 1024,) and got.block_thread_extent == (1024,)


def test_prime_extent_with_must_divide_degenerates_to_one_thread_blocks():
    """KernelCfg note in workdiv/WorkDivHelpers.hpp:326-329: a prime grid thread extent forces block extent 1."""
    p = wd.make_props(1, grid_block_extent_max=(2**31 - 1,), grid_block_count_max=2**31 - 1, block_thread_extent_max=(1024,),
                      block_thread_count_max=1024, thread_elem_extent_max=(2**31 - 1,), thread_elem_count_max=2**31 - 1)
    got = wd.sub_divide_grid_elems((1000003,), (1,), p, 1024, True, U)
    assert got.block_thread_extent == (1,) and got.grid_block_extent == (1000003,)
    got = wd.sub_divide_grid_elems((1000003,), (1,), p, 1024, False, U)
    assert got.block_thread_extent == (1024,) and got.grid_block_extent == (977,)


def test_is_valid_work_div_rejects_violations():
    p = props_2d()
    ok = wd.WorkDivMembers((4, 4), (16, 16), (1, 1))
    assert wd.is_valid_work_div(ok, p)
    assert not wd.is_valid_work_div(wd.WorkDivMembers((4, 4), (512, 1), (1, 1)), p)  # block extent > max in dim 0
    assert not wd.is_valid_work_div(wd.WorkDivMembers((4, 4), (32, 32), (1, 1)), p)  # 1024 threads > 512
    assert not wd.is_valid_work_div(wd.WorkDivMembers((2048, 1), (1, 1), (1, 1)), p)  # grid extent > max
    assert not wd.is_valid_work_div(wd.WorkDivMembers((4, 4), (16, 16), (8, 8)), p)  # 64 elems > 16
    assert not wd.is_valid_work_div(wd.WorkDivMembers((0, 4), (16, 16), (1, 1)), p)  # zero extent
    assert not wd.is_valid_work_div(ok, p, kernel_block_thread_count_max=128)  # kernel limit 128 < 256
