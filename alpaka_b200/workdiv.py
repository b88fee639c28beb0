"""Work-division selection (host logic) through the C ABI: alpaka::subDivideGridElems / getValidWorkDiv /
isValidWorkDiv (reference: include/alpaka/workdiv/WorkDivHelpers.hpp:133-396, 406-549)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Sequence

from . import _lib
from ._lib import AccDevProps, check

EQUAL_EXTENT, CLOSE_TO_EQUAL_EXTENT, UNRESTRICTED = 0, 1, 2
U64_MAX = 2**64 - 1


@dataclass(frozen=True)
class WorkDivMembers:
    """alpaka::WorkDivMembers (workdiv/WorkDivMembers.hpp:18-97); vectors in alpaka order (slowest first)."""

    grid_block_extent: tuple
    block_thread_extent: tuple
    thread_elem_extent: tuple


def make_props(dim: int, *, grid_block_extent_max, grid_block_count_max, block_thread_extent_max, block_thread_count_max,
               thread_elem_extent_max, thread_elem_count_max, multi_processor_count=1, shared_mem_size_bytes=0,
               global_mem_size_bytes=0) -> AccDevProps:
    p = AccDevProps()
    p.multi_processor_count = multi_processor_count
    for i in range(dim):
        p.grid_block_extent_max[i] = grid_block_extent_max[i]
        p.block_thread_extent_max[i] = block_thread_extent_max[i]
        p.thread_elem_extent_max[i] = thread_elem_extent_max[i]
    p.grid_block_count_max = grid_block_count_max
    p.block_thread_count_max = block_thread_count_max
    p.thread_elem_count_max = thread_elem_count_max
    p.shared_mem_size_bytes = shared_mem_size_bytes
    p.global_mem_size_bytes = global_mem_size_bytes
    return p


def get_acc_dev_props(dev_idx: int, dim: int) -> AccDevProps:
    p = AccDevProps()
    check(_lib.load().b200_acc_dev_props_get(dev_idx, dim, C.byref(p)))
    return p


def sub_divide_grid_elems(grid_elem_extent: Sequence[int], thread_elem_extent: Sequence[int], props: AccDevProps,
                          kernel_block_thread_count_max: int = 0, block_thread_must_divide_grid_thread_extent: bool = True,
                          restriction: int = UNRESTRICTED) -> WorkDivMembers:
    dim = len(grid_elem_extent)
    arr = C.c_uint64 * dim
    g, b, e = arr(), arr(), arr()
    check(
        _lib.load().b200_subdivide_grid_elems(
            dim, arr(*grid_elem_extent), arr(*thread_elem_extent), C.byref(props), kernel_block_thread_count_max,
            int(block_thread_must_divide_grid_thread_extent), restriction, g, b, e,
        )
    )
    return WorkDivMembers(tuple(g), tuple(b), tuple(e))


def is_valid_work_div(wd: WorkDivMembers, props: AccDevProps, kernel_block_thread_count_max: int = 0) -> bool:
    dim = len(wd.grid_block_extent)
    arr = C.c_uint64 * dim
    ok = C.c_int(0)
    check(
        _lib.load().b200_is_valid_work_div(
            dim, arr(*wd.grid_block_extent), arr(*wd.block_thread_extent), arr(*wd.thread_elem_extent), C.byref(props),
            kernel_block_thread_count_max, C.byref(ok),
        )
    )
    return bool(ok.value)
