// alpaka_b200/csrc/b200_workdiv.cpp -- work-division selection and validation (host logic, no device needed).
//
// Replaces alpaka::subDivideGridElems (include/alpaka/workdiv/WorkDivHelpers.hpp:133-309), isValidWorkDiv
// (:406-549) and the device-property query behind getAccDevProps (acc/AccGpuUniformCudaHipRt.hpp:113-187).
// Behaviour is pinned by the reference's device-independent known-answer tests
// (test/unit/workDiv/src/WorkDivHelpersTest.cpp:34-180), replayed in tests/test_workdiv.py.
// All vectors use alpaka order: index 0 is the slowest dimension.
#include "b200/b200.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace b200
{
    int fail(int code, char const* what, char const* file, int line);
}

namespace
{
    using u64 = uint64_t;

    u64 prod(u64 const* v, int dim)
    {
        u64 p = 1;
        for(int i = 0; i < dim; ++i)
            p *= v[i];
        return p;
    }

    u64 divCeil(u64 a, u64 b)
    {
        return (a + b - 1) / b;
    }

    // floor(value^(1/n)) by search (core/Utility.hpp:42-62 nthRootFloor)
    u64 nthRootFloor(u64 value, int n)
    {
        u64 lo = 0, hi = value;
        while(lo < hi)
        {
            u64 const mid = lo + (hi - lo + 1) / 2;
            // mid^n <= value ?
            u64 p = 1;
            bool over = false;
            for(int i = 0; i < n; ++i)
            {
                if(mid != 0 && p > value / mid)
                {
                    over = true;
                    break;
                }
                p *= mid;
            }
            if(!over && p <= value)
                lo = mid;
            else
                hi = mid - 1;
        }
        return lo;
    }

    bool propsValid(b200_acc_dev_props const& p, int dim)
    {
        if(p.grid_block_count_max < 1 || p.block_thread_count_max < 1 || p.thread_elem_count_max < 1)
            return false;
        for(int i = 0; i < dim; ++i)
            if(p.grid_block_extent_max[i] < 1 || p.block_thread_extent_max[i] < 1 || p.thread_elem_extent_max[i] < 1)
                return false;
        return true;
    }

    // largest d <= maxDivisor with dividend % d == 0
    u64 divisorAtMost(u64 dividend, u64 maxDivisor)
    {
        u64 d = maxDivisor;
        while(dividend % d != 0)
            --d;
        return d;
    }
} // namespace

extern "C"
{
    int b200_subdivide_grid_elems(
        int dim,
        uint64_t const* gridElemExtent,
        uint64_t const* threadElemExtentIn,
        b200_acc_dev_props const* props,
        uint64_t kernelBlockThreadCountMax,
        int mustDivide,
        int restriction,
        uint64_t* gridBlockExtent,
        uint64_t* blockThreadExtent,
        uint64_t* threadElemExtent)
    {
        if(dim < 1 || dim > 4 || !gridElemExtent || !threadElemExtentIn || !props || !gridBlockExtent
           || !blockThreadExtent || !threadElemExtent)
            return b200::fail(B200_EINVAL, "subdivide_grid_elems arguments", __FILE__, __LINE__);
        if(restriction < B200_SUBDIV_EQUAL_EXTENT || restriction > B200_SUBDIV_UNRESTRICTED)
            return b200::fail(B200_EINVAL, "restriction", __FILE__, __LINE__);
        if(!propsValid(*props, dim))
            return b200::fail(B200_EINVAL, "acc dev props", __FILE__, __LINE__);
        for(int i = 0; i < dim; ++i)
            if(gridElemExtent[i] < 1 || threadElemExtentIn[i] < 1
               || threadElemExtentIn[i] > props->thread_elem_extent_max[i])
                return b200::fail(B200_EINVAL, "extents", __FILE__, __LINE__);
        if(prod(threadElemExtentIn, dim) > props->thread_elem_count_max)
            return b200::fail(B200_EINVAL, "thread elem count", __FILE__, __LINE__);

        // elements per thread never exceed the grid; threads needed per dimension
        u64 gridThreadExtent[4];
        for(int i = 0; i < dim; ++i)
        {
            threadElemExtent[i] = std::min(threadElemExtentIn[i], gridElemExtent[i]);
            gridThreadExtent[i] = divCeil(gridElemExtent[i], threadElemExtent[i]);
        }

        // start from the largest block the device allows, clipped to the grid
        for(int i = 0; i < dim; ++i)
            blockThreadExtent[i] = std::min(props->block_thread_extent_max[i], gridThreadExtent[i]);
        if(restriction == B200_SUBDIV_EQUAL_EXTENT)
        {
            u64 m = *std::min_element(blockThreadExtent, blockThreadExtent + dim);
            std::fill(blockThreadExtent, blockThreadExtent + dim, m != 0 ? m : 1);
        }

        u64 const countMax = kernelBlockThreadCountMax != 0 ? kernelBlockThreadCountMax : props->block_thread_count_max;
        for(int i = 0; i < dim; ++i)
            blockThreadExtent[i] = std::min(blockThreadExtent[i], countMax);

        if(countMax == 1)
        {
            std::fill(blockThreadExtent, blockThreadExtent + dim, nthRootFloor(countMax, dim));
        }
        else if(prod(blockThreadExtent, dim) > countMax)
        {
            if(restriction == B200_SUBDIV_EQUAL_EXTENT)
            {
                std::fill(blockThreadExtent, blockThreadExtent + dim, nthRootFloor(countMax, dim));
            }
            else if(restriction == B200_SUBDIV_CLOSE_TO_EQUAL_EXTENT)
            {
                // halve the (first) largest extent until the block fits
                while(prod(blockThreadExtent, dim) > countMax)
                {
                    int const imax = int(std::max_element(blockThreadExtent, blockThreadExtent + dim) - blockThreadExtent);
                    blockThreadExtent[imax] /= 2;
                }
            }
            else
            {
                // halve the smallest extent that is still > 1, never touching the fastest dimension
                // (the reference's min_element runs over [begin, end-1) with 1s ordered last)
                while(prod(blockThreadExtent, dim) > countMax)
                {
                    int pick = 0;
                    for(int i = 1; i < dim - 1; ++i)
                    {
                        u64 const a = blockThreadExtent[i], b = blockThreadExtent[pick];
                        bool const less = (a == 1) ? false : (b == 1) ? true : a < b;
                        if(less)
                            pick = i;
                    }
                    blockThreadExtent[pick] /= 2;
                }
            }
        }

        if(mustDivide)
        {
            if(restriction == B200_SUBDIV_EQUAL_EXTENT)
            {
                // greatest d such that, in every dimension, gridThreadExtent[i] % d == 0 and the co-divisor
                // gridThreadExtent[i] / d' construction of the reference is respected: the reference collects, per
                // dimension, { g / k : k in [1, min(g, b)], g % k == 0 } and takes the largest common member.
                auto members = [&](int i)
                {
                    std::vector<u64> v;
                    u64 const g = gridThreadExtent[i];
                    u64 const lim = std::min(g, blockThreadExtent[i]);
                    for(u64 k = 1; k <= lim; ++k)
                        if(g % k == 0)
                            v.push_back(g / k);
                    std::sort(v.begin(), v.end());
                    return v;
                };
                // NB: the reference intersects dimension 0 with dimension i for each i and keeps only the LAST
                // intersection (dim 0 with dim-1); reproduced as is.
                std::vector<u64> common = members(0);
                if(dim > 1)
                {
                    std::vector<u64> const last = members(dim - 1);
                    std::vector<u64> inter;
                    std::set_intersection(common.begin(), common.end(), last.begin(), last.end(), std::back_inserter(inter));
                    common.swap(inter);
                }
                u64 const d = common.empty() ? 1 : common.back();
                std::fill(blockThreadExtent, blockThreadExtent + dim, d);
            }
            else
            {
                for(int i = 0; i < dim; ++i)
                    blockThreadExtent[i] = divisorAtMost(gridThreadExtent[i], blockThreadExtent[i]);
            }
        }

        for(int i = 0; i < dim; ++i)
            gridBlockExtent[i] = divCeil(gridThreadExtent[i], blockThreadExtent[i]);

        // final clamp to the device limits
        for(int i = 0; i < dim; ++i)
        {
            gridBlockExtent[i] = std::min(gridBlockExtent[i], props->grid_block_extent_max[i]);
            blockThreadExtent[i] = std::min(blockThreadExtent[i], props->block_thread_extent_max[i]);
            if(props->thread_elem_extent_max[i] < threadElemExtentIn[i])
                threadElemExtent[i] = props->thread_elem_extent_max[i];
        }
        return 0;
    }

    int b200_is_valid_work_div(
        int dim,
        uint64_t const* gridBlockExtent,
        uint64_t const* blockThreadExtent,
        uint64_t const* threadElemExtent,
        b200_acc_dev_props const* props,
        uint64_t kernelBlockThreadCountMax,
        int* isValid)
    {
        if(dim < 1 || dim > 4 || !gridBlockExtent || !blockThreadExtent || !threadElemExtent || !props || !isValid)
            return b200::fail(B200_EINVAL, "is_valid_work_div arguments", __FILE__, __LINE__);
        *isValid = 0;
        // workdiv/WorkDivHelpers.hpp:406-470
        if(prod(gridBlockExtent, dim) == 0 || prod(blockThreadExtent, dim) == 0 || prod(threadElemExtent, dim) == 0)
            return 0;
        if(props->grid_block_count_max < prod(gridBlockExtent, dim))
            return 0;
        if(props->block_thread_count_max < prod(blockThreadExtent, dim))
            return 0;
        if(props->thread_elem_count_max < prod(threadElemExtent, dim))
            return 0;
        if(kernelBlockThreadCountMax != 0 && kernelBlockThreadCountMax < prod(blockThreadExtent, dim))
            return 0;
        for(int i = 0; i < dim; ++i)
        {
            if(props->grid_block_extent_max[i] < gridBlockExtent[i]
               || props->block_thread_extent_max[i] < blockThreadExtent[i]
               || props->thread_elem_extent_max[i] < threadElemExtent[i])
                return 0;
        }
        *isValid = 1;
        return 0;
    }
}
