// include/alpaka/test/Check.hpp -- device-side assertion used by test kernels (reference: include/alpaka/test/Check.hpp):
// a kernel takes `bool* success` and records failures instead of aborting.
#pragma once

#include <cstdio>

#define ALPAKA_CHECK(success, expression)                                                                             \
    do                                                                                                                \
    {                                                                                                                 \
        if(!(expression))                                                                                             \
        {                                                                                                             \
            printf("ALPAKA_CHECK failed because '!(%s)'\n", #expression);                                             \
            success = false;                                                                                          \
        }                                                                                                             \
    } while(0)
