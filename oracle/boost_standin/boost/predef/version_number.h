/* Minimal stand-in for <boost/predef/version_number.h>: TEST INFRASTRUCTURE ONLY.
 * Boost is not installed in this image; the reference (alpaka) only needs the version-number
 * macro family from Boost.Predef to pick compiler/OS code paths (reference:
 * include/alpaka/version.hpp:7, include/alpaka/core/BoostPredef.hpp:8). Nothing here computes
 * anything on the hot path. Used only when building oracle/_ref from /root/reference. */
#ifndef B200_ORACLE_BOOST_PREDEF_VERSION_NUMBER_H
#define B200_ORACLE_BOOST_PREDEF_VERSION_NUMBER_H

#define BOOST_VERSION_NUMBER(major, minor, patch) \
    ((((major) % 100) * 10000000) + (((minor) % 100) * 100000) + ((patch) % 100000))
#define BOOST_VERSION_NUMBER_MAX BOOST_VERSION_NUMBER(99, 99, 99999)
#define BOOST_VERSION_NUMBER_ZERO BOOST_VERSION_NUMBER(0, 0, 0)
#define BOOST_VERSION_NUMBER_MIN BOOST_VERSION_NUMBER(0, 0, 1)
#define BOOST_VERSION_NUMBER_AVAILABLE BOOST_VERSION_NUMBER_MIN
#define BOOST_VERSION_NUMBER_NOT_AVAILABLE BOOST_VERSION_NUMBER_ZERO
#define BOOST_VERSION_NUMBER_MAJOR(N) (((N) / 10000000) % 100)
#define BOOST_VERSION_NUMBER_MINOR(N) (((N) / 100000) % 100)
#define BOOST_VERSION_NUMBER_PATCH(N) ((N) % 100000)

/* decimal VVRRP (e.g. __CUDACC_VER-style) and YYYYMMDD encodings */
#define BOOST_PREDEF_MAKE_10_VVRRP(V) BOOST_VERSION_NUMBER(((V) / 1000) % 100, ((V) / 10) % 100, (V) % 10)
#define BOOST_PREDEF_MAKE_YYYYMMDD(V) \
    BOOST_VERSION_NUMBER((((V) / 10000) % 10000) - 1970, ((V) / 100) % 100, (V) % 100)

#endif
