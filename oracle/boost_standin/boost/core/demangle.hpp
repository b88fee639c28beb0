/* Minimal stand-in for <boost/core/demangle.hpp>: TEST INFRASTRUCTURE ONLY
 * (reference use: include/alpaka/core/DemangleTypeNames.hpp:9,19). */
#ifndef B200_ORACLE_BOOST_CORE_DEMANGLE_HPP
#define B200_ORACLE_BOOST_CORE_DEMANGLE_HPP
#include <cstdlib>
#include <cxxabi.h>
#include <string>

namespace boost::core
{
    inline std::string demangle(char const* name)
    {
        int status = 0;
        char* p = abi::__cxa_demangle(name, nullptr, nullptr, &status);
        std::string r = (status == 0 && p) ? p : name;
        std::free(p);
        return r;
    }
} // namespace boost::core
#endif
