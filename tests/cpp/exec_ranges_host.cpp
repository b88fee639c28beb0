// tests/cpp/exec_ranges_host.cpp -- host-only driver for tests/test_exec_ranges_cpu.py: prints the index sequences of the
// three range types behind uniformElements / uniformGroups / uniformGroupElements / independentGroup* (include/alpaka/
// b200/Exec.hpp) for parameters given on stdin, one case per line:
//   R start run pitch extent      RunHopRange   (uniformElementsAlong, independentGroupElementsAlong)
//   H start pitch extent          HopRange      (uniformGroupsAlong, independentGroupsAlong)
//   G origin lo hi                GroupRange    (uniformGroupElementsAlong), printed as global:local
// Built with plain g++ (no CUDA): the ranges are ordinary host/device value types.
#include <alpaka/alpaka.hpp>

#include <cstdint>
#include <iostream>
#include <sstream>
#include <string>

template<typename TIdx>
void run(std::istream& in)
{
    std::string line;
    while(std::getline(in, line))
    {
        std::istringstream ls(line);
        char kind = 0;
        long long p[4] = {};
        ls >> kind;
        for(int i = 0; i < 4; ++i)
            ls >> p[i];
        std::ostringstream out;
        if(kind == 'R')
            for(auto i : alpaka::b200x::RunHopRange<TIdx>(TIdx(p[0]), TIdx(p[1]), TIdx(p[2]), TIdx(p[3])))
                out << static_cast<long long>(i) << ' ';
        else if(kind == 'H')
            for(auto i : alpaka::b200x::HopRange<TIdx>(TIdx(p[0]), TIdx(p[1]), TIdx(p[2])))
                out << static_cast<long long>(i) << ' ';
        else if(kind == 'G')
            for(auto e : alpaka::b200x::GroupRange<TIdx>(TIdx(p[0]), TIdx(p[1]), TIdx(p[2])))
                out << static_cast<long long>(e.global) << ':' << static_cast<long long>(e.local) << ' ';
        else
            continue;
        std::cout << out.str() << '\n';
    }
}

int main(int argc, char** argv)
{
    std::string const t = argc > 1 ? argv[1] : "u32";
    if(t == "u32")
        run<std::uint32_t>(std::cin);
    else if(t == "i32")
        run<std::int32_t>(std::cin);
    else if(t == "u64")
        run<std::uint64_t>(std::cin);
    else
        return 2;
    return 0;
}
