// tests/cpp/heat_schedule_host.cpp -- host-only driver for tests/test_heat_schedule_cpu.py: prints the launch depths
// alpaka::b200::heatNextDepth (include/alpaka/b200/Heat2D.hpp) chooses for "n depth minDepth" lines on stdin, the way
// Heat2DStepper::steps / Heat2DSlabs::steps consume them. Built with plain g++ (no CUDA).
#include <alpaka/alpaka.hpp>

#include <cstdint>
#include <iostream>
#include <sstream>
#include <string>

auto main() -> int
{
    std::string line;
    while(std::getline(std::cin, line))
    {
        std::istringstream ls(line);
        long long n = 0;
        int depth = 0, minDepth = 1;
        ls >> n >> depth >> minDepth;
        std::ostringstream out;
        auto left = static_cast<std::uint32_t>(n);
        while(left > 0)
        {
            auto const k = alpaka::b200::heatNextDepth(left, depth, minDepth);
            if(k == 0)
            {
                out << "X";
                break;
            }
            out << k << ' ';
            left -= k;
        }
        std::cout << out.str() << '\n';
    }
    return 0;
}
