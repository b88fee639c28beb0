// examples/common/cli.hpp -- tiny --key=value command line reader and raw binary file I/O shared by the example drivers.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace cli
{
    class Args
    {
    public:
        Args(int argc, char** argv)
        {
            for(int i = 1; i < argc; ++i)
            {
                std::string a = argv[i];
                if(a.rfind("--", 0) != 0)
                    throw std::runtime_error("unexpected argument: " + a);
                a = a.substr(2);
                auto const eq = a.find('=');
                if(eq == std::string::npos)
                    m_kv[a] = "1";
                else
                    m_kv[a.substr(0, eq)] = a.substr(eq + 1);
            }
        }
        [[nodiscard]] auto has(std::string const& k) const -> bool
        {
            return m_kv.count(k) != 0;
        }
        [[nodiscard]] auto str(std::string const& k, std::string const& dflt = "") const -> std::string
        {
            auto const it = m_kv.find(k);
            return it == m_kv.end() ? dflt : it->second;
        }
        [[nodiscard]] auto u64(std::string const& k, std::uint64_t dflt) const -> std::uint64_t
        {
            auto const it = m_kv.find(k);
            return it == m_kv.end() ? dflt : std::strtoull(it->second.c_str(), nullptr, 0);
        }
        [[nodiscard]] auto f64(std::string const& k, double dflt) const -> double
        {
            auto const it = m_kv.find(k);
            return it == m_kv.end() ? dflt : std::strtod(it->second.c_str(), nullptr);
        }

    private:
        std::map<std::string, std::string> m_kv;
    };

    inline void writeFile(std::string const& path, void const* data, std::size_t bytes)
    {
        std::FILE* f = std::fopen(path.c_str(), "wb");
        if(f == nullptr || std::fwrite(data, 1, bytes, f) != bytes)
            throw std::runtime_error("cannot write " + path);
        std::fclose(f);
    }

    inline void readFile(std::string const& path, void* data, std::size_t bytes)
    {
        std::FILE* f = std::fopen(path.c_str(), "rb");
        if(f == nullptr || std::fread(data, 1, bytes, f) != bytes)
            throw std::runtime_error("cannot read " + std::to_string(bytes) + " bytes from " + path);
        std::fclose(f);
    }
} // namespace cli
