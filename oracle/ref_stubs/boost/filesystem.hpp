// Stand-in for <boost/filesystem.hpp> and the two boost::algorithm helpers the reference's template loader
// uses — TEST INFRASTRUCTURE (oracle/_ref build only).  std::filesystem has the same interface for these calls.
#pragma once
#include <filesystem>
#include <string>
#include <vector>
namespace boost {
namespace filesystem = std::filesystem;
struct is_any_of { std::string set; explicit is_any_of(const char* s) : set(s) {} };
inline void split(std::vector<std::string>& out, const std::string& in, const is_any_of& sep) {
    out.clear();
    std::string cur;
    for (char c : in) {
        if (sep.set.find(c) != std::string::npos) { out.push_back(cur); cur.clear(); }
        else cur.push_back(c);
    }
    out.push_back(cur);
}
}  // namespace boost
