// Stand-in for yaml-cpp: flat "key: value  # comment" files, which is all cfg/options*.yaml contain.
#pragma once
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
namespace YAML {
class Node {
public:
    Node() = default;
    explicit Node(std::string scalar) : scalar_(std::move(scalar)) {}
    Node operator[](const std::string& key) const {
        auto it = map_.find(key);
        if (it == map_.end()) throw std::runtime_error("yaml stand-in: missing key " + key);
        return Node(it->second);
    }
    template <typename T> T as() const {
        if constexpr (std::is_same<T, std::string>::value) return scalar_;
        else if constexpr (std::is_same<T, bool>::value) return scalar_ == "true" || scalar_ == "True" || scalar_ == "1";
        else { std::stringstream ss(scalar_); T v{}; ss >> v; return v; }
    }
    std::map<std::string, std::string> map_;
private:
    std::string scalar_;
};
inline Node LoadFile(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("yaml stand-in: cannot open " + path);
    Node n;
    std::string line;
    while (std::getline(f, line)) {
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        const size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        auto trim = [](std::string s) { const size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r"); return a == std::string::npos ? std::string() : s.substr(a, b - a + 1); };
        n.map_[trim(line.substr(0, colon))] = trim(line.substr(colon + 1));
    }
    return n;
}
}  // namespace YAML
