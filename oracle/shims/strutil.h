// TEST INFRASTRUCTURE ONLY -- minimal stand-in for the un-vendored `strutil` header the
// reference includes (main.cpp:9, utils.h:7, LoaderGPGPU.h:6).  Only the four calls the
// reference makes are provided.  Not product code; never linked into libmpc_b200.so.
#pragma once
#include <string>
#include <vector>
namespace strutil {
inline std::vector<std::string> split(const std::string& s, const std::string& delim) {
  std::vector<std::string> out;
  size_t pos = 0, hit;
  if (delim.empty()) { out.push_back(s); return out; }
  while ((hit = s.find(delim, pos)) != std::string::npos) {
    out.push_back(s.substr(pos, hit - pos));
    pos = hit + delim.size();
  }
  out.push_back(s.substr(pos));
  return out;
}
inline std::vector<std::string> split(const std::string& s, char delim) {
  return split(s, std::string(1, delim));
}
inline bool replace_all(std::string& s, const std::string& target, const std::string& repl) {
  if (target.empty()) return false;
  bool found = false;
  size_t pos = 0;
  while ((pos = s.find(target, pos)) != std::string::npos) {
    s.replace(pos, target.size(), repl);
    pos += repl.size();
    found = true;
  }
  return found;
}
inline bool ends_with(const std::string& s, const std::string& suffix) {
  return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}
inline bool contains(const std::string& s, const std::string& sub) {
  return s.find(sub) != std::string::npos;
}
}  // namespace strutil
