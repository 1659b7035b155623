#include "utils.h"

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

bool isFileExists(const std::string& filePath) {
  std::ifstream f(filePath);
  return f.good();
}

std::vector<std::string> splitString(const std::string& s, const std::string& delim) {
  std::vector<std::string> out;
  size_t pos = 0, hit;
  while (!delim.empty() && (hit = s.find(delim, pos)) != std::string::npos) {
    out.push_back(s.substr(pos, hit - pos));
    pos = hit + delim.size();
  }
  out.push_back(s.substr(pos));
  return out;
}

bool replaceAll(std::string& s, const std::string& target, const std::string& repl) {
  bool found = false;
  size_t pos = 0;
  while (!target.empty() && (pos = s.find(target, pos)) != std::string::npos) {
    s.replace(pos, target.size(), repl);
    pos += repl.size();
    found = true;
  }
  return found;
}

bool endsWith(const std::string& s, const std::string& suffix) {
  return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

std::string parseConfig(const std::string& configPath) {
  std::vector<std::string> parts = splitString(configPath, "/");
  std::string name = parts.back();
  if (!replaceAll(name, ".json", "")) {
    printf("String parsing have failed!\n");  // utils.cpp:16
    exit(1);
  }
  return name;
}

std::string formatDouble(double v) {
  if (std::isnan(v)) return std::signbit(v) ? "-nan" : "nan";
  if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
  if (v == 0.0) return std::signbit(v) ? "-0" : "0";
  char buf[64];
  auto res = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);  // shortest round trip
  std::string sci(buf, res.ptr);
  std::string out;
  size_t p = 0;
  if (sci[0] == '-') { out = "-"; p = 1; }
  size_t e = sci.find('e');
  std::string digits;
  for (size_t i = p; i < e; i++)
    if (sci[i] != '.') digits += sci[i];
  int exp10 = atoi(sci.c_str() + e + 1);
  const int n = (int)digits.size();
  if (exp10 >= -4 && exp10 < 16) {
    if (exp10 >= 0) {
      if (n <= exp10 + 1) {
        out += digits + std::string((size_t)(exp10 + 1 - n), '0');
      } else {
        out += digits.substr(0, (size_t)exp10 + 1) + "." + digits.substr((size_t)exp10 + 1);
      }
    } else {
      out += "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
    }
  } else {
    out += digits.substr(0, 1);
    if (n > 1) out += "." + digits.substr(1);
    char eb[16];
    snprintf(eb, sizeof(eb), "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
    out += eb;
  }
  return out;
}

extern "C" int mpcb_format_double(double v, char* out, int cap) {
  std::string s = formatDouble(v);
  if ((int)s.size() + 1 > cap) return -1;
  memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}
