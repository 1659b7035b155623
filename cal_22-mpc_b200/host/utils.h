// Small host utilities: file test, config-name stem (reference src/utils.{h,cpp}) and `{}`-style number
// formatting so that the CSV files are byte-identical to what the reference writes through fmt::format.
#ifndef MPCB_UTILS_H_
#define MPCB_UTILS_H_

#include <cstdint>
#include <string>
#include <vector>

bool isFileExists(const std::string& filePath);          // utils.cpp:3-7
std::string parseConfig(const std::string& configPath);  // utils.cpp:9-21: file name without ".json"
std::vector<std::string> splitString(const std::string& s, const std::string& delim);
bool replaceAll(std::string& s, const std::string& target, const std::string& repl);
bool endsWith(const std::string& s, const std::string& suffix);

// fmt::format("{}", double): shortest round-trip digits; fixed notation for 1e-4 <= |x| < 1e16, otherwise
// d.ddde+XX; no trailing ".0" on integral values; "inf" / "nan" (SURVEY.md section 7 "CSV doubles").
std::string formatDouble(double v);
extern "C" int mpcb_format_double(double v, char* out, int cap);  // for the tests

#endif
