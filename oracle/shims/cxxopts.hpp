// TEST INFRASTRUCTURE ONLY -- the slice of cxxopts the reference's main.cpp:37-64 uses:
// Options(name).add_options()("s,long", desc[, value<T>()]) ; parse(argc, argv) ;
// result.count(name) ; result[name].as<std::string>() ; options.help().
#pragma once
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
namespace cxxopts {
struct ValueBase { virtual ~ValueBase() {} };
template <class T> struct TypedValue : ValueBase {};
template <class T> std::shared_ptr<ValueBase> value() { return std::make_shared<TypedValue<T>>(); }
struct OptionSpec { std::string s, l, desc; bool takes; };
class OptionValue {
 public:
  std::string v;
  template <class T> T as() const;
};
template <> inline std::string OptionValue::as<std::string>() const { return v; }
template <> inline int OptionValue::as<int>() const { return std::stoi(v); }
class ParseResult {
 public:
  size_t count(const std::string& k) const { auto it = counts.find(k); return it == counts.end() ? 0 : it->second; }
  const OptionValue& operator[](const std::string& k) const { return vals.at(k); }
  std::map<std::string, size_t> counts;
  std::map<std::string, OptionValue> vals;
};
class Options;
class Adder {
 public:
  explicit Adder(Options& o) : o_(o) {}
  Adder& operator()(const std::string& names, const std::string& desc, std::shared_ptr<ValueBase> v = nullptr);
 private:
  Options& o_;
};
class Options {
 public:
  explicit Options(std::string prog, std::string help = "") : prog_(std::move(prog)), help_(std::move(help)) {}
  Adder add_options() { return Adder(*this); }
  ParseResult parse(int argc, char** argv) {
    ParseResult r;
    for (int i = 1; i < argc; i++) {
      std::string a = argv[i];
      const OptionSpec* spec = nullptr;
      std::string inlineVal; bool hasInline = false;
      if (a.rfind("--", 0) == 0) {
        std::string name = a.substr(2);
        size_t eq = name.find('=');
        if (eq != std::string::npos) { inlineVal = name.substr(eq + 1); name = name.substr(0, eq); hasInline = true; }
        for (auto& s : specs) if (s.l == name) spec = &s;
      } else if (a.size() >= 2 && a[0] == '-') {
        std::string name = a.substr(1, 1);
        if (a.size() > 2) { inlineVal = a.substr(2); hasInline = true; }
        for (auto& s : specs) if (s.s == name) spec = &s;
      }
      if (!spec) continue;
      r.counts[spec->l]++;
      if (spec->takes) {
        if (hasInline) r.vals[spec->l].v = inlineVal;
        else if (i + 1 < argc) r.vals[spec->l].v = argv[++i];
      }
    }
    return r;
  }
  std::string help() const {
    std::ostringstream os;
    os << help_ << "\nUsage:\n  " << prog_ << " [OPTION...]\n\n";
    for (auto& s : specs) {
      std::string left = "  -" + s.s + ", --" + s.l + (s.takes ? " arg" : "");
      os << left;
      for (size_t k = left.size(); k < 22; k++) os << ' ';
      os << s.desc << "\n";
    }
    return os.str();
  }
  std::vector<OptionSpec> specs;
 private:
  std::string prog_, help_;
};
inline Adder& Adder::operator()(const std::string& names, const std::string& desc, std::shared_ptr<ValueBase> v) {
  OptionSpec s; s.desc = desc; s.takes = (bool)v;
  size_t c = names.find(',');
  if (c == std::string::npos) { s.l = names; } else { s.s = names.substr(0, c); s.l = names.substr(c + 1); }
  o_.specs.push_back(s);
  return *this;
}
}  // namespace cxxopts
