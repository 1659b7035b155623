// TEST INFRASTRUCTURE ONLY -- the subset of the jsoncpp API that the reference's
// VPC.cpp:84-327 calls, implemented over nlohmann::json (header shipped with the image).
// jsoncpp semantics reproduced: operator[] on a missing key/index yields a null value,
// and asInt/asFloat/asBool/asString on null yield 0 / 0.0f / false / "".
#pragma once
#include <istream>
#include <map>
#include <memory>
#include <string>
#include <nlohmann/json.hpp>
#define JSONCPP_STRING std::string
namespace Json {
class Value {
 public:
  Value() : j_(nullptr) {}
  explicit Value(const nlohmann::json& j) : j_(j) {}
  Value(Value&&) = default;
  Value& operator=(Value&&) = default;
  Value& operator=(bool b) { j_ = b; kids_.clear(); return *this; }
  Value& operator[](const char* key) { return child(std::string(key)); }
  Value& operator[](const std::string& key) { return child(key); }
  Value& operator[](int idx) {
    std::string k = "#" + std::to_string(idx);
    auto it = kids_.find(k);
    if (it != kids_.end()) return *it->second;
    auto v = std::make_unique<Value>();
    if (j_.is_array() && idx >= 0 && (size_t)idx < j_.size()) v->j_ = j_[(size_t)idx];
    return *(kids_[k] = std::move(v));
  }
  bool isNull() const { return j_.is_null(); }
  int asInt() const {
    if (j_.is_number_integer()) return j_.get<int>();
    if (j_.is_number()) return (int)j_.get<double>();
    if (j_.is_boolean()) return j_.get<bool>() ? 1 : 0;
    return 0;
  }
  float asFloat() const {
    if (j_.is_number()) return (float)j_.get<double>();
    if (j_.is_boolean()) return j_.get<bool>() ? 1.0f : 0.0f;
    return 0.0f;
  }
  bool asBool() const {
    if (j_.is_boolean()) return j_.get<bool>();
    if (j_.is_number()) return j_.get<double>() != 0.0;
    return false;
  }
  std::string asString() const {
    if (j_.is_string()) return j_.get<std::string>();
    return std::string();
  }
  unsigned size() const { return (j_.is_array() || j_.is_object()) ? (unsigned)j_.size() : 0u; }
  nlohmann::json j_;

 private:
  Value& child(const std::string& key) {
    auto it = kids_.find(key);
    if (it != kids_.end()) return *it->second;
    auto v = std::make_unique<Value>();
    if (j_.is_object()) {
      auto f = j_.find(key);
      if (f != j_.end()) v->j_ = *f;
    }
    return *(kids_[key] = std::move(v));
  }
  std::map<std::string, std::unique_ptr<Value>> kids_;
};
class CharReaderBuilder {
 public:
  Value& operator[](const char* key) { return settings_[key]; }
 private:
  std::map<std::string, Value> settings_;
};
inline bool parseFromStream(CharReaderBuilder&, std::istream& in, Value* root, std::string* errs) {
  try {
    nlohmann::json j = nlohmann::json::parse(in);
    *root = Value(j);
    return true;
  } catch (const std::exception& e) {
    if (errs) *errs = e.what();
    return false;
  }
}
}  // namespace Json
