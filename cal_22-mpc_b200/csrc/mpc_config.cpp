// Config front end of libmpc_b200: reference-format JSON -> flat mpc_config_pod.
//
// Replaces VPC::parseConfig (reference src/compressor/VPC.cpp:72-330).  The reference reads the
// file with jsoncpp; that library is not a dependency here, so a small recursive-descent JSON
// reader lives in this file.  jsoncpp lookup semantics the reference relies on are kept: a
// missing key or index reads as null, and null converts to 0 / 0.0f / false / "".
//
// Configs for which the reference has undefined behaviour (out-of-range table entries, a
// module order its hard-coded casts cannot handle, ...) are rejected instead of reproduced.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "mpc_capi.h"

namespace {

struct JValue {
  enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<JValue> arr;
  std::vector<std::pair<std::string, JValue>> obj;

  const JValue& get(const std::string& key) const {
    static const JValue null_value;
    if (type != Obj) return null_value;
    const JValue* hit = &null_value;
    for (auto& kv : obj)
      if (kv.first == key) hit = &kv.second;  // last duplicate wins, as in jsoncpp
    return *hit;
  }
  const JValue& at(size_t i) const {
    static const JValue null_value;
    if (type != Arr || i >= arr.size()) return null_value;
    return arr[i];
  }
  bool is_null() const { return type == Null; }
  int as_int() const {
    if (type == Num) return (int)num;
    if (type == Bool) return b ? 1 : 0;
    return 0;
  }
  float as_float() const {
    if (type == Num) return (float)num;
    if (type == Bool) return b ? 1.0f : 0.0f;
    return 0.0f;
  }
  bool as_bool() const {
    if (type == Bool) return b;
    if (type == Num) return num != 0.0;
    return false;
  }
  std::string as_string() const { return type == Str ? str : std::string(); }
};

class JParser {
 public:
  explicit JParser(const std::string& s) : s_(s) {}
  bool parse(JValue& out, std::string& err) {
    try {
      skip();
      out = value();
      skip();
      if (p_ != s_.size()) fail("trailing characters");
      return true;
    } catch (const std::string& e) {
      err = e;
      return false;
    }
  }

 private:
  [[noreturn]] void fail(const std::string& what) {
    throw std::string("JSON syntax error at offset ") + std::to_string(p_) + ": " + what;
  }
  void skip() {
    for (;;) {
      while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) p_++;
      // jsoncpp accepts C/C++ comments by default; the reference's "collecComments" key is a typo
      if (p_ + 1 < s_.size() && s_[p_] == '/' && s_[p_ + 1] == '/') {
        while (p_ < s_.size() && s_[p_] != '\n') p_++;
      } else if (p_ + 1 < s_.size() && s_[p_] == '/' && s_[p_ + 1] == '*') {
        size_t e = s_.find("*/", p_ + 2);
        if (e == std::string::npos) fail("unterminated comment");
        p_ = e + 2;
      } else {
        return;
      }
    }
  }
  JValue value() {
    if (p_ >= s_.size()) fail("unexpected end");
    char c = s_[p_];
    if (c == '{') return object();
    if (c == '[') return array();
    if (c == '"') { JValue v; v.type = JValue::Str; v.str = string(); return v; }
    if (s_.compare(p_, 4, "true") == 0) { p_ += 4; JValue v; v.type = JValue::Bool; v.b = true; return v; }
    if (s_.compare(p_, 5, "false") == 0) { p_ += 5; JValue v; v.type = JValue::Bool; v.b = false; return v; }
    if (s_.compare(p_, 4, "null") == 0) { p_ += 4; return JValue(); }
    return number();
  }
  JValue number() {
    size_t b = p_;
    if (p_ < s_.size() && (s_[p_] == '-' || s_[p_] == '+')) p_++;
    while (p_ < s_.size() && (isdigit((unsigned char)s_[p_]) || s_[p_] == '.' || s_[p_] == 'e' || s_[p_] == 'E' ||
                              s_[p_] == '-' || s_[p_] == '+'))
      p_++;
    if (b == p_) fail("unexpected character");
    JValue v;
    v.type = JValue::Num;
    try {
      v.num = std::stod(s_.substr(b, p_ - b));
    } catch (...) {
      fail("bad number");
    }
    return v;
  }
  std::string string() {
    std::string out;
    p_++;  // opening quote
    while (p_ < s_.size() && s_[p_] != '"') {
      char c = s_[p_++];
      if (c == '\\') {
        if (p_ >= s_.size()) fail("bad escape");
        char e = s_[p_++];
        switch (e) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            if (p_ + 4 > s_.size()) fail("bad \\u escape");
            unsigned cp = (unsigned)std::stoul(s_.substr(p_, 4), nullptr, 16);
            p_ += 4;
            if (cp < 0x80) out += (char)cp;
            else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
            else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: out += e;
        }
      } else {
        out += c;
      }
    }
    if (p_ >= s_.size()) fail("unterminated string");
    p_++;
    return out;
  }
  JValue array() {
    JValue v;
    v.type = JValue::Arr;
    p_++;
    skip();
    if (p_ < s_.size() && s_[p_] == ']') { p_++; return v; }
    for (;;) {
      skip();
      v.arr.push_back(value());
      skip();
      if (p_ < s_.size() && s_[p_] == ',') { p_++; continue; }
      if (p_ < s_.size() && s_[p_] == ']') { p_++; return v; }
      fail("expected , or ]");
    }
  }
  JValue object() {
    JValue v;
    v.type = JValue::Obj;
    p_++;
    skip();
    if (p_ < s_.size() && s_[p_] == '}') { p_++; return v; }
    for (;;) {
      skip();
      if (p_ >= s_.size() || s_[p_] != '"') fail("expected string key");
      std::string k = string();
      skip();
      if (p_ >= s_.size() || s_[p_] != ':') fail("expected :");
      p_++;
      skip();
      v.obj.emplace_back(k, value());
      skip();
      if (p_ < s_.size() && s_[p_] == ',') { p_++; continue; }
      if (p_ < s_.size() && s_[p_] == '}') { p_++; return v; }
      fail("expected , or }");
    }
  }
  const std::string& s_;
  size_t p_ = 0;
};

void set_err(char* err, size_t n, const char* fmt, ...) {
  if (!err || n == 0) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err, n, fmt, ap);
  va_end(ap);
}

const char* kPatternNames[] = {"ZerosPattern", "SingleOnePattern", "TwoConsecutiveOnesPattern", "MaskingPattern",
                               "UncompressedPattern"};

int build_pod(const JValue& root, mpc_config_pod* out, char* err, size_t err_len) {
  memset(out, 0, sizeof(*out));
  const JValue& ov = root.get("overview");
  const int n = ov.get("num_modules").as_int();  // VPC.cpp:99
  const int L = ov.get("lineSize").as_int();     // VPC.cpp:101
  if (n < 1 || n > MPC_MAX_MODULES) {
    set_err(err, err_len, "overview.num_modules = %d is outside [1, %d]", n, MPC_MAX_MODULES);
    return MPC_E_CONFIG;
  }
  out->num_modules = n;
  out->line_size = L;
  if (ov.get("encoding_bits").is_null()) {  // VPC.cpp:102-108
    int bits = (int)ceil(log2f((float)(n + 1)));
    for (int i = 0; i <= n; i++) out->enc_bits[i] = bits;
  } else {  // VPC.cpp:109-116: list index j <-> cluster j-1
    for (int i = 0; i <= n; i++) out->enc_bits[i] = ov.get("encoding_bits").at((size_t)i).as_int();
  }
  bool line_fn_set = false;
  for (int i = 0; i < n; i++) {
    const JValue& spec = root.get("modules").get(std::to_string(i));  // VPC.cpp:126
    mpc_module_pod& m = out->modules[i];
    std::string name = spec.get("name").as_string();
    if (name == "PredComp") {
      m.kind = MPC_MOD_PREDCOMP;
      const JValue& sub = spec.get("submodules");
      const JValue& ps = sub.get("ResidueModule").get("PredictorModule");  // VPC.cpp:139
      std::string pname = ps.get("name").as_string();
      const int pl = ps.get("LineSize").as_int();
      m.root = ps.get("RootIndex").as_int();
      const bool tabled = (pname == "WeightBasePredictor" || pname == "DiffBasePredictor");
      if (pname == "WeightBasePredictor") m.predictor = MPC_PRED_WEIGHT;
      else if (pname == "DiffBasePredictor") m.predictor = MPC_PRED_DIFF;
      else if (pname == "OneBasePredictor") m.predictor = MPC_PRED_ONE;
      else if (pname == "ConsecutiveBasePredictor") m.predictor = MPC_PRED_CONSEC;
      else {
        // message of VPC.cpp:180-184 (adjacent string literals there collapse the inner quotes)
        set_err(err, err_len, "%s is not a valid predictor module. Check the config file.", pname.c_str());
        return MPC_E_CONFIG;
      }
      if (tabled) {
        // Diff/Weight take their length from the predictor's own LineSize (PredictorModule.cpp:46,91);
        // a different length makes every later stage index out of range in the reference.
        if (pl != L) {
          set_err(err, err_len, "module %d: PredictorModule.LineSize = %d differs from overview.lineSize = %d", i, pl, L);
          return MPC_E_CONFIG;
        }
        if (L < 1 || L > MPC_MAX_LINE) {
          set_err(err, err_len, "overview.lineSize = %d is not supported (32, 64 or 128)", L);
          return MPC_E_CONFIG;
        }
        for (int j = 0; j < L; j++) {
          int bi = ps.get("BaseIndexTable").at((size_t)j).as_int();
          if (j != m.root && (bi < 0 || bi >= L)) {
            set_err(err, err_len, "module %d: BaseIndexTable[%d] = %d is outside the line", i, j, bi);
            return MPC_E_CONFIG;
          }
          m.base[j] = (uint8_t)(bi < 0 || bi >= L ? 0 : bi);
          if (m.predictor == MPC_PRED_DIFF) {
            m.diff[j] = (uint8_t)ps.get("DiffTable").at((size_t)j).as_int();  // PredictorModule.cpp:104
          } else if (j != m.root) {
            float w = ps.get("WeightTable").at((size_t)j).as_float();
            if (!(w > 0.0f) || std::isinf(w)) {
              set_err(err, err_len, "module %d: WeightTable[%d] = %g has no defined shift distance", i, j, (double)w);
              return MPC_E_CONFIG;
            }
            int s = (int)log2f(w);  // PredictorModule.cpp:31
            if (s < -8) s = -8;     // |shift| >= 8 already yields 0 for a byte
            if (s > 8) s = 8;
            m.shift[j] = (int8_t)s;
          }
        }
      }
      m.consecutive_xor = sub.get("XORModule").get("consecutiveXOR").as_bool() ? 1 : 0;  // VPC.cpp:194
      const JValue& sc = sub.get("ScanModule");
      m.table_size = sc.get("TableSize").as_int();  // VPC.cpp:198
      if (m.table_size < 0 || m.table_size > 8 * MPC_MAX_LINE || (L > 0 && m.table_size > 8 * L)) {
        set_err(err, err_len, "module %d: ScanModule.TableSize = %d exceeds 8 * lineSize", i, m.table_size);
        return MPC_E_CONFIG;
      }
      for (int j = 0; j < m.table_size; j++) {
        int r = sc.get("Rows").at((size_t)j).as_int();
        int c = sc.get("Cols").at((size_t)j).as_int();
        if (r < 0 || r > 7 || c < 0 || c >= L) {
          set_err(err, err_len, "module %d: scan entry %d = (row %d, col %d) is outside the 8 x %d bit-plane array", i, j, r, c, L);
          return MPC_E_CONFIG;
        }
        m.scan_row[j] = (uint8_t)r;
        m.scan_col[j] = (uint8_t)c;
      }
      // FPCModule block: parsed for its names only (VPC.cpp:209-303); its content never reaches
      // the encoder (VPC.h:278 uses a default-constructed FPCModule).
      const JValue& fpc = sub.get("FPCModule");
      int np = fpc.get("num_modules").as_int();
      for (int j = 0; j < np; j++) {
        std::string pn = fpc.get(std::to_string(j)).get("name").as_string();
        bool ok = false;
        for (const char* k : kPatternNames) ok = ok || pn == k;
        if (!ok) {
          set_err(err, err_len, "%s is not a valid pattern module. Check the config file.", pn.c_str());
          return MPC_E_CONFIG;
        }
      }
    } else if (name == "AllZero") {
      m.kind = MPC_MOD_ALLZERO;
      out->has_wordsame = 0;  // last AllZero/AllWordSame parsed decides the line function, VPC.cpp:312,318
      line_fn_set = true;
    } else if (name == "ByteplaneAllSame" || name == "AllWordSame") {
      m.kind = MPC_MOD_ALLWORDSAME;
      out->has_wordsame = 1;
      line_fn_set = true;
    } else {
      set_err(err, err_len, "\"%s\" is not a valid compression module. Check the config file.", name.c_str());
      return MPC_E_CONFIG;
    }
  }
  if (!line_fn_set) {
    set_err(err, err_len, "config has neither an AllZero nor an AllWordSame module (the reference would call an unset function pointer)");
    return MPC_E_CONFIG;
  }
  out->first_predcomp = out->has_wordsame ? 2 : 1;
  return mpc_config_validate(out, err, err_len);
}

}  // namespace

extern "C" int mpc_config_validate(const mpc_config_pod* c, char* err, size_t err_len) {
  if (!c) return MPC_E_ARG;
  const int L = c->line_size, n = c->num_modules;
  if (!(L == 32 || L == 64 || L == 128)) {
    set_err(err, err_len, "overview.lineSize = %d is not supported (32, 64 or 128)", L);
    return MPC_E_CONFIG;
  }
  if (n < 1 || n > MPC_MAX_MODULES) {
    set_err(err, err_len, "overview.num_modules = %d is outside [1, %d]", n, MPC_MAX_MODULES);
    return MPC_E_CONFIG;
  }
  // The reference casts module 0 to AllZeroModule (VPC.cpp:336), module 1 to AllWordSameModule
  // (VPC.cpp:353) and every module from first_predcomp on to PredCompModule (VPC.cpp:374).
  if (c->modules[0].kind != MPC_MOD_ALLZERO) {
    set_err(err, err_len, "module 0 must be AllZero");
    return MPC_E_CONFIG;
  }
  if (c->has_wordsame && (n < 2 || c->modules[1].kind != MPC_MOD_ALLWORDSAME)) {
    set_err(err, err_len, "an AllWordSame module must be module 1");
    return MPC_E_CONFIG;
  }
  if (c->first_predcomp != (c->has_wordsame ? 2 : 1)) {
    set_err(err, err_len, "first_predcomp is inconsistent with has_wordsame");
    return MPC_E_CONFIG;
  }
  for (int i = c->first_predcomp; i < n; i++) {
    const mpc_module_pod& m = c->modules[i];
    if (m.kind != MPC_MOD_PREDCOMP) {
      set_err(err, err_len, "module %d must be PredComp (modules after AllZero/AllWordSame are cast to PredCompModule)", i);
      return MPC_E_CONFIG;
    }
    if (m.root < 0 || m.root >= L) {
      set_err(err, err_len, "module %d: RootIndex = %d is outside the line", i, m.root);
      return MPC_E_CONFIG;
    }
    if (m.predictor == MPC_PRED_CONSEC && m.root != 0) {
      set_err(err, err_len, "module %d: ConsecutiveBasePredictor needs RootIndex 0 (the reference reads index -1 otherwise)", i);
      return MPC_E_CONFIG;
    }
    if (m.predictor < MPC_PRED_ONE || m.predictor > MPC_PRED_WEIGHT) {
      set_err(err, err_len, "module %d: unknown predictor id %d", i, m.predictor);
      return MPC_E_CONFIG;
    }
    if (m.table_size < 0 || m.table_size > 8 * L) {
      set_err(err, err_len, "module %d: ScanModule.TableSize = %d exceeds 8 * lineSize", i, m.table_size);
      return MPC_E_CONFIG;
    }
    for (int j = 0; j < m.table_size; j++)
      if (m.scan_row[j] > 7 || m.scan_col[j] >= L) {
        set_err(err, err_len, "module %d: scan entry %d is outside the bit-plane array", i, j);
        return MPC_E_CONFIG;
      }
    if (m.predictor == MPC_PRED_DIFF || m.predictor == MPC_PRED_WEIGHT)
      for (int j = 0; j < L; j++)
        if (j != m.root && m.base[j] >= L) {
          set_err(err, err_len, "module %d: BaseIndexTable[%d] is outside the line", i, j);
          return MPC_E_CONFIG;
        }
  }
  for (int i = 0; i <= n; i++)
    if (c->enc_bits[i] < 0 || c->enc_bits[i] > MPC_MAX_ENC_BITS) {
      set_err(err, err_len, "encoding_bits[%d] = %d is outside [0, %d]", i, c->enc_bits[i], MPC_MAX_ENC_BITS);
      return MPC_E_CONFIG;
    }
  return MPC_OK;
}

extern "C" int mpc_config_from_json_text(const char* text, mpc_config_pod* out, char* err, size_t err_len) {
  if (!text || !out) return MPC_E_ARG;
  std::string s(text), perr;
  JValue root;
  JParser p(s);
  if (!p.parse(root, perr)) {
    set_err(err, err_len, "%s", perr.c_str());
    return MPC_E_CONFIG;
  }
  return build_pod(root, out, err, err_len);
}

extern "C" int mpc_config_from_json_file(const char* path, mpc_config_pod* out, char* err, size_t err_len) {
  if (!path || !out) return MPC_E_ARG;
  std::ifstream f(path);
  if (!f.is_open()) {
    set_err(err, err_len, "Invalid File! \"%s\" is not valid path.", path);  // VPC.cpp:79
    return MPC_E_IO;
  }
  std::stringstream ss;
  ss << f.rdbuf();
  std::string text = ss.str();
  int rc = mpc_config_from_json_text(text.c_str(), out, err, err_len);
  if (rc == MPC_E_CONFIG && err && strncmp(err, "JSON syntax", 11) == 0) {
    std::string first(err);
    set_err(err, err_len, "%s\nParsing ERROR! \"%s\" is not valid json file.", first.c_str(), path);  // VPC.cpp:92-93
  }
  return rc;
}
