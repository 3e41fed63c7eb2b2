// Byte-level BPE (tokenizer.json) reader / encoder / decoder -- see op/byte_bpe.h.
#include "op/byte_bpe.h"

#include <algorithm>
#include <climits>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>

namespace op {
namespace {
#include "unicode_tables.inc"

// ---- minimal JSON ------------------------------------------------------------------------------------
struct JValue {
  enum Type { kNull, kBool, kNumber, kString, kArray, kObject } type = kNull;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<JValue> arr;
  std::vector<std::pair<std::string, JValue>> obj;
  const JValue* get(const char* key) const {
    if (type != kObject) return nullptr;
    for (const auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
};

void append_utf8(std::string& out, uint32_t cp) {
  if (cp < 0x80) {
    out.push_back(static_cast<char>(cp));
  } else if (cp < 0x800) {
    out.push_back(static_cast<char>(0xC0 | (cp >> 6)));
    out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
  } else if (cp < 0x10000) {
    out.push_back(static_cast<char>(0xE0 | (cp >> 12)));
    out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
    out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
  } else {
    out.push_back(static_cast<char>(0xF0 | (cp >> 18)));
    out.push_back(static_cast<char>(0x80 | ((cp >> 12) & 0x3F)));
    out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
    out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
  }
}

struct JsonParser {
  const char* p;
  const char* end;
  bool ok = true;
  void ws() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p;
  }
  bool lit(const char* s) {
    const size_t n = std::strlen(s);
    if (static_cast<size_t>(end - p) >= n && std::memcmp(p, s, n) == 0) {
      p += n;
      return true;
    }
    return false;
  }
  int hex4() {
    if (end - p < 4) return -1;
    int v = 0;
    for (int i = 0; i < 4; ++i) {
      const char c = p[i];
      v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0';
      else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
      else return -1;
    }
    p += 4;
    return v;
  }
  std::string string() {
    std::string out;
    if (p >= end || *p != '"') {
      ok = false;
      return out;
    }
    ++p;
    while (p < end && *p != '"') {
      if (*p != '\\') {
        out.push_back(*p++);
        continue;
      }
      if (++p >= end) break;
      const char e = *p++;
      switch (e) {
        case 'n': out.push_back('\n'); break;
        case 't': out.push_back('\t'); break;
        case 'r': out.push_back('\r'); break;
        case 'b': out.push_back('\b'); break;
        case 'f': out.push_back('\f'); break;
        case 'u': {
          int hi = hex4();
          if (hi < 0) { ok = false; return out; }
          uint32_t cp = static_cast<uint32_t>(hi);
          if (hi >= 0xD800 && hi <= 0xDBFF && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
            p += 2;
            const int lo = hex4();
            if (lo < 0xDC00 || lo > 0xDFFF) { ok = false; return out; }
            cp = 0x10000 + ((static_cast<uint32_t>(hi) - 0xD800) << 10) + (static_cast<uint32_t>(lo) - 0xDC00);
          }
          append_utf8(out, cp);
          break;
        }
        default: out.push_back(e);  // \" \\ \/
      }
    }
    if (p >= end) ok = false;
    else ++p;
    return out;
  }
  JValue value(int depth = 0) {
    JValue v;
    ws();
    if (p >= end || depth > 64) {
      ok = false;
      return v;
    }
    if (*p == '{') {
      ++p;
      v.type = JValue::kObject;
      ws();
      if (p < end && *p == '}') { ++p; return v; }
      while (ok) {
        ws();
        std::string key = string();
        ws();
        if (!ok || p >= end || *p != ':') { ok = false; break; }
        ++p;
        v.obj.emplace_back(std::move(key), value(depth + 1));
        ws();
        if (p < end && *p == ',') { ++p; continue; }
        if (p < end && *p == '}') { ++p; break; }
        ok = false;
      }
    } else if (*p == '[') {
      ++p;
      v.type = JValue::kArray;
      ws();
      if (p < end && *p == ']') { ++p; return v; }
      while (ok) {
        v.arr.push_back(value(depth + 1));
        ws();
        if (p < end && *p == ',') { ++p; continue; }
        if (p < end && *p == ']') { ++p; break; }
        ok = false;
      }
    } else if (*p == '"') {
      v.type = JValue::kString;
      v.str = string();
    } else if (lit("true")) {
      v.type = JValue::kBool, v.b = true;
    } else if (lit("false")) {
      v.type = JValue::kBool;
    } else if (lit("null")) {
    } else {
      char* stop = nullptr;
      v.type = JValue::kNumber;
      v.num = std::strtod(p, &stop);
      if (stop == p || stop > end) ok = false;
      else p = stop;
    }
    return v;
  }
};

// ---- Unicode helpers -----------------------------------------------------------------------------------
template <size_t N>
bool in_ranges(const uint32_t (&table)[N][2], uint32_t cp) {
  size_t lo = 0, hi = N;
  while (lo < hi) {
    const size_t mid = (lo + hi) / 2;
    if (cp < table[mid][0]) hi = mid;
    else if (cp > table[mid][1]) lo = mid + 1;
    else return true;
  }
  return false;
}
bool is_letter(uint32_t c) { return in_ranges(kUnicodeLetters, c); }
bool is_number(uint32_t c) { return in_ranges(kUnicodeNumbers, c); }
bool is_space(uint32_t c) {  // \s = Unicode White_Space
  return (c >= 9 && c <= 13) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) ||
         c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}
bool is_newline(uint32_t c) { return c == '\r' || c == '\n'; }

struct Cps {
  std::vector<uint32_t> cp;
  std::vector<size_t> off;  // byte offset of each code point, plus the end
};
Cps decode_utf8(std::string_view s) {
  Cps r;
  size_t i = 0;
  while (i < s.size()) {
    const unsigned char c = static_cast<unsigned char>(s[i]);
    int n = c < 0x80 ? 1 : (c >> 5) == 0x6 ? 2 : (c >> 4) == 0xE ? 3 : (c >> 3) == 0x1E ? 4 : 1;
    if (i + n > s.size()) n = 1;
    uint32_t cp = n == 1 ? c : n == 2 ? (c & 0x1F) : n == 3 ? (c & 0x0F) : (c & 0x07);
    bool good = n == 1 ? c < 0x80 : true;
    for (int k = 1; k < n && good; ++k) {
      const unsigned char cc = static_cast<unsigned char>(s[i + k]);
      if ((cc & 0xC0) != 0x80) good = false;
      cp = (cp << 6) | (cc & 0x3F);
    }
    if (!good) {
      cp = 0xFFFD;  // stray byte: its own piece, never a letter or digit
      n = 1;
    }
    r.cp.push_back(cp);
    r.off.push_back(i);
    i += n;
  }
  r.off.push_back(s.size());
  return r;
}

// GPT-2 bytes_to_unicode: printable bytes map to themselves, the rest to U+0100...
const std::vector<std::string>& byte_to_printable() {
  static const std::vector<std::string> table = [] {
    std::vector<std::string> t(256);
    int n = 0;
    for (int b = 0; b < 256; ++b) {
      const bool keep = (b >= 0x21 && b <= 0x7E) || (b >= 0xA1 && b <= 0xAC) || (b >= 0xAE && b <= 0xFF);
      append_utf8(t[b], keep ? static_cast<uint32_t>(b) : static_cast<uint32_t>(256 + n++));
    }
    return t;
  }();
  return table;
}
const std::unordered_map<std::string, unsigned char>& printable_to_byte() {
  static const std::unordered_map<std::string, unsigned char> map = [] {
    std::unordered_map<std::string, unsigned char> m;
    const auto& t = byte_to_printable();
    for (int b = 0; b < 256; ++b) m.emplace(t[b], static_cast<unsigned char>(b));
    return m;
  }();
  return map;
}
}  // namespace

std::string ByteBpeModel::load(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return "cannot open " + path;
  std::ostringstream ss;
  ss << f.rdbuf();
  return load_from_json(ss.str());
}

std::string ByteBpeModel::load_from_json(std::string_view json) {
  JsonParser parser{json.data(), json.data() + json.size()};
  const JValue root = parser.value();
  if (!parser.ok || root.type != JValue::kObject) return "not a JSON object";
  const JValue* model = root.get("model");
  const JValue* vocab = model ? model->get("vocab") : nullptr;
  const JValue* merges = model ? model->get("merges") : nullptr;
  if (!vocab || vocab->type != JValue::kObject || !merges || merges->type != JValue::kArray)
    return "tokenizer.json has no model.vocab / model.merges (not a BPE tokenizer)";
  if (const JValue* t = model->get("type"); t && t->type == JValue::kString && t->str != "BPE")
    return "model.type is " + t->str + ", expected BPE";
  if (const JValue* im = model->get("ignore_merges")) ignore_merges_ = im->type == JValue::kBool && im->b;

  vocab_.clear(), merge_rank_.clear(), added_.clear(), id_to_token_.clear(), id_is_added_.clear();
  int32_t max_id = -1;
  for (const auto& kv : vocab->obj) {
    if (kv.second.type != JValue::kNumber) return "non-numeric id in model.vocab";
    const int32_t id = static_cast<int32_t>(kv.second.num);
    vocab_.emplace(kv.first, id);
    max_id = std::max(max_id, id);
  }
  const JValue* added = root.get("added_tokens");
  if (added && added->type == JValue::kArray) {
    for (const JValue& a : added->arr) {
      const JValue* id = a.get("id");
      const JValue* content = a.get("content");
      if (!id || !content || id->type != JValue::kNumber || content->type != JValue::kString) continue;
      added_.emplace_back(content->str, static_cast<int32_t>(id->num));
      max_id = std::max(max_id, static_cast<int32_t>(id->num));
    }
  }
  if (max_id < 0) return "empty vocabulary";
  id_to_token_.assign(static_cast<size_t>(max_id) + 1, std::string());
  id_is_added_.assign(static_cast<size_t>(max_id) + 1, false);
  for (const auto& kv : vocab_)
    if (kv.second >= 0) id_to_token_[kv.second] = kv.first;
  for (const auto& a : added_)
    if (a.second >= 0) id_to_token_[a.second] = a.first, id_is_added_[a.second] = true;
  std::sort(added_.begin(), added_.end(),
            [](const auto& x, const auto& y) { return x.first.size() > y.first.size(); });

  int32_t rank = 0;
  for (const JValue& m : merges->arr) {  // ["a", "b"] (tokenizers >= 0.20) or "a b"
    std::string key;
    if (m.type == JValue::kArray && m.arr.size() == 2 && m.arr[0].type == JValue::kString &&
        m.arr[1].type == JValue::kString)
      key = m.arr[0].str + " " + m.arr[1].str;
    else if (m.type == JValue::kString)
      key = m.str;
    else
      return "malformed entry in model.merges";
    merge_rank_.emplace(std::move(key), rank++);
  }

  // \p{N}{1,3} (Llama-3) or \p{N} (Qwen2, and the reference's PAT_STR, encode.cpp:60-61)
  max_digits_ = 1;
  const std::string_view all(json);
  if (all.find("\\\\p{N}{1,3}") != std::string_view::npos) max_digits_ = 3;
  return "";
}

int32_t ByteBpeModel::token_to_id(const std::string& content) const {
  for (const auto& a : added_)
    if (a.first == content) return a.second;
  const auto it = vocab_.find(content);
  return it == vocab_.end() ? -1 : it->second;
}

// (?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+
std::vector<std::string> ByteBpeModel::split(std::string_view text) const {
  std::vector<std::string> pieces;
  const Cps u = decode_utf8(text);
  const size_t n = u.cp.size();
  auto at = [&](size_t i) -> uint32_t { return i < n ? u.cp[i] : 0xFFFFFFFFu; };
  auto lower = [](uint32_t c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; };
  size_t i = 0;
  while (i < n) {
    size_t j = i;  // end of the match (exclusive)
    const uint32_t c = u.cp[i];
    // 1. contractions
    if (c == '\'') {
      const uint32_t a = lower(at(i + 1)), b = lower(at(i + 2));
      if (a == 's' || a == 't' || a == 'm' || a == 'd') j = i + 2;
      else if ((a == 'r' && b == 'e') || (a == 'v' && b == 'e') || (a == 'l' && b == 'l')) j = i + 3;
      else if (a == 0x17F) j = i + 2;  // U+017F LONG S case-folds to 's'
    }
    // 2. [^\r\n\p{L}\p{N}]?\p{L}+
    if (j == i) {
      size_t k = i;
      if (!is_newline(c) && !is_letter(c) && !is_number(c) && is_letter(at(i + 1))) k = i + 1;
      if (is_letter(at(k))) {
        while (is_letter(at(k))) ++k;
        j = k;
      }
    }
    // 3. \p{N}{1,max}
    if (j == i && is_number(c)) {
      size_t k = i;
      while (k < i + static_cast<size_t>(max_digits_) && is_number(at(k))) ++k;
      j = k;
    }
    // 4.  ?[^\s\p{L}\p{N}]+[\r\n]*
    if (j == i) {
      auto other = [&](uint32_t x) { return x != 0xFFFFFFFFu && !is_space(x) && !is_letter(x) && !is_number(x); };
      size_t k = i;
      if (c == ' ' && other(at(i + 1))) k = i + 1;
      if (other(at(k))) {
        while (other(at(k))) ++k;
        while (is_newline(at(k))) ++k;
        j = k;
      }
    }
    // 5-7. whitespace
    if (j == i && is_space(c)) {
      size_t run = i;
      while (is_space(at(run))) ++run;
      size_t last_nl = SIZE_MAX;
      for (size_t k = i; k < run; ++k)
        if (is_newline(u.cp[k])) last_nl = k;
      if (last_nl != SIZE_MAX) j = last_nl + 1;           // \s*[\r\n]+
      else if (run == n) j = run;                         // \s+(?!\S) at the end of the text
      else if (run - i > 1) j = run - 1;                  // \s+(?!\S): leave one blank for the next word
      else j = run;                                       // \s+
    }
    if (j == i) j = i + 1;  // (unreachable: every code point starts some alternative)
    pieces.emplace_back(text.substr(u.off[i], u.off[j] - u.off[i]));
    i = j;
  }
  return pieces;
}

void ByteBpeModel::encode_piece(std::string_view piece, std::vector<int32_t>& out) const {
  const auto& b2p = byte_to_printable();
  std::vector<std::string> sym;
  std::string whole;
  for (unsigned char b : piece) {
    sym.push_back(b2p[b]);
    whole += b2p[b];
  }
  if (ignore_merges_) {
    const auto it = vocab_.find(whole);
    if (it != vocab_.end()) {
      out.push_back(it->second);
      return;
    }
  }
  while (sym.size() > 1) {
    int best_rank = INT_MAX;
    size_t best = 0;
    for (size_t k = 0; k + 1 < sym.size(); ++k) {
      const auto it = merge_rank_.find(sym[k] + " " + sym[k + 1]);
      if (it != merge_rank_.end() && it->second < best_rank) best_rank = it->second, best = k;
    }
    if (best_rank == INT_MAX) break;
    sym[best] += sym[best + 1];
    sym.erase(sym.begin() + static_cast<long>(best) + 1);
  }
  for (const std::string& s : sym) {
    const auto it = vocab_.find(s);
    if (it != vocab_.end()) out.push_back(it->second);
    // (a byte-level vocabulary always holds the 256 single-byte symbols; nothing to drop)
  }
}

std::vector<int32_t> ByteBpeModel::encode(std::string_view text) const {
  std::vector<int32_t> ids;
  size_t pos = 0;
  while (pos < text.size()) {
    // earliest added token at or after pos; on the same start the longest wins
    size_t hit = std::string_view::npos, hit_len = 0;
    int32_t hit_id = -1;
    for (const auto& a : added_) {
      if (a.first.empty()) continue;
      const size_t at = text.find(a.first, pos);
      if (at != std::string_view::npos && (at < hit || (at == hit && a.first.size() > hit_len)))
        hit = at, hit_len = a.first.size(), hit_id = a.second;
    }
    const size_t stop = hit == std::string_view::npos ? text.size() : hit;
    if (stop > pos)
      for (const std::string& piece : split(text.substr(pos, stop - pos))) encode_piece(piece, ids);
    if (hit == std::string_view::npos) break;
    ids.push_back(hit_id);
    pos = hit + hit_len;
  }
  return ids;
}

std::string ByteBpeModel::decode(const std::vector<int32_t>& ids) const {
  const auto& p2b = printable_to_byte();
  std::string out;
  for (int32_t id : ids) {
    if (id < 0 || id >= vocab_size()) continue;
    const std::string& tok = id_to_token_[id];
    if (id_is_added_[id]) {
      out += tok;  // literal content, as cpp-tiktoken's _decode_native does for special tokens
      continue;
    }
    const Cps u = decode_utf8(tok);
    for (size_t k = 0; k < u.cp.size(); ++k) {
      const auto it = p2b.find(tok.substr(u.off[k], u.off[k + 1] - u.off[k]));
      if (it != p2b.end()) out.push_back(static_cast<char>(it->second));
    }
  }
  return out;
}
}  // namespace op
