// SentencePiece BPE model reader / encoder / decoder -- see op/spm_bpe.h.
#include "op/spm_bpe.h"

#include <cstring>
#include <fstream>
#include <queue>
#include <sstream>

namespace op {
namespace {
constexpr const char kSpace[] = "\xE2\x96\x81";  // U+2581, SentencePiece's visible blank

// ---- protobuf wire format (just what ModelProto needs) ---------------------------------------------
struct Reader {
  const unsigned char* p;
  const unsigned char* end;
  bool ok = true;
  bool done() const { return p >= end || !ok; }
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 64; shift += 7) {
      if (p >= end) break;
      const unsigned char b = *p++;
      v |= static_cast<uint64_t>(b & 0x7F) << shift;
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
  std::string_view bytes() {
    const uint64_t n = varint();
    if (!ok || n > static_cast<uint64_t>(end - p)) {
      ok = false;
      return {};
    }
    std::string_view s(reinterpret_cast<const char*>(p), static_cast<size_t>(n));
    p += n;
    return s;
  }
  uint32_t fixed32() {
    if (end - p < 4) {
      ok = false;
      return 0;
    }
    uint32_t v;
    std::memcpy(&v, p, 4);
    p += 4;
    return v;
  }
  void skip(int wire) {
    switch (wire) {
      case 0: varint(); break;
      case 1: if (end - p < 8) ok = false; else p += 8; break;
      case 2: bytes(); break;
      case 5: fixed32(); break;
      default: ok = false;
    }
  }
};
Reader reader_of(std::string_view s) {
  return Reader{reinterpret_cast<const unsigned char*>(s.data()),
                reinterpret_cast<const unsigned char*>(s.data()) + s.size()};
}

// length of the UTF-8 character starting at s[0] (1 for a malformed lead byte)
int utf8_len(std::string_view s) {
  const unsigned char c = static_cast<unsigned char>(s[0]);
  int n = c < 0x80 ? 1 : (c >> 5) == 0x6 ? 2 : (c >> 4) == 0xE ? 3 : (c >> 3) == 0x1E ? 4 : 1;
  if (n > static_cast<int>(s.size())) return 1;
  for (int i = 1; i < n; ++i)
    if ((static_cast<unsigned char>(s[i]) & 0xC0) != 0x80) return 1;
  return n;
}
bool valid_utf8_char(std::string_view s, int n) {
  const unsigned char c = static_cast<unsigned char>(s[0]);
  if (n == 1) return c < 0x80;
  uint32_t cp = n == 2 ? (c & 0x1F) : n == 3 ? (c & 0x0F) : (c & 0x07);
  for (int i = 1; i < n; ++i) cp = (cp << 6) | (static_cast<unsigned char>(s[i]) & 0x3F);
  if (n == 2) return cp >= 0x80;
  if (n == 3) return cp >= 0x800 && !(cp >= 0xD800 && cp <= 0xDFFF);
  return cp >= 0x10000 && cp <= 0x10FFFF;
}
}  // namespace

std::string SpmBpeModel::load(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return "cannot open " + path;
  std::ostringstream ss;
  ss << f.rdbuf();
  return load_from_bytes(ss.str());
}

std::string SpmBpeModel::load_from_bytes(std::string_view blob) {
  pieces_.clear();
  merge_vocab_.clear();
  int model_type = 1;  // TrainerSpec default: UNIGRAM
  bool has_charsmap = false;
  Reader r = reader_of(blob);
  while (!r.done()) {
    const uint64_t key = r.varint();
    const int field = static_cast<int>(key >> 3), wire = static_cast<int>(key & 7);
    if (field == 1 && wire == 2) {  // repeated SentencePiece pieces
      Reader m = reader_of(r.bytes());
      Piece pc;
      while (!m.done()) {
        const uint64_t k = m.varint();
        const int fd = static_cast<int>(k >> 3), w = static_cast<int>(k & 7);
        if (fd == 1 && w == 2) pc.text = std::string(m.bytes());
        else if (fd == 2 && w == 5) { const uint32_t u = m.fixed32(); std::memcpy(&pc.score, &u, 4); }
        else if (fd == 3 && w == 0) pc.type = static_cast<int>(m.varint());
        else m.skip(w);
      }
      if (!m.ok) return "malformed SentencePiece entry";
      pieces_.push_back(std::move(pc));
    } else if (field == 2 && wire == 2) {  // TrainerSpec
      Reader m = reader_of(r.bytes());
      while (!m.done()) {
        const uint64_t k = m.varint();
        const int fd = static_cast<int>(k >> 3), w = static_cast<int>(k & 7);
        if (w == 0 && (fd == 3 || fd == 35 || (fd >= 40 && fd <= 43))) {
          const int64_t v = static_cast<int64_t>(m.varint());
          if (fd == 3) model_type = static_cast<int>(v);
          else if (fd == 35) byte_fallback_ = v != 0;
          else if (fd == 40) unk_id_ = static_cast<int32_t>(v);
          else if (fd == 41) bos_id_ = static_cast<int32_t>(v);
          else if (fd == 42) eos_id_ = static_cast<int32_t>(v);
        } else if (fd == 44 && w == 2) {
          unk_surface_ = std::string(m.bytes());
        } else {
          m.skip(w);
        }
      }
      if (!m.ok) return "malformed TrainerSpec";
    } else if (field == 3 && wire == 2) {  // NormalizerSpec
      Reader m = reader_of(r.bytes());
      while (!m.done()) {
        const uint64_t k = m.varint();
        const int fd = static_cast<int>(k >> 3), w = static_cast<int>(k & 7);
        if (fd == 2 && w == 2) has_charsmap = !m.bytes().empty();
        else if (fd == 3 && w == 0) add_dummy_prefix_ = m.varint() != 0;
        else if (fd == 4 && w == 0) remove_extra_whitespaces_ = m.varint() != 0;
        else if (fd == 5 && w == 0) escape_whitespaces_ = m.varint() != 0;
        else m.skip(w);
      }
      if (!m.ok) return "malformed NormalizerSpec";
    } else {
      r.skip(wire);
    }
  }
  if (!r.ok || pieces_.empty()) return "not a SentencePiece model file";
  if (model_type != 2) return "only BPE SentencePiece models are supported (this one is model_type " +
                              std::to_string(model_type) + ")";
  if (has_charsmap)
    return "the model carries a compiled normalisation map (e.g. nmt_nfkc); only the identity rule "
           "used by Llama tokenizers is supported";
  for (int b = 0; b < 256; ++b) byte_piece_[b] = -1;
  for (int32_t id = 0; id < static_cast<int32_t>(pieces_.size()); ++id) {
    const Piece& pc = pieces_[id];
    if (pc.type == kNormal || pc.type == kUserDefined || pc.type == kUnused) {
      merge_vocab_.emplace(std::string_view(pc.text), id);  // first definition wins
      if (pc.type == kUserDefined) has_user_defined_ = true;
    } else if (pc.type == kByte && pc.text.size() == 6 && pc.text.compare(0, 3, "<0x") == 0) {
      byte_piece_[std::stoi(pc.text.substr(3, 2), nullptr, 16)] = id;
    }
  }
  if (has_user_defined_) return "models with user-defined symbols are not supported";
  if (byte_fallback_)
    for (int b = 0; b < 256; ++b)
      if (byte_piece_[b] < 0) return "byte_fallback model without all 256 byte pieces";
  return "";
}

// normalizer.cc semantics for a spec without a character map (rule "identity")
std::string SpmBpeModel::normalize(std::string_view text) const {
  std::string out;
  if (text.empty()) return out;
  size_t begin = 0, end = text.size();
  if (remove_extra_whitespaces_) {
    while (begin < end && text[begin] == ' ') ++begin;
    while (end > begin && text[end - 1] == ' ') --end;
    if (begin == end) return out;
  }
  const std::string_view blank = escape_whitespaces_ ? std::string_view(kSpace) : std::string_view(" ");
  if (add_dummy_prefix_) out.append(blank);
  bool prev_space = false;
  for (size_t i = begin; i < end;) {
    const std::string_view rest = text.substr(i, end - i);
    const int n = utf8_len(rest);
    if (n == 1 && rest[0] == ' ') {
      if (!(remove_extra_whitespaces_ && prev_space)) out.append(blank);
      prev_space = true;
    } else {
      if (valid_utf8_char(rest, n)) out.append(rest.substr(0, n));
      else out.append("\xEF\xBF\xBD");  // malformed byte -> U+FFFD, one byte consumed
      prev_space = false;
    }
    i += valid_utf8_char(rest, n) ? n : 1;
  }
  return out;
}

std::vector<int32_t> SpmBpeModel::encode(std::string_view text) const {
  std::vector<int32_t> ids;
  const std::string norm = normalize(text);
  if (norm.empty()) return ids;

  struct Symbol {
    int prev, next;
    size_t off, len;  // span of `norm`; len 0 = merged away
  };
  std::vector<Symbol> sym;
  for (size_t i = 0; i < norm.size();) {
    const int n = utf8_len(std::string_view(norm).substr(i));
    sym.push_back({static_cast<int>(sym.size()) - 1, -1, i, static_cast<size_t>(n)});
    i += n;
  }
  for (size_t i = 0; i + 1 < sym.size(); ++i) sym[i].next = static_cast<int>(i) + 1;

  struct Pair {
    int left, right;
    float score;
    size_t size;
  };
  auto worse = [](const Pair& a, const Pair& b) {  // bpe_model.cc SymbolPairComparator
    return a.score < b.score || (a.score == b.score && a.left > b.left);
  };
  std::priority_queue<Pair, std::vector<Pair>, decltype(worse)> agenda(worse);
  auto consider = [&](int left, int right) {
    if (left < 0 || right < 0) return;
    const std::string_view cat(norm.data() + sym[left].off, sym[left].len + sym[right].len);
    const auto it = merge_vocab_.find(cat);
    if (it == merge_vocab_.end()) return;
    agenda.push({left, right, pieces_[it->second].score, cat.size()});
  };
  for (size_t i = 1; i < sym.size(); ++i) consider(static_cast<int>(i) - 1, static_cast<int>(i));
  while (!agenda.empty()) {
    const Pair top = agenda.top();
    agenda.pop();
    Symbol& l = sym[top.left];
    Symbol& r = sym[top.right];
    if (l.len == 0 || r.len == 0 || l.len + r.len != top.size) continue;  // stale entry
    l.len += r.len;
    l.next = r.next;
    if (r.next >= 0) sym[r.next].prev = top.left;
    r.len = 0;
    consider(l.prev, top.left);
    consider(top.left, l.next);
  }
  for (int i = 0; i != -1; i = sym[i].next) {
    const std::string_view piece(norm.data() + sym[i].off, sym[i].len);
    const auto it = merge_vocab_.find(piece);
    // (UNUSED pieces would be re-split along their merge history; Llama vocabularies have none)
    if (it != merge_vocab_.end() && pieces_[it->second].type != kUnused) {
      ids.push_back(it->second);
    } else if (byte_fallback_) {
      for (unsigned char b : piece) ids.push_back(byte_piece_[b]);
    } else if (ids.empty() || ids.back() != unk_id_) {
      ids.push_back(unk_id_);  // a run of unknown characters is ONE <unk> (sentencepiece_processor.cc)
    }
  }
  return ids;
}

std::string SpmBpeModel::decode(const std::vector<int32_t>& ids) const {
  std::string text, bytes;
  auto flush_bytes = [&]() {  // a run of byte pieces is one UTF-8 string; bad bytes -> U+FFFD each
    for (size_t i = 0; i < bytes.size();) {
      const std::string_view rest = std::string_view(bytes).substr(i);
      const int n = utf8_len(rest);
      if (valid_utf8_char(rest, n)) {
        text.append(rest.substr(0, n));
        i += n;
      } else {
        text.append("\xEF\xBF\xBD");
        i += 1;
      }
    }
    bytes.clear();
  };
  // sentencepiece_processor.cc Decode: the blank the normaliser put in front of the text is not part
  // of it.  With add_dummy_prefix one leading blank is consumed; with remove_extra_whitespaces every
  // leading blank is, until some text has been produced.
  bool at_start = true;
  for (int32_t id : ids) {
    if (id < 0 || id >= piece_size()) continue;
    const Piece& pc = pieces_[id];
    if (pc.type == kByte) {
      bytes.push_back(static_cast<char>(std::stoi(pc.text.substr(3, 2), nullptr, 16)));
      at_start = false;
      continue;
    }
    flush_bytes();
    if (pc.type == kControl) continue;
    if (pc.type == kUnknown) {
      text.append(unk_surface_);
      at_start = false;
      continue;
    }
    std::string_view s(pc.text);
    bool consumed = false;
    if (at_start && (add_dummy_prefix_ || remove_extra_whitespaces_) && s.substr(0, 3) == kSpace) {
      s.remove_prefix(3);
      consumed = !remove_extra_whitespaces_;
    }
    while (!s.empty()) {
      if (s.substr(0, 3) == kSpace) {
        text.push_back(' ');
        s.remove_prefix(3);
      } else {
        text.push_back(s[0]);
        s.remove_prefix(1);
      }
    }
    if (consumed || !text.empty()) at_start = false;
  }
  flush_bytes();
  return text;
}
}  // namespace op
