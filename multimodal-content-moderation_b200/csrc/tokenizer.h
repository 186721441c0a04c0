// CLIP byte-level BPE tokenizer (host side, C++), SURVEY 8f rank 4: the step in front of `input_ids` /
// `attention_mask` of the hot path.
//
// Replaces, for the CLIP backends, the reference's
//     tok = self.tok(text, padding="max_length", truncation=True, max_length=self.max_len, return_attention_mask=True)
// (R/src/data/dataset.py:148-165, R/scripts/inference.py:168-180, R/sagemaker/inference.py:230-239), where `self.tok` is
// Hugging Face's CLIPTokenizer.  The algorithm restated here is the pipeline that class configures
// (HF/models/clip/tokenization_clip.py:68-118, executed by the `tokenizers` library):
//     added special tokens "<|startoftext|>" / "<|endoftext|>" are cut out of the RAW text and mapped to their ids
//     normalizer      NFC -> every run of White_Space becomes one ' ' -> Unicode lower-casing, character by character
//                     (the `tokenizers` Lowercase normalizer has no final-sigma rule: capital sigma is always U+03C3)
//     pre-tokenizer   Split(<|startoftext|>|<|endoftext|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+),
//                     matches kept, the rest (white space) dropped; then ByteLevel, which (use_regex = true) re-splits
//                     every piece with GPT-2's pattern -- a no-op for all pieces but the two special-token literals,
//                     which fall apart into "<|", the letters and "|>": exactly what the Split produces without those
//                     two alternatives, so the scanner below leaves them out -- and maps UTF-8 bytes to the 256
//                     printable stand-in characters of GPT-2
//     model           BPE with end_of_word_suffix "</w>", unknown symbol -> unk (= "<|endoftext|>"), lowest-rank merge
//                     first, leftmost first among equals
//     post-processor  [bos] tokens [eos]; truncation to max_len (tokens cut to max_len - 2), padding with the pad token
//                     (= "<|endoftext|>") and attention_mask 1 / 0
// The Unicode data (csrc/unicode_tables.h) is probed from the `tokenizers` library itself, code point by code point
// (tools/gen_unicode_tables.py), so "identical to CLIPTokenizer" holds for the installed library version.
// The vocabulary (vocab.json) and the merge list (merges.txt) are the checkpoint's own files; none ship with this
// repository (the image has no network), the tests pin the algorithm against CLIPTokenizer on synthetic vocabularies.
//
// No CUDA here: tokenisation is string work on the host in the reference as well.  Batches are split over threads.
#pragma once

#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#include "unicode_tables.h"

namespace mmcm_tok {

// ------------------------------------------------------------------------------------------------ Unicode helpers
inline bool in_ranges(const uint32_t (*r)[2], int n, uint32_t cp) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (cp < r[mid][0]) hi = mid - 1;
    else if (cp > r[mid][1]) lo = mid + 1;
    else return true;
  }
  return false;
}
inline bool is_letter(uint32_t cp) {
  if (cp < 128) return (cp >= 'a' && cp <= 'z') || (cp >= 'A' && cp <= 'Z');
  return in_ranges(mmcm_uni::kLetter, mmcm_uni::kLetterCount, cp);
}
inline bool is_number(uint32_t cp) {
  if (cp < 128) return cp >= '0' && cp <= '9';
  return in_ranges(mmcm_uni::kNumber, mmcm_uni::kNumberCount, cp);
}
inline bool is_space(uint32_t cp) {
  if (cp < 128) return cp == ' ' || (cp >= 9 && cp <= 13);
  for (int i = 0; i < mmcm_uni::kSpaceCount; ++i)
    if (mmcm_uni::kSpace[i] == cp) return true;
  return false;
}

inline uint8_t combining_class(uint32_t cp) {
  if (cp < 0x300) return 0;
  int lo = 0, hi = mmcm_uni::kCccCount - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (cp < mmcm_uni::kCcc[mid].lo) hi = mid - 1;
    else if (cp > mmcm_uni::kCcc[mid].hi) lo = mid + 1;
    else return mmcm_uni::kCcc[mid].ccc;
  }
  return 0;
}
inline const mmcm_uni::Decomp* find_decomp(uint32_t cp) {
  int lo = 0, hi = mmcm_uni::kDecompCount - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (cp < mmcm_uni::kDecomp[mid].cp) hi = mid - 1;
    else if (cp > mmcm_uni::kDecomp[mid].cp) lo = mid + 1;
    else return &mmcm_uni::kDecomp[mid];
  }
  return nullptr;
}
inline uint32_t compose_pair(uint32_t a, uint32_t b) {
  // Hangul: L + V -> LV, LV + T -> LVT
  if (a >= 0x1100 && a < 0x1113 && b >= 0x1161 && b < 0x1176) return 0xAC00 + ((a - 0x1100) * 21 + (b - 0x1161)) * 28;
  if (a >= 0xAC00 && a < 0xD7A4 && (a - 0xAC00) % 28 == 0 && b > 0x11A7 && b < 0x11C3) return a + (b - 0x11A7);
  int lo = 0, hi = mmcm_uni::kCompCount - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const mmcm_uni::Comp& c = mmcm_uni::kComp[mid];
    if (a < c.a || (a == c.a && b < c.b)) hi = mid - 1;
    else if (a > c.a || (a == c.a && b > c.b)) lo = mid + 1;
    else return c.cp;
  }
  return 0;
}
inline void decompose(uint32_t cp, std::vector<uint32_t>& out) {
  if (cp >= 0xAC00 && cp < 0xD7A4) {   // Hangul syllable
    const uint32_t s = cp - 0xAC00;
    out.push_back(0x1100 + s / 588);
    out.push_back(0x1161 + (s % 588) / 28);
    if (s % 28) out.push_back(0x11A7 + s % 28);
    return;
  }
  const mmcm_uni::Decomp* d = cp >= 0xC0 ? find_decomp(cp) : nullptr;
  if (!d) { out.push_back(cp); return; }
  for (int k = 0; k < 4 && d->to[k]; ++k) out.push_back(d->to[k]);   // the table holds the FULL decomposition
}

// Unicode Normalization Form C (UAX #15): canonical decomposition, canonical ordering, canonical composition
inline void nfc(std::vector<uint32_t>& s) {
  bool plain = true;                     // everything below U+0300 is already in NFC
  for (uint32_t c : s)
    if (c >= 0x300) { plain = false; break; }
  if (plain) return;
  std::vector<uint32_t> d;
  d.reserve(s.size() + 8);
  for (uint32_t c : s) decompose(c, d);
  for (size_t i = 1; i < d.size(); ++i) {          // canonical ordering: stable sort of runs of non-starters by ccc
    const uint8_t cc = combining_class(d[i]);
    if (!cc) continue;
    size_t j = i;
    while (j > 0) {
      const uint8_t pc = combining_class(d[j - 1]);
      if (pc <= cc) break;
      std::swap(d[j], d[j - 1]);
      --j;
    }
  }
  std::vector<uint32_t> o;
  o.reserve(d.size());
  size_t starter = (size_t)-1;
  uint8_t last_cc = 0;
  for (size_t i = 0; i < d.size(); ++i) {
    const uint32_t c = d[i];
    const uint8_t cc = combining_class(c);
    if (starter != (size_t)-1) {
      const bool blocked = (o.size() - 1 > starter) && (last_cc == 0 || last_cc >= cc);   // something uncombined in between
      if (!blocked) {
        const uint32_t comp = compose_pair(o[starter], c);
        if (comp) { o[starter] = comp; continue; }
      }
    }
    if (cc == 0) { starter = o.size(); last_cc = 0; }
    else last_cc = cc;
    o.push_back(c);
  }
  s.swap(o);
}

inline const mmcm_uni::Lower* find_lower(uint32_t cp) {
  int lo = 0, hi = mmcm_uni::kLowerCount - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (cp < mmcm_uni::kLower[mid].cp) hi = mid - 1;
    else if (cp > mmcm_uni::kLower[mid].cp) lo = mid + 1;
    else return &mmcm_uni::kLower[mid];
  }
  return nullptr;
}
// full lower-casing of every character on its own (char::to_lowercase of the `tokenizers` Lowercase normalizer: one
// character may become up to three, e.g. U+0130 -> "i" + U+0307; no context rules, so no final sigma)
inline void lowercase(const std::vector<uint32_t>& s, std::vector<uint32_t>& out) {
  out.clear();
  out.reserve(s.size());
  for (size_t i = 0; i < s.size(); ++i) {
    const uint32_t c = s[i];
    if (c < 128) { out.push_back((c >= 'A' && c <= 'Z') ? c + 32 : c); continue; }
    const mmcm_uni::Lower* l = find_lower(c);
    if (!l) { out.push_back(c); continue; }
    for (int k = 0; k < 3 && l->to[k]; ++k) out.push_back(l->to[k]);
  }
}

inline void utf8_decode(const char* p, size_t n, std::vector<uint32_t>& out) {
  out.clear();
  const unsigned char* s = reinterpret_cast<const unsigned char*>(p);
  for (size_t i = 0; i < n;) {
    uint32_t c = s[i];
    int len = 1;
    if (c < 0x80) {}
    else if ((c >> 5) == 6 && i + 1 < n) { c = ((c & 31) << 6) | (s[i + 1] & 63); len = 2; }
    else if ((c >> 4) == 14 && i + 2 < n) { c = ((c & 15) << 12) | ((s[i + 1] & 63) << 6) | (s[i + 2] & 63); len = 3; }
    else if ((c >> 3) == 30 && i + 3 < n) { c = ((c & 7) << 18) | ((s[i + 1] & 63) << 12) | ((s[i + 2] & 63) << 6) | (s[i + 3] & 63); len = 4; }
    else c = 0xFFFD;
    out.push_back(c);
    i += len;
  }
}
inline void utf8_append(std::string& o, uint32_t c) {
  if (c < 0x80) o.push_back((char)c);
  else if (c < 0x800) { o.push_back((char)(0xC0 | (c >> 6))); o.push_back((char)(0x80 | (c & 63))); }
  else if (c < 0x10000) { o.push_back((char)(0xE0 | (c >> 12))); o.push_back((char)(0x80 | ((c >> 6) & 63))); o.push_back((char)(0x80 | (c & 63))); }
  else { o.push_back((char)(0xF0 | (c >> 18))); o.push_back((char)(0x80 | ((c >> 12) & 63))); o.push_back((char)(0x80 | ((c >> 6) & 63))); o.push_back((char)(0x80 | (c & 63))); }
}

// ------------------------------------------------------------------------------------------------ the tokenizer
struct PairHash {
  size_t operator()(const std::pair<int32_t, int32_t>& p) const {
    return std::hash<uint64_t>()(((uint64_t)(uint32_t)p.first << 32) | (uint32_t)p.second);
  }
};

class ClipTokenizer {
 public:
  // returns "" on success, else the error message
  std::string load(const char* vocab_path, const char* merges_path) {
    std::string err = load_vocab(vocab_path);
    if (!err.empty()) return err;
    auto b = vocab_.find(kBos), e = vocab_.find(kEos);
    if (b == vocab_.end() || e == vocab_.end()) return "vocabulary lacks <|startoftext|> / <|endoftext|>";
    bos_ = b->second;
    eos_ = e->second;
    // GPT-2's byte <-> printable character table (HF pre_tokenizers.ByteLevel)
    int n = 0;
    for (int byte = 0; byte < 256; ++byte) {
      const bool printable = (byte >= '!' && byte <= '~') || (byte >= 0xA1 && byte <= 0xAC) || (byte >= 0xAE && byte <= 0xFF);
      const uint32_t cp = printable ? (uint32_t)byte : (uint32_t)(256 + n++);
      std::string ch;
      utf8_append(ch, cp);
      byte_sym_[byte] = lookup(ch);
      byte_sym_end_[byte] = lookup(ch + "</w>");
    }
    return load_merges(merges_path);
  }

  int vocab_size() const { return (int)vocab_.size(); }
  int bos() const { return bos_; }
  int eos() const { return eos_; }

  // ids / mask: max_len entries each
  void encode(const char* text, size_t len, int max_len, int64_t* ids, int64_t* mask,
              std::unordered_map<std::string, std::vector<int32_t>>& cache) const {
    std::vector<int32_t> toks;
    toks.reserve(96);
    // added special tokens are matched on the raw text, before normalisation
    size_t pos = 0;
    while (pos <= len) {
      size_t next = len, slen = 0;
      int sid = -1;
      for (int k = 0; k < 2; ++k) {
        const std::string& sp = k == 0 ? kBos : kEos;
        if (sp.size() > len - pos) continue;
        const char* f = std::search(text + pos, text + len, sp.begin(), sp.end());
        if (f != text + len && (size_t)(f - text) < next) { next = f - text; slen = sp.size(); sid = k == 0 ? bos_ : eos_; }
      }
      if (next > pos) encode_segment(text + pos, next - pos, toks, cache);
      if (sid < 0) break;
      toks.push_back(sid);
      pos = next + slen;
    }
    const int body = std::min((int)toks.size(), std::max(0, max_len - 2));
    int o = 0;
    if (max_len >= 1) { ids[o] = bos_; mask[o++] = 1; }
    for (int i = 0; i < body && o < max_len; ++i) { ids[o] = toks[i]; mask[o++] = 1; }
    if (o < max_len) { ids[o] = eos_; mask[o++] = 1; }
    for (; o < max_len; ++o) { ids[o] = eos_; mask[o] = 0; }   // pad_token == "<|endoftext|>"
  }

 private:
  const std::string kBos = "<|startoftext|>", kEos = "<|endoftext|>";
  std::unordered_map<std::string, int32_t> vocab_;
  std::vector<std::string> id_to_tok_;
  std::unordered_map<std::pair<int32_t, int32_t>, std::pair<int32_t, int32_t>, PairHash> merges_;   // (a, b) -> (rank, id)
  int32_t bos_ = 0, eos_ = 0;
  int32_t byte_sym_[256], byte_sym_end_[256];

  int32_t lookup(const std::string& s) const {
    auto it = vocab_.find(s);
    return it == vocab_.end() ? -1 : it->second;
  }

  // normalise + pre-tokenise one stretch of text between special tokens
  void encode_segment(const char* p, size_t n, std::vector<int32_t>& toks,
                      std::unordered_map<std::string, std::vector<int32_t>>& cache) const {
    std::vector<uint32_t> raw, ws, s;
    utf8_decode(p, n, raw);
    nfc(raw);
    ws.reserve(raw.size());
    for (size_t i = 0; i < raw.size();) {            // \s+ -> ' '
      if (is_space(raw[i])) {
        while (i < raw.size() && is_space(raw[i])) ++i;
        ws.push_back(' ');
      } else ws.push_back(raw[i++]);
    }
    lowercase(ws, s);
    // the Split regex, alternative by alternative, leftmost match, first alternative that matches
    static const char* kContr[] = {"'s", "'t", "'re", "'ve", "'m", "'ll", "'d"};
    const size_t L = s.size();
    auto match_ascii = [&](size_t i, const char* lit) -> size_t {
      size_t k = 0;
      for (; lit[k]; ++k)
        if (i + k >= L || s[i + k] != (uint32_t)(unsigned char)lit[k]) return 0;
      return k;
    };
    std::string word;
    for (size_t i = 0; i < L;) {
      size_t m = 0;
      if (s[i] == '\'') for (const char* c : kContr) if ((m = match_ascii(i, c))) break;
      if (!m) {
        if (is_letter(s[i])) { m = 1; while (i + m < L && is_letter(s[i + m])) ++m; }
        else if (is_number(s[i])) m = 1;
        else if (!is_space(s[i])) { m = 1; while (i + m < L && !is_space(s[i + m]) && !is_letter(s[i + m]) && !is_number(s[i + m])) ++m; }
      }
      if (!m) { ++i; continue; }                     // white space between matches: dropped
      word.clear();
      for (size_t k = 0; k < m; ++k) utf8_append(word, s[i + k]);
      bpe(word, toks, cache);
      i += m;
    }
  }

  // one pre-token (UTF-8) -> token ids
  void bpe(const std::string& word, std::vector<int32_t>& toks,
           std::unordered_map<std::string, std::vector<int32_t>>& cache) const {
    auto hit = cache.find(word);
    if (hit != cache.end()) { toks.insert(toks.end(), hit->second.begin(), hit->second.end()); return; }
    std::vector<int32_t> sym(word.size());
    for (size_t i = 0; i < word.size(); ++i) {
      const unsigned char b = (unsigned char)word[i];
      const int32_t id = (i + 1 == word.size()) ? byte_sym_end_[b] : byte_sym_[b];
      sym[i] = id >= 0 ? id : eos_;                  // unknown symbol -> unk_token (= "<|endoftext|>")
    }
    while (sym.size() > 1) {
      int best = -1, best_rank = INT32_MAX, best_id = -1;
      for (size_t i = 0; i + 1 < sym.size(); ++i) {
        auto it = merges_.find({sym[i], sym[i + 1]});
        if (it != merges_.end() && it->second.first < best_rank) { best_rank = it->second.first; best = (int)i; best_id = it->second.second; }
      }
      if (best < 0) break;
      sym[best] = best_id;
      sym.erase(sym.begin() + best + 1);
    }
    if (cache.size() < (1u << 16)) cache.emplace(word, sym);
    toks.insert(toks.end(), sym.begin(), sym.end());
  }

  static bool read_file(const char* path, std::string& out) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out.append(buf, n);
    fclose(f);
    return true;
  }

  // vocab.json: one flat JSON object {"token": id, ...}
  std::string load_vocab(const char* path) {
    std::string js;
    if (!read_file(path, js)) return std::string("cannot read vocabulary file '") + path + "'";
    size_t i = 0;
    auto skip = [&]() { while (i < js.size() && (js[i] == ' ' || js[i] == '\n' || js[i] == '\r' || js[i] == '\t')) ++i; };
    skip();
    if (i >= js.size() || js[i] != '{') return "vocab.json: expected '{'";
    ++i;
    while (true) {
      skip();
      if (i < js.size() && js[i] == '}') break;
      if (i >= js.size() || js[i] != '"') return "vocab.json: expected a string key";
      ++i;
      std::string key;
      while (i < js.size() && js[i] != '"') {
        if (js[i] != '\\') { key.push_back(js[i++]); continue; }
        if (++i >= js.size()) return "vocab.json: bad escape";
        const char e = js[i++];
        if (e == 'u') {
          auto hex4 = [&](uint32_t& v) { if (i + 4 > js.size()) return false; v = (uint32_t)strtoul(js.substr(i, 4).c_str(), nullptr, 16); i += 4; return true; };
          uint32_t cp, lo;
          if (!hex4(cp)) return "vocab.json: bad \\u escape";
          if (cp >= 0xD800 && cp < 0xDC00 && i + 6 <= js.size() && js[i] == '\\' && js[i + 1] == 'u') {
            i += 2;
            if (!hex4(lo)) return "vocab.json: bad surrogate pair";
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          }
          utf8_append(key, cp);
        } else {
          const char* m = strchr("\"\"\\\\//b\bf\fn\nr\rt\t", e);
          if (!m) return "vocab.json: unknown escape";
          key.push_back(m[1]);
        }
      }
      ++i;
      skip();
      if (i >= js.size() || js[i] != ':') return "vocab.json: expected ':'";
      ++i;
      skip();
      char* end = nullptr;
      const long id = strtol(js.c_str() + i, &end, 10);
      if (end == js.c_str() + i || id < 0) return "vocab.json: expected a non-negative integer id";
      i = end - js.c_str();
      vocab_[key] = (int32_t)id;
      skip();
      if (i < js.size() && js[i] == ',') ++i;
    }
    return "";
  }

  // merges.txt: "#version" header, then one "left right" pair per line, rank = line order
  std::string load_merges(const char* path) {
    std::string txt;
    if (!read_file(path, txt)) return std::string("cannot read merges file '") + path + "'";
    int rank = 0;
    size_t i = 0;
    while (i < txt.size()) {
      size_t e = txt.find('\n', i);
      if (e == std::string::npos) e = txt.size();
      std::string line = txt.substr(i, e - i);
      i = e + 1;
      if (!line.empty() && line.back() == '\r') line.pop_back();
      if (line.empty() || line.rfind("#version", 0) == 0) continue;
      const size_t sp = line.find(' ');
      if (sp == std::string::npos) return "merges.txt: line without a separator: '" + line + "'";
      const std::string a = line.substr(0, sp), b = line.substr(sp + 1);
      const int32_t ia = lookup(a), ib = lookup(b), iab = lookup(a + b);
      if (ia < 0 || ib < 0 || iab < 0) return "merges.txt: token of merge '" + line + "' is not in the vocabulary";
      merges_.emplace(std::make_pair(ia, ib), std::make_pair(rank, iab));
      ++rank;
    }
    return "";
  }
};

// texts[i] .. texts[i] + lens[i]; ids / mask: [n, max_len] int64
inline void encode_batch(const ClipTokenizer& tok, const char* const* texts, const int64_t* lens, int n, int max_len,
                         int64_t* ids, int64_t* mask, int n_threads) {
  if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
  n_threads = std::max(1, std::min(n_threads, (n + 63) / 64));
  auto work = [&](int t) {
    std::unordered_map<std::string, std::vector<int32_t>> cache;
    for (int i = t; i < n; i += n_threads)
      tok.encode(texts[i], (size_t)lens[i], max_len, ids + (size_t)i * max_len, mask + (size_t)i * max_len, cache);
  };
  if (n_threads == 1) { work(0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
  for (auto& x : th) x.join();
}

}  // namespace mmcm_tok
