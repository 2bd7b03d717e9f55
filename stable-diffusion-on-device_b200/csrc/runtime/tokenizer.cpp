// CLIP byte-pair tokenizer — see tokenizer.h.  Restated from the behaviour of the reference's libsdod::Tokenizer
// (csrc/libsdod/src/tokenizer.cpp), stage by stage; each stage cites the lines it follows.  No locale calls, no wide-character library.
#include "tokenizer.h"

#include <cstdio>
#include <fstream>

namespace sdod {

namespace {

struct CpRange { uint32_t lo, hi; };
#include "unicode_tables.inc"

template <size_t N>
bool in_ranges(const CpRange (&t)[N], uint32_t cp) {
    size_t a = 0, b = N;
    while (a < b) {
        const size_t m = (a + b) / 2;
        if (cp < t[m].lo) b = m;
        else if (cp > t[m].hi) a = m + 1;
        else return true;
    }
    return false;
}
bool is_blank(uint32_t cp) { return in_ranges(kBlankRanges, cp); }     // iswblank
bool is_letter(uint32_t cp) { return in_ranges(kAlphaRanges, cp); }    // iswalpha (glibc counts non-ASCII digits as letters)
bool is_digit(uint32_t cp) { return in_ranges(kDigitRanges, cp); }     // iswdigit: ASCII 0-9 only

// One UTF-8 sequence at s[0..n): its length, or -1 when glibc's mbtowc fails under a UTF-8 locale — stray / truncated / overlong sequences
// and surrogates.  Like glibc (probed in this image, 2.39) it still accepts the pre-RFC-3629 forms: lead bytes 0xF5..0xFD, i.e. 4- to 6-byte
// sequences up to 0x7FFFFFFF; such code points belong to no character class and tokenize as "other" symbols, byte by byte.
int utf8_decode(const unsigned char* s, size_t n, uint32_t* cp) {
    if (n == 0) return -1;
    const unsigned char c = s[0];
    if (c < 0x80) { *cp = c; return 1; }
    int len;
    uint32_t v, min;
    if (c >= 0xC2 && c <= 0xDF) { len = 2; v = c & 0x1F; min = 0x80; }
    else if (c >= 0xE0 && c <= 0xEF) { len = 3; v = c & 0x0F; min = 0x800; }
    else if (c >= 0xF0 && c <= 0xF7) { len = 4; v = c & 0x07; min = 0x10000; }
    else if (c >= 0xF8 && c <= 0xFB) { len = 5; v = c & 0x03; min = 0x200000; }
    else if (c >= 0xFC && c <= 0xFD) { len = 6; v = c & 0x01; min = 0x4000000; }
    else return -1;
    if (n < static_cast<size_t>(len)) return -1;
    for (int i = 1; i < len; ++i) {
        if ((s[i] & 0xC0) != 0x80) return -1;
        v = (v << 6) | (s[i] & 0x3F);
    }
    if (v < min || (v >= 0xD800 && v <= 0xDFFF)) return -1;
    *cp = v;
    return len;
}

void utf8_append(std::string& out, uint32_t cp) {
    if (cp < 0x80) out.push_back(static_cast<char>(cp));
    else if (cp < 0x800) { out.push_back(static_cast<char>(0xC0 | (cp >> 6))); out.push_back(static_cast<char>(0x80 | (cp & 0x3F))); }
    else if (cp < 0x10000) {
        out.push_back(static_cast<char>(0xE0 | (cp >> 12))); out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
        out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
    } else {
        out.push_back(static_cast<char>(0xF0 | (cp >> 18))); out.push_back(static_cast<char>(0x80 | ((cp >> 12) & 0x3F)));
        out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F))); out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
    }
}

[[noreturn]] void bad_utf8() { throw TokenizerError("Invalid UTF-8 string"); }      // tokenizer.cpp:77,179

// CLIP's bytes_to_unicode: printable Latin-1 bytes map to themselves, the other 68 bytes to U+0100.. in byte order.
uint32_t byte_to_codepoint(unsigned b) {
    auto kept = [](unsigned x) { return (x >= 0x21 && x <= 0x7E) || (x >= 0xA1 && x <= 0xAC) || (x >= 0xAE); };
    if (kept(b)) return b;
    unsigned n = 0;
    for (unsigned x = 0; x < b; ++x) n += kept(x) ? 0 : 1;
    return 256 + n;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ stages
// Leading / trailing blanks dropped, runs of blanks -> one ASCII space, ASCII upper case -> lower case.  (The reference lower-cases with the
// one-argument C tolower on the decoded wide character — tokenizer.cpp:86 — which under a UTF-8 locale maps A-Z only; checked against the
// compiled reference: "É" keeps its bytes.)
std::string Tokenizer::sanitize(const std::string& str) {
    std::string out;
    out.reserve(str.size());
    const unsigned char* p = reinterpret_cast<const unsigned char*>(str.data());
    size_t left = str.size();
    bool seen_char = false, last_blank = false;
    while (left) {
        uint32_t cp;
        const int len = utf8_decode(p, left, &cp);
        if (len < 0) bad_utf8();
        if (cp == 0) break;                                   // mbtowc returns 0 on NUL: the reference stops there (tokenizer.cpp:78-81)
        const bool blank = is_blank(cp);
        if (!blank) {
            seen_char = true;
            if (cp >= 'A' && cp <= 'Z') out.push_back(static_cast<char>(cp + 32));
            else out.append(reinterpret_cast<const char*>(p), static_cast<size_t>(len));
        } else if (seen_char && !last_blank) {
            out.push_back(' ');
        }
        last_blank = blank;
        p += len;
        left -= static_cast<size_t>(len);
    }
    if (seen_char && last_blank) out.pop_back();
    return out;
}

// The CLIP pattern 's|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+ as the reference's hand-written scanner implements it
// (tokenizer.cpp:114-222): contractions are tried first at every token start, a digit is a token of its own, letters and "other"
// characters form maximal runs, blanks separate.
std::vector<std::string> Tokenizer::split_words(const std::string& s) {
    std::vector<std::string> words;
    const unsigned char* p = reinterpret_cast<const unsigned char*>(s.data());
    const unsigned char* const end = p + s.size();
    while (p < end) {
        const size_t rem = static_cast<size_t>(end - p);
        if (rem > 1 && p[0] == '\'') {
            size_t n = 0;
            if (p[1] == 's' || p[1] == 't' || p[1] == 'm' || p[1] == 'd') n = 2;
            else if (rem > 2 && ((p[1] == 'r' && p[2] == 'e') || (p[1] == 'v' && p[2] == 'e') || (p[1] == 'l' && p[2] == 'l'))) n = 3;
            if (n) { words.emplace_back(reinterpret_cast<const char*>(p), n); p += n; continue; }
        }
        uint32_t cp;
        int len = utf8_decode(p, rem, &cp);
        if (len <= 0) bad_utf8();
        if (is_digit(cp)) { words.emplace_back(reinterpret_cast<const char*>(p), static_cast<size_t>(len)); p += len; continue; }
        const bool letters = is_letter(cp);
        if (!letters && is_blank(cp)) { p += len; continue; }            // matches nothing: skip
        const unsigned char* q = p + len;
        while (q < end) {
            len = utf8_decode(q, static_cast<size_t>(end - q), &cp);
            if (len <= 0) bad_utf8();
            const bool l = is_letter(cp);
            if (letters ? !l : (l || is_digit(cp) || is_blank(cp))) break;
            q += len;
        }
        words.emplace_back(reinterpret_cast<const char*>(p), static_cast<size_t>(q - p));
        p = q;
    }
    return words;
}

// Every byte of the word becomes one printable symbol (UTF-8 encoded), tokenizer.cpp:24-53.
std::string Tokenizer::byte_symbols(const std::string& word) {
    static const struct Table {
        std::string sym[256];
        Table() { for (unsigned b = 0; b < 256; ++b) utf8_append(sym[b], byte_to_codepoint(b)); }
    } table;
    std::string out;
    out.reserve(word.size() * 2);
    for (unsigned char c : word) out += table.sym[c];
    return out;
}

// ------------------------------------------------------------------------------------------------ vocabulary
Tokenizer::Tokenizer(const std::string& bpe_file) {
    std::ifstream in(bpe_file, std::ios::binary);
    if (!in) throw TokenizerError("Tokenizer file " + bpe_file + " does not exist");       // tokenizer.cpp:230-231
    std::string line;
    unsigned next_token = 0, next_rank = 0;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        const size_t sp = line.find(' ');
        if (sp == std::string::npos) {
            ids_.emplace(line, static_cast<token_type>(next_token++));
        } else {                                               // a merge: its product is the next token, its position the rank
            ids_.emplace(line.substr(0, sp) + line.substr(sp + 1), static_cast<token_type>(next_token++));
            ranks_.emplace(line, next_rank++);
        }
    }
    finish(next_token);
}

Tokenizer::Tokenizer() {
    unsigned next_token = 0;
    for (int pass = 0; pass < 2; ++pass) {
        // gen_tokenizer_file.py:33-34: the symbols in bytes_to_unicode() order (kept bytes first, then the remapped ones), then again + "</w>"
        std::vector<unsigned> order;
        for (unsigned b = 0; b < 256; ++b) if (byte_to_codepoint(b) == b) order.push_back(b);
        for (unsigned b = 0; b < 256; ++b) if (byte_to_codepoint(b) != b) order.push_back(b);
        for (unsigned b : order) {
            std::string sym;
            utf8_append(sym, byte_to_codepoint(b));
            if (pass) sym += "</w>";
            ids_.emplace(sym, static_cast<token_type>(next_token++));
        }
    }
    finish(next_token);
}

void Tokenizer::finish(unsigned next_token) {
    if (next_token + 2 > 65536) throw TokenizerError("Tokenizer vocabulary does not fit 16-bit token ids");
    start_ = static_cast<token_type>(next_token);
    end_ = static_cast<token_type>(next_token + 1);
}

// ------------------------------------------------------------------------------------------------ encode
std::vector<Tokenizer::token_type> Tokenizer::encode(const std::string& utf8, unsigned context_len, bool* deviated) const {
    if (deviated) *deviated = false;
    std::vector<token_type> out;
    out.reserve(context_len);
    out.push_back(start_);
    const std::string clean = sanitize(utf8);
    for (const std::string& w : split_words(clean)) bpe(out, byte_symbols(w), context_len - 1, deviated);
    while (out.size() < context_len) out.push_back(end_);
    return out;
}

void Tokenizer::bpe(std::vector<token_type>& out, const std::string& symbols, unsigned max_len, bool* deviated) const {
    if (out.size() >= max_len) return;
    auto id_of = [&](const std::string& sym) {
        auto it = ids_.find(sym);
        if (it == ids_.end()) throw TokenizerError("symbol missing from the tokenizer vocabulary");
        return it->second;
    };
    // one entry per symbol character, the last one carrying the end-of-word marker
    std::vector<std::string> word;
    {
        const unsigned char* p = reinterpret_cast<const unsigned char*>(symbols.data());
        size_t left = symbols.size();
        while (left) {
            uint32_t cp;
            const int len = utf8_decode(p, left, &cp);
            if (len <= 0) bad_utf8();
            word.emplace_back(reinterpret_cast<const char*>(p), static_cast<size_t>(len));
            p += len;
            left -= static_cast<size_t>(len);
        }
    }
    if (word.empty()) return;
    word.back() += "</w>";
    std::vector<std::string> next;
    while (word.size() > 1) {
        // lowest-ranked adjacent pair (first occurrence)
        const unsigned none = ~0u;
        unsigned best = none;
        size_t at = 0;
        for (size_t i = 0; i + 1 < word.size(); ++i) {
            auto it = ranks_.find(word[i] + ' ' + word[i + 1]);
            if (it != ranks_.end() && it->second < best) { best = it->second; at = i; }
        }
        if (best == none) break;
        const std::string first = word[at], second = word[at + 1];
        // The reference's merge pass (tokenizer.cpp:339-356) holds back a symbol equal to `first` and decides on the next one.  Two quirks
        // follow and are kept, because they decide the token ids the reference produces: a held `first` followed by another `first` is
        // emitted together with it (the second one is not itself held), and a `first` held at the end of the word is dropped.
        next.clear();
        bool held = false;
        for (const std::string& w : word) {
            if (held) {
                if (w == second) next.push_back(first + second);
                else { next.push_back(first); next.push_back(w); }
                held = false;
            } else if (w == first) {
                held = true;
            } else {
                next.push_back(w);
            }
        }
        if (next == word) {
            // The pass changed nothing although the pair occurs (e.g. [a, a, b] with pair (a, b)): the reference repeats the same pass for ever.
            // Documented deviation: merge every non-overlapping occurrence left to right (the textbook BPE step) and carry on.
            if (deviated) *deviated = true;
            next.clear();
            for (size_t i = 0; i < word.size();) {
                if (i + 1 < word.size() && word[i] == first && word[i + 1] == second) { next.push_back(first + second); i += 2; }
                else next.push_back(word[i++]);
            }
        }
        word.swap(next);
    }
    for (const std::string& w : word) {
        out.push_back(id_of(w));
        if (out.size() >= max_len) return;
    }
}

}  // namespace sdod
