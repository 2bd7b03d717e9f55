// CLIP byte-pair tokenizer of the prompt path.
//
// Replaces libsdod::Tokenizer (reference csrc/libsdod/src/tokenizer.{h,cpp}): same vocabulary file (`ctokenizer.txt`: one symbol per line,
// then one "first second" merge per line; token id = line index, start / end tokens appended — tokenizer.cpp:228-255, written by
// gen_tokenizer_file.py:27-42), same 77-slot output (start token, BPE ids, padded with the end token — tokenizer.cpp:258-276), and the same
// token ids for every prompt on which the reference terminates.  Unlike the reference it does not touch the process locale
// (tokenizer.cpp:259-261 switches LC_ALL to en_US.utf8 around every call and fails on hosts without that locale): UTF-8 is decoded here
// and the character classes come from tables frozen from glibc (unicode_tables.inc, tools/gen_unicode_tables.py).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace sdod {

struct TokenizerError : std::runtime_error {      // maps to LIBSDOD_INVALID_ARGUMENT at the C API (reference: ErrorCode::INVALID_ARGUMENT)
    using std::runtime_error::runtime_error;
};

class Tokenizer {
public:
    using token_type = uint16_t;                  // tokenizer.h:20

    // vocabulary + merges from a ctokenizer.txt
    explicit Tokenizer(const std::string& bpe_file);
    // byte-level vocabulary without merges (the 256 byte symbols and their "</w>" forms, ids as gen_tokenizer_file.py:33-34 lays them out):
    // what "random-init" contexts use, where no vocabulary file exists
    Tokenizer();

    // [start, ids..., end, end, ...] of exactly context_len entries
    // *deviated (optional) is set when a merge pass of the reference would not terminate on this prompt (see bpe()) and the textbook merge was
    // used instead; for every other prompt the ids equal the reference's.
    std::vector<token_type> encode(const std::string& utf8, unsigned context_len = 77, bool* deviated = nullptr) const;

    token_type start_token() const { return start_; }
    token_type end_token() const { return end_; }
    size_t vocab_size() const { return static_cast<size_t>(end_) + 1; }
    size_t merges() const { return ranks_.size(); }

    // the three stages, exposed for the tests
    static std::string sanitize(const std::string& utf8);                       // tokenizer.cpp:55-109
    static std::vector<std::string> split_words(const std::string& clean);      // tokenizer.cpp:114-222
    static std::string byte_symbols(const std::string& word);                   // tokenizer.cpp:24-53

private:
    std::unordered_map<std::string, token_type> ids_;
    std::unordered_map<std::string, unsigned> ranks_;      // key: first + ' ' + second
    token_type start_ = 0, end_ = 0;

    void finish(unsigned next_token);
    void bpe(std::vector<token_type>& out, const std::string& symbols, unsigned max_len, bool* deviated) const;    // tokenizer.cpp:279-369
};

}  // namespace sdod
