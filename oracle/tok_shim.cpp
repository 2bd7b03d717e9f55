// ORACLE — TEST INFRASTRUCTURE ONLY.
// extern "C" shim over the reference's own libsdod::Tokenizer (/root/reference/csrc/libsdod/src/tokenizer.{h,cpp}), compiled where the sources
// lie; no reference source is copied into this repo.  Output: oracle/_ref/libtok_ref.so (git-ignored, travels to the GPU box).
// The reference switches LC_ALL to "en_US.utf8" around every call (tokenizer.cpp:259-261); this image only has C.utf8, where that setlocale
// fails and leaves the current locale in place — so the shim selects C.utf8 first (same UTF-8 decoding and glibc character classes).
#include "tokenizer.h"

#include <clocale>
#include <cstdlib>
#include <cstring>

#define EXPORT extern "C" __attribute__((visibility("default")))

EXPORT void* tok_ref_create(const char* path) {
    try { return new libsdod::Tokenizer(path); } catch (...) { return nullptr; }
}
EXPORT void tok_ref_destroy(void* h) { delete static_cast<libsdod::Tokenizer*>(h); }
// returns the number of ids written (== context_len), or -1 when the reference throws (invalid UTF-8)
EXPORT int tok_ref_tokenize(void* h, const char* utf8, unsigned short* out, unsigned context_len) {
    std::setlocale(LC_ALL, "C.utf8");
    std::mbtowc(nullptr, nullptr, 0);      // reset mbtowc's internal state: after one invalid sequence it would reject every later prompt
    try {
        auto v = static_cast<libsdod::Tokenizer*>(h)->tokenize(std::string(utf8), context_len);
        std::memcpy(out, v.data(), v.size() * sizeof(unsigned short));
        return static_cast<int>(v.size());
    } catch (...) {
        return -1;
    }
}
