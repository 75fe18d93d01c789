// ORACLE support (test infrastructure only): C entry points over the REAL reference hashing code.
// Built only where /root/reference exists (the build container), by oracle/Makefile, from the
// reference sources where they lie (cpp/core/md5.cpp, cpp/core/sha2.cpp -- the only two reference
// translation units on this path that compile standalone) into
// oracle/_ref/libkc_ref_hash.so.  No reference source is copied into this repository.
#include "core/md5.h"
#include "core/sha2.h"

#include <cstddef>
#include <cstdint>

extern "C" {
// MD5::get(const uint8_t*, size_t, uint32_t[4])  (cpp/core/md5.h)
void kref_md5(const uint8_t* msg, size_t len, uint32_t out[4]) { MD5::get(msg, len, out); }
// SHA2::get256(const uint8_t*, size_t, uint64_t[4])  (cpp/core/sha2.h; used by Rand::init, rand.cpp:293)
void kref_sha256_u64(const uint8_t* msg, size_t len, uint64_t out[4]) { SHA2::get256(msg, len, out); }
void kref_sha256_bytes(const uint8_t* msg, size_t len, uint8_t out[32]) { SHA2::get256(msg, len, out); }
}
