// Shared by the host-side file readers (fast5.cu, vbz.cu): one error type that never crosses the C ABI, and the
// decompressors the HDF5 filter pipeline calls.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace ndhost {

struct Error {
  std::string msg;
};

[[noreturn]] inline void fail(const std::string& m) { throw Error{m}; }

// One or more concatenated Zstandard frames (RFC 8878) -> out (replaced).  More than `limit` bytes of output is an error.
void zstd_decompress(const uint8_t* src, size_t n, std::vector<uint8_t>& out, size_t limit);

// One chunk written by ONT's VBZ HDF5 filter (id 32020), client values cd[0..ncd) = {vbz version, integer size, zig-zag
// delta, zstd level} -> the chunk's little-endian integers.
void vbz_decompress(const uint8_t* src, size_t n, const uint32_t* cd, int ncd, std::vector<uint8_t>& out, size_t limit);

}  // namespace ndhost
