// npzwrite.cpp -- the training-data file of the reference, written from C: TrainingWriteBuffers::writeToZipFile
// (cpp/dataio/trainingwrite.cpp:566-587) = one zip archive with five deflated members, each a numpy array with the 256-byte
// version-1.0 header NumpyBuffer prepares (cpp/dataio/numpywrite.cpp:110-222: "{'descr':'<f4','fortran_order':False,'shape':(N,C)}",
// space-padded, newline-terminated), member names without an extension (binaryInputNCHWPacked, globalInputNC,
// policyTargetsNCMove, globalTargetsNC, valueTargetsNCHW), which numpy.load and python/shuffle.py read as an .npz.
// The reference goes through libzip; this is a direct zip writer over zlib's raw deflate (already a dependency of the model-file
// reader).  The file is written under a temporary name and renamed, like TrainingDataWriter does (trainingwrite.cpp:760-770).
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "kc_internal.h"

namespace {

constexpr int TOTAL_HEADER_BYTES = 256;   // numpywrite.h:24

std::string npyHeader(const char* descr, const std::vector<int64_t>& shape) {
  std::string s(TOTAL_HEADER_BYTES, ' ');
  s[0] = (char)0x93; s[1] = 'N'; s[2] = 'U'; s[3] = 'M'; s[4] = 'P'; s[5] = 'Y'; s[6] = 1; s[7] = 0;
  s[8] = (char)((TOTAL_HEADER_BYTES - 10) & 0xff); s[9] = (char)((TOTAL_HEADER_BYTES - 10) >> 8);
  std::string dict = std::string("{'descr':'") + descr + "','fortran_order':False,'shape':(";
  for(size_t i = 0; i < shape.size(); i++) {
    dict += std::to_string(shape[i]);
    if(i + 1 < shape.size() || shape.size() == 1) dict += ",";
  }
  dict += ")}";
  memcpy(&s[10], dict.data(), dict.size());
  s[TOTAL_HEADER_BYTES - 1] = '\n';
  return s;
}

struct Member { std::string name; uint32_t crc, compSize, rawSize, offset; };

void put16(std::string& o, uint32_t v) { o.push_back((char)(v & 0xff)); o.push_back((char)((v >> 8) & 0xff)); }
void put32(std::string& o, uint32_t v) { put16(o, v & 0xffff); put16(o, v >> 16); }

// raw deflate of header + payload
int deflateMember(const std::string& header, const void* data, size_t bytes, std::string& out, uint32_t& crc) {
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if(deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return kc::fail("kc_training_write_npz: deflateInit2 failed");
  out.resize(deflateBound(&zs, (uLong)(header.size() + bytes)) + 64);
  zs.next_out = (Bytef*)&out[0]; zs.avail_out = (uInt)out.size();
  zs.next_in = (Bytef*)header.data(); zs.avail_in = (uInt)header.size();
  int rc = deflate(&zs, bytes ? Z_NO_FLUSH : Z_FINISH);
  if(bytes && rc == Z_OK) {
    zs.next_in = (Bytef*)data; zs.avail_in = (uInt)bytes;
    rc = deflate(&zs, Z_FINISH);
  }
  const bool ok = rc == Z_STREAM_END;
  out.resize(zs.total_out);
  deflateEnd(&zs);
  if(!ok) return kc::fail("kc_training_write_npz: deflate failed");
  crc = (uint32_t)crc32(crc32(0L, (const Bytef*)header.data(), (uInt)header.size()), (const Bytef*)data, (uInt)bytes);
  return 0;
}

}  // namespace

extern "C" int kc_training_write_npz(const char* path, int numRows, int xSize, int ySize, const uint8_t* binaryInputNCHWPacked, const float* globalInputNC,
                                     const int16_t* policyTargetsNCMove, const float* globalTargetsNC, const int8_t* valueTargetsNCHW) {
  KC_CHECK(path && numRows >= 0 && xSize >= 2 && ySize >= 2 && xSize <= KC_MAX_LEN && ySize <= KC_MAX_LEN, "kc_training_write_npz: bad argument");
  KC_CHECK(numRows == 0 || (binaryInputNCHWPacked && globalInputNC && policyTargetsNCMove && globalTargetsNC && valueTargetsNCHW),
           "kc_training_write_npz: null array");
  const int64_t N = numRows, HW = (int64_t)xSize * ySize, packed = (HW + 7) / 8;
  struct Src { const char* name; const char* descr; std::vector<int64_t> shape; const void* data; size_t elt; };
  const Src srcs[5] = {
    {"binaryInputNCHWPacked", "|u1", {N, KC_NUM_SPATIAL_V1, packed}, binaryInputNCHWPacked, 1},   // trainingwrite.h:127-131
    {"globalInputNC", "<f4", {N, KC_NUM_GLOBAL_V1}, globalInputNC, 4},
    {"policyTargetsNCMove", "<i2", {N, 2, 4 * HW}, policyTargetsNCMove, 2},
    {"globalTargetsNC", "<f4", {N, 64}, globalTargetsNC, 4},
    {"valueTargetsNCHW", "|i1", {N, 5, ySize, xSize}, valueTargetsNCHW, 1},
  };
  const std::string tmp = std::string(path) + ".tmp";
  FILE* f = fopen(tmp.c_str(), "wb");
  KC_CHECK(f, std::string("kc_training_write_npz: cannot open ") + tmp);
  std::vector<Member> members;
  uint64_t offset = 0;
  bool ioOk = true;
  for(const Src& s : srcs) {
    size_t bytes = s.elt;
    for(int64_t d : s.shape) bytes *= (size_t)d;
    if(bytes + TOTAL_HEADER_BYTES >= 0xffffffffULL || offset >= 0xffffffffULL) { fclose(f); remove(tmp.c_str()); return kc::fail("kc_training_write_npz: member of 4 GiB or more (zip64 is not written): write fewer rows per file"); }
    const std::string header = npyHeader(s.descr, s.shape);
    std::string comp; uint32_t crc = 0;
    if(deflateMember(header, s.data, bytes, comp, crc)) { fclose(f); remove(tmp.c_str()); return 1; }
    Member m{s.name, crc, (uint32_t)comp.size(), (uint32_t)(bytes + TOTAL_HEADER_BYTES), (uint32_t)offset};
    std::string local;
    put32(local, 0x04034b50); put16(local, 20); put16(local, 0); put16(local, 8); put16(local, 0); put16(local, 0x21);   // 1980-01-01
    put32(local, m.crc); put32(local, m.compSize); put32(local, m.rawSize); put16(local, (uint32_t)m.name.size()); put16(local, 0);
    local += m.name;
    ioOk = ioOk && fwrite(local.data(), 1, local.size(), f) == local.size() && fwrite(comp.data(), 1, comp.size(), f) == comp.size();
    offset += local.size() + comp.size();
    members.push_back(m);
  }
  std::string central;
  for(const Member& m : members) {
    put32(central, 0x02014b50); put16(central, 20); put16(central, 20); put16(central, 0); put16(central, 8); put16(central, 0); put16(central, 0x21);
    put32(central, m.crc); put32(central, m.compSize); put32(central, m.rawSize); put16(central, (uint32_t)m.name.size());
    put16(central, 0); put16(central, 0); put16(central, 0); put16(central, 0); put32(central, 0); put32(central, m.offset);
    central += m.name;
  }
  std::string eocd;
  put32(eocd, 0x06054b50); put16(eocd, 0); put16(eocd, 0); put16(eocd, (uint32_t)members.size()); put16(eocd, (uint32_t)members.size());
  put32(eocd, (uint32_t)central.size()); put32(eocd, (uint32_t)offset); put16(eocd, 0);
  ioOk = ioOk && fwrite(central.data(), 1, central.size(), f) == central.size() && fwrite(eocd.data(), 1, eocd.size(), f) == eocd.size();
  ioOk = (fclose(f) == 0) && ioOk;
  if(!ioOk) { remove(tmp.c_str()); return kc::fail(std::string("kc_training_write_npz: write error on ") + tmp); }
  if(rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return kc::fail(std::string("kc_training_write_npz: cannot rename to ") + path); }
  return 0;
}
