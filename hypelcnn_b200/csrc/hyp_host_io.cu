// Host-only helpers behind the C ABI: the CRC-32C of the TFRecord framing and the LZW decoder of compressed TIFF strips.
// No device code; kept out of the engine's translation unit.
#include "hyp_common.cuh"

using namespace hyp;

extern "C" {

// host-only: TIFF LZW (MSB-first codes of 9..12 bits, ClearCode 256, EndOfInformation 257, the code width grows one
// code early) — the strips of compressed scene files (hypelcnn_b200/utilities/tiff_io.py; the reference reads them
// through tifffile).  Table entries are (prefix code, last byte, length); a string is written back to front.
int hyp_tiff_lzw_decode(const void* data, uint64_t len, void* out, uint64_t out_capacity, uint64_t* out_len) {
  HYP_CHECK_ARG(out_len && (data || len == 0) && (out || out_capacity == 0), "null argument");
  const uint8_t* src = static_cast<const uint8_t*>(data);
  uint8_t* dst = static_cast<uint8_t*>(out);
  static thread_local uint16_t prefix[4096];
  static thread_local uint8_t last[4096], first[4096];
  static thread_local uint32_t length[4096];
  for (int i = 0; i < 256; i++) {
    prefix[i] = 0xffff;
    last[i] = first[i] = (uint8_t)i;
    length[i] = 1;
  }
  uint64_t pos = 0, written = 0;
  uint32_t buffer = 0;
  int buffered = 0, width = 9, next = 258, previous = -1;
  bool corrupt = false;
  while (written < out_capacity) {
    while (buffered < width && pos < len) {
      buffer = (buffer << 8) | src[pos++];
      buffered += 8;
    }
    if (buffered < width) break;
    const int code = (int)((buffer >> (buffered - width)) & ((1u << width) - 1u));
    buffered -= width;
    if (code == 256) {  // ClearCode
      width = 9;
      next = 258;
      previous = -1;
      continue;
    }
    if (code == 257) break;  // EndOfInformation
    if (previous < 0) {
      if (code >= 256) { corrupt = true; break; }
    } else {
      if (code > next || next >= 4096) { corrupt = true; break; }
      // new string = previous string + first byte of this code's string (of the previous one when the code is new)
      prefix[next] = (uint16_t)previous;
      first[next] = first[previous];
      last[next] = (code < next) ? first[code] : first[previous];
      length[next] = length[previous] + 1;
      next++;
    }
    const uint32_t n = length[code];
    const uint64_t room = out_capacity - written;
    const uint32_t keep = n <= room ? n : (uint32_t)room;  // a string may run past the strip's last byte
    int walk = code;
    for (uint32_t k = n; k-- > 0;) {
      if (k < keep) dst[written + k] = last[walk];
      if (prefix[walk] != 0xffff) walk = prefix[walk];
    }
    written += keep;
    previous = code;
    if (next + 1 >= (1 << width) && width < 12) width++;
  }
  *out_len = written;
  HYP_CHECK_ARG(!corrupt, "corrupt LZW stream");
  return HYP_OK;
}

// host-only: CRC-32C (Castagnoli, reflected 0x82F63B78), slicing-by-8 — the checksum of the TFRecord framing
// (importer/TFRecordImporter.py, utilities/tfrecord_writer.py read / write TFRecord files through TensorFlow)
int hyp_crc32c(const void* data, uint64_t len, uint32_t* crc_inout) {
  HYP_CHECK_ARG(crc_inout && (data || len == 0), "null argument");
  static uint32_t table[8][256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t n = 0; n < 256; n++) {
      uint32_t c = n;
      for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      table[0][n] = c;
    }
    for (uint32_t n = 0; n < 256; n++)
      for (int t = 1; t < 8; t++) table[t][n] = (table[t - 1][n] >> 8) ^ table[0][table[t - 1][n] & 0xffu];
    ready = true;
  }
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = ~*crc_inout;
  while (len >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;  // little-endian hosts (x86-64, aarch64)
    c = table[7][w & 0xff] ^ table[6][(w >> 8) & 0xff] ^ table[5][(w >> 16) & 0xff] ^ table[4][(w >> 24) & 0xff] ^
        table[3][(w >> 32) & 0xff] ^ table[2][(w >> 40) & 0xff] ^ table[1][(w >> 48) & 0xff] ^ table[0][w >> 56];
    p += 8;
    len -= 8;
  }
  while (len--) c = (c >> 8) ^ table[0][(c ^ *p++) & 0xffu];
  *crc_inout = ~c;
  return HYP_OK;
}


}  // extern "C"
