// png_gray.cpp — a small PNG reader for height-field assets.
//
// MuJoCo's compiler reads `<hfield file="terrain.png">` by asking its PNG library for an 8-bit GREY image whatever the
// file's own colour type is, then flips the rows so that image row 0 is the far (+y) edge (mjCHField::LoadPNG); the
// reference loads such models through mj_loadXML (mujoco_ros/src/mujoco_env.cpp:771-911).  The vendored PNG codec of the
// reference (lodepng) serves its offscreen renderer and is out of scope; this file restates just the decode direction
// from the format specifications: PNG (ISO/IEC 15948: chunks, CRC-32, five scanline filters), zlib (RFC 1950: header,
// Adler-32) and DEFLATE (RFC 1951: stored, fixed-Huffman and dynamic-Huffman blocks).  Interlaced (Adam7) files are
// refused by name.  The grey conversion follows what that library does for this request: red channel of colour pixels,
// high byte of 16-bit samples, (value * 255) / (2^depth - 1) for shallow greys, palette lookup first.
#include "png_gray.h"

#include <cstring>

namespace b2mj {

namespace {

uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

uint32_t crc32_of(const uint8_t* p, size_t n) {
  static uint32_t table[256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = c & 1 ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    ready = true;
  }
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

// ---- DEFLATE -------------------------------------------------------------------------------------------------------
struct BitReader {
  const uint8_t* p;
  size_t n, pos = 0;
  uint32_t acc = 0;
  int have = 0;
  bool overrun = false;
  uint32_t bits(int k) {  // k <= 16, least-significant bit first
    while (have < k) {
      if (pos >= n) { overrun = true; return 0; }
      acc |= (uint32_t)p[pos++] << have;
      have += 8;
    }
    const uint32_t v = acc & ((1u << k) - 1);
    acc >>= k;
    have -= k;
    return v;
  }
  void align() { acc = 0; have = 0; }
};

// canonical Huffman code: count[len] codes of each length, symbols sorted by (length, value)
struct Huffman {
  uint16_t count[16];
  uint16_t symbol[288];
  bool build(const uint8_t* lengths, int n) {
    std::memset(count, 0, sizeof(count));
    for (int i = 0; i < n; i++) count[lengths[i]]++;
    count[0] = 0;
    int left = 1;
    for (int len = 1; len < 16; len++) {
      left = (left << 1) - count[len];
      if (left < 0) return false;  // over-subscribed
    }
    uint16_t offs[16];
    offs[1] = 0;
    for (int len = 1; len < 15; len++) offs[len + 1] = offs[len] + count[len];
    for (int i = 0; i < n; i++)
      if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
    return true;
  }
  int decode(BitReader& br) const {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len < 16; len++) {
      code |= (int)br.bits(1);
      if (br.overrun) return -1;
      const int c = count[len];
      if (code - c < first) return symbol[index + (code - first)];
      index += c;
      first = (first + c) << 1;
      code <<= 1;
    }
    return -1;
  }
};

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

bool inflate_block(BitReader& br, const Huffman& lit, const Huffman& dist, std::vector<uint8_t>& out, size_t limit) {
  for (;;) {
    const int sym = lit.decode(br);
    if (sym < 0) return false;
    if (sym < 256) {
      if (out.size() >= limit) return false;
      out.push_back((uint8_t)sym);
    } else if (sym == 256) {
      return true;
    } else {
      if (sym > 285) return false;
      const int len = kLenBase[sym - 257] + (int)br.bits(kLenExtra[sym - 257]);
      const int ds = dist.decode(br);
      if (ds < 0 || ds > 29) return false;
      const size_t d = kDistBase[ds] + br.bits(kDistExtra[ds]);
      if (br.overrun || d > out.size() || out.size() + len > limit) return false;
      const size_t from = out.size() - d;
      for (int k = 0; k < len; k++) out.push_back(out[from + k]);  // may overlap its own output: byte by byte
    }
  }
}

bool inflate(const uint8_t* p, size_t n, std::vector<uint8_t>& out, size_t limit) {
  BitReader br{p, n};
  for (;;) {
    const int last = (int)br.bits(1), type = (int)br.bits(2);
    if (br.overrun) return false;
    if (type == 0) {
      br.align();
      if (br.pos + 4 > n) return false;
      const unsigned len = p[br.pos] | p[br.pos + 1] << 8, nlen = p[br.pos + 2] | p[br.pos + 3] << 8;
      br.pos += 4;
      if ((len ^ nlen) != 0xFFFF || br.pos + len > n || out.size() + len > limit) return false;
      out.insert(out.end(), p + br.pos, p + br.pos + len);
      br.pos += len;
    } else if (type == 1) {
      uint8_t l[288], d[30];
      for (int i = 0; i < 288; i++) l[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
      std::memset(d, 5, sizeof(d));
      Huffman lit, dist;
      lit.build(l, 288);
      dist.build(d, 30);
      if (!inflate_block(br, lit, dist, out, limit)) return false;
    } else if (type == 2) {
      const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
      if (br.overrun || nlen > 286 || ndist > 30) return false;
      static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      uint8_t cl[19] = {0};
      for (int i = 0; i < ncode; i++) cl[order[i]] = (uint8_t)br.bits(3);
      Huffman clh;
      if (!clh.build(cl, 19)) return false;
      uint8_t lens[286 + 30] = {0};
      for (int i = 0; i < nlen + ndist;) {
        const int sym = clh.decode(br);
        if (sym < 0) return false;
        if (sym < 16) {
          lens[i++] = (uint8_t)sym;
        } else {
          int rep;
          uint8_t val = 0;
          if (sym == 16) {
            if (i == 0) return false;
            val = lens[i - 1];
            rep = 3 + (int)br.bits(2);
          } else if (sym == 17) {
            rep = 3 + (int)br.bits(3);
          } else {
            rep = 11 + (int)br.bits(7);
          }
          if (br.overrun || i + rep > nlen + ndist) return false;
          while (rep--) lens[i++] = val;
        }
      }
      if (lens[256] == 0) return false;  // no end-of-block code
      Huffman lit, dist;
      if (!lit.build(lens, nlen) || !dist.build(lens + nlen, ndist)) return false;
      if (!inflate_block(br, lit, dist, out, limit)) return false;
    } else {
      return false;
    }
    if (last) return !br.overrun;
  }
}

bool zlib_decompress(const std::vector<uint8_t>& z, std::vector<uint8_t>& out, size_t expect, std::string& err) {
  if (z.size() < 6 || (z[0] & 0x0F) != 8 || ((z[0] << 8 | z[1]) % 31) != 0 || (z[1] & 0x20)) {
    err = "bad zlib header";
    return false;
  }
  out.reserve(expect);
  if (!inflate(z.data() + 2, z.size() - 6, out, expect)) {
    err = "corrupt DEFLATE stream";
    return false;
  }
  uint32_t a = 1, b = 0;
  for (uint8_t v : out) {
    a = (a + v) % 65521;
    b = (b + a) % 65521;
  }
  if ((b << 16 | a) != be32(z.data() + z.size() - 4)) {
    err = "Adler-32 mismatch";
    return false;
  }
  return true;
}

int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
  return pa <= pb && pa <= pc ? a : pb <= pc ? b : c;
}

}  // namespace

bool png_decode_gray8(const uint8_t* bytes, size_t n, std::vector<uint8_t>& image, unsigned& width, unsigned& height,
                      std::string& err) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n'};
  if (n < 8 || std::memcmp(bytes, sig, 8) != 0) {
    err = "not a PNG file (bad signature)";
    return false;
  }
  size_t pos = 8;
  bool have_ihdr = false, have_end = false;
  int depth = 0, ctype = 0;
  std::vector<uint8_t> idat, palette;
  while (pos + 12 <= n && !have_end) {
    const uint32_t len = be32(bytes + pos);
    const uint8_t* tag = bytes + pos + 4;
    if ((size_t)len > n - pos - 12) {
      err = "truncated chunk";
      return false;
    }
    const uint8_t* body = tag + 4;
    if (crc32_of(tag, 4 + (size_t)len) != be32(body + len)) {
      err = "chunk CRC mismatch";
      return false;
    }
    if (!std::memcmp(tag, "IHDR", 4)) {
      if (len != 13) { err = "bad IHDR"; return false; }
      width = be32(body);
      height = be32(body + 4);
      depth = body[8];
      ctype = body[9];
      if (body[10] != 0 || body[11] != 0) { err = "unknown compression / filter method"; return false; }
      if (body[12] != 0) { err = "interlaced (Adam7) PNG files are not supported: save the image non-interlaced"; return false; }
      have_ihdr = true;
    } else if (!std::memcmp(tag, "PLTE", 4)) {
      palette.assign(body, body + len);
    } else if (!std::memcmp(tag, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + len);
    } else if (!std::memcmp(tag, "IEND", 4)) {
      have_end = true;
    } else if (!(tag[0] & 0x20)) {
      err = "unknown critical chunk";
      return false;
    }
    pos += 12 + (size_t)len;
  }
  if (!have_ihdr || !have_end || idat.empty()) {
    err = "missing IHDR / IDAT / IEND";
    return false;
  }
  const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
  const bool depth_ok = ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                        : ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8)
                                     : (depth == 8 || depth == 16);
  if (!channels || !depth_ok) { err = "invalid colour type / bit depth"; return false; }
  if (!width || !height || (uint64_t)width * height > (1u << 26)) { err = "image size out of range"; return false; }
  if (ctype == 3 && palette.size() < 3) { err = "palette image without PLTE"; return false; }

  const int bpp_bits = channels * depth;
  const size_t stride = ((size_t)width * bpp_bits + 7) / 8, bpp = (size_t)(bpp_bits + 7) / 8;
  std::vector<uint8_t> raw;
  if (!zlib_decompress(idat, raw, (stride + 1) * height, err)) return false;
  if (raw.size() != (stride + 1) * height) { err = "image data has the wrong length"; return false; }

  // undo the scanline filters in place (row r starts at r * (stride + 1) with its filter byte)
  std::vector<uint8_t> zero(stride, 0);
  for (unsigned r = 0; r < height; r++) {
    uint8_t* cur = raw.data() + (size_t)r * (stride + 1) + 1;
    const uint8_t* up = r ? cur - (stride + 1) : zero.data();
    const int f = cur[-1];
    if (f > 4) { err = "unknown scanline filter"; return false; }
    for (size_t i = 0; i < stride; i++) {
      const int a = i >= bpp ? cur[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
      const int add = f == 0 ? 0 : f == 1 ? a : f == 2 ? b : f == 3 ? (a + b) / 2 : paeth(a, b, c);
      cur[i] = (uint8_t)(cur[i] + add);
    }
  }

  image.resize((size_t)width * height);
  for (unsigned r = 0; r < height; r++) {
    const uint8_t* row = raw.data() + (size_t)r * (stride + 1) + 1;
    for (unsigned c = 0; c < width; c++) {
      unsigned v;
      if (depth == 16) {
        v = row[(size_t)c * channels * 2];  // high byte of the first (grey / red) sample
      } else if (depth == 8) {
        v = row[(size_t)c * channels];
      } else {
        const size_t bit = (size_t)c * depth;
        v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1);
        if (ctype == 0) v = v * 255 / ((1u << depth) - 1);
      }
      if (ctype == 3) {
        if ((size_t)v * 3 + 2 >= palette.size()) { err = "palette index out of range"; return false; }
        v = palette[(size_t)v * 3];
      }
      image[(size_t)r * width + c] = (uint8_t)v;
    }
  }
  return true;
}

}  // namespace b2mj
