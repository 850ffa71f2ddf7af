// png_gray.h — PNG file -> 8-bit grey image, for height-field assets (<hfield file="*.png">).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace b2mj {

// Decodes a PNG held in memory into width * height grey bytes, row-major from the top row.  Returns false and sets err on
// a malformed / truncated / interlaced stream.  Colour images contribute their red channel, 16-bit samples their high
// byte, samples below 8 bits are scaled to 0..255, palette entries are looked up first.
bool png_decode_gray8(const uint8_t* bytes, size_t n, std::vector<uint8_t>& image, unsigned& width, unsigned& height,
                      std::string& err);

}  // namespace b2mj
