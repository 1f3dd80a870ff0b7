// bmp_decode.h — Windows BMP -> 8-bit gray the way cv::imread(path, 0) returns it (the reference reads frames and
// templates through that call: framegrabber/filegrabber.cpp:106-109, feature.cpp:119).  Uncompressed files with a
// BITMAPCOREHEADER / BITMAPINFOHEADER / V4 / V5 header: 1, 4 and 8 bits per pixel with a palette, 24 and 32 bits per
// pixel (also as BI_BITFIELDS with the plain BGRA masks); bottom-up or top-down rows.  Colour -> gray with OpenCV's 14-bit weights
// Y = (R*4899 + G*9617 + B*1868 + 8192) >> 14 (palette entries are converted once); 32-bit pixels instead as
// (uchar)(0.299f R + 0.587f G + 0.114f B) in single precision, truncated -- checked against cv2.imread over all 2^24 colours.  RLE-compressed files and other bit-field layouts
// are rejected (empty result, like a failed imread).  tests/test_host_shim.py compares with cv2.imread(path, 0).
#pragma once
#include <cstdint>
#include <vector>

namespace sl2bmp {

inline bool decode_gray(const uint8_t *d, size_t n, std::vector<uint8_t> &gray, int &W, int &H) {
  auto u16 = [&](size_t o) { return (uint32_t)d[o] | ((uint32_t)d[o + 1] << 8); };
  auto u32 = [&](size_t o) { return u16(o) | (u16(o + 2) << 16); };
  if (n < 26 || d[0] != 'B' || d[1] != 'M') return false;
  const uint32_t off = u32(10), hs = u32(14);
  int bpp;
  long w, h;
  uint32_t comp = 0, ncol = 0;
  int entry = 4;  // bytes per palette entry
  if (hs == 12) {
    w = (int16_t)u16(18);
    h = (int16_t)u16(20);
    bpp = (int)u16(24);
    entry = 3;
  } else if (hs >= 40 && 14 + (size_t)hs <= n) {
    w = (int32_t)u32(18);
    h = (int32_t)u32(22);
    bpp = (int)u16(28);
    comp = u32(30);
    ncol = u32(46);
  } else {
    return false;
  }
  const bool topdown = h < 0;
  if (topdown) h = -h;
  if (comp == 3 && bpp == 32 && n >= 66 && u32(54) == 0x00ff0000u && u32(58) == 0x0000ff00u && u32(62) == 0x000000ffu)
    comp = 0;  // BI_BITFIELDS with the masks of plain BGRA (what cv::imwrite produces for 4 channels)
  if (w <= 0 || h <= 0 || w > 32768 || h > 32768 || comp != 0) return false;
  if (bpp != 1 && bpp != 4 && bpp != 8 && bpp != 24 && bpp != 32) return false;
  auto luma = [](int r, int g, int b) { return (uint8_t)((r * 4899 + g * 9617 + b * 1868 + 8192) >> 14); };
  uint8_t pal[256] = {0};
  if (bpp <= 8) {
    const uint32_t cap = 1u << bpp;
    if (ncol == 0 || ncol > cap) ncol = cap;
    const size_t po = 14 + (size_t)hs;
    if (po + (size_t)ncol * entry > n) return false;
    for (uint32_t i = 0; i < ncol; ++i) pal[i] = luma(d[po + i * entry + 2], d[po + i * entry + 1], d[po + i * entry]);
  }
  const size_t rowb = (((size_t)w * bpp + 31) / 32) * 4;
  if ((size_t)off + rowb * (size_t)h > n) return false;
  W = (int)w;
  H = (int)h;
  gray.assign((size_t)W * H, 0);
  for (long y = 0; y < h; ++y) {
    const uint8_t *row = d + off + rowb * (size_t)(topdown ? y : h - 1 - y);
    uint8_t *g = gray.data() + (size_t)y * W;
    for (long x = 0; x < w; ++x) {
      switch (bpp) {
        case 1: g[x] = pal[(row[x >> 3] >> (7 - (x & 7))) & 1]; break;
        case 4: g[x] = pal[(row[x >> 1] >> ((x & 1) ? 0 : 4)) & 15]; break;
        case 8: g[x] = pal[row[x]]; break;
        case 24: g[x] = luma(row[3 * x + 2], row[3 * x + 1], row[3 * x]); break;
        default: {  // 32 bits per pixel: OpenCV's reader goes through single-precision floats and truncates
          const float fr = 0.299f * (float)row[4 * x + 2], fg = 0.587f * (float)row[4 * x + 1];
          const float fb = 0.114f * (float)row[4 * x];
          const float s2 = fr + fg;
          g[x] = (uint8_t)(s2 + fb);
          break;
        }
      }
    }
  }
  return true;
}

}  // namespace sl2bmp
