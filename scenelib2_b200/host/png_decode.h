// png_decode.h — PNG -> 8-bit gray, self-contained (no zlib / libpng / OpenCV in this image).
//
// The reference reads frames and templates with cv::imread(path, 0) (framegrabber/filegrabber.cpp:106-109,
// feature.cpp:119), i.e. any image format OpenCV knows, converted to one 8-bit gray channel.  The host shim
// decodes PGM (P5 / P2) and PNG itself: RFC 1950 / 1951 inflate (stored, fixed and dynamic Huffman blocks),
// the five PNG scanline filters, colour types gray / gray+alpha / RGB / RGBA / palette at bit depths 1-16, and
// the gray conversion cv::imread(path, 0) gets from libpng (Y = (R*9797 + G*19234 + B*3737) >> 15 on 8-bit samples).
// Interlaced (Adam7) files are rejected (empty result, like a failed imread).  JPEG: jpeg_decode.h.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace sl2png {

struct BitReader {
  const uint8_t *p, *end;
  uint32_t buf = 0;
  int cnt = 0;
  bool ok = true;
  BitReader(const uint8_t *b, const uint8_t *e) : p(b), end(e) {}
  uint32_t bits(int n) {  // n <= 16, LSB first
    while (cnt < n) {
      if (p >= end) {
        ok = false;
        return 0;
      }
      buf |= (uint32_t)(*p++) << cnt;
      cnt += 8;
    }
    const uint32_t v = buf & ((1u << n) - 1u);
    buf >>= n;
    cnt -= n;
    return v;
  }
  void align() {
    buf = 0;
    cnt = 0;
  }
};

struct Huffman {  // canonical code, decoded bit by bit (frames are small; speed is irrelevant next to disk I/O)
  uint16_t count[16], symbol[288];
  bool build(const uint8_t *len, int n) {
    std::memset(count, 0, sizeof(count));
    for (int i = 0; i < n; ++i) ++count[len[i]];
    count[0] = 0;
    int left = 1;
    for (int l = 1; l < 16; ++l) {
      left = (left << 1) - count[l];
      if (left < 0) return false;
    }
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    for (int i = 0; i < n; ++i)
      if (len[i]) symbol[offs[len[i]]++] = (uint16_t)i;
    return true;
  }
  int decode(BitReader &br) const {
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; ++l) {
      code |= (int)br.bits(1);
      if (!br.ok) return -1;
      const int c = count[l];
      if (code - c < first) return symbol[index + (code - first)];
      index += c;
      first += c;
      first <<= 1;
      code <<= 1;
    }
    return -1;
  }
};

inline bool inflate(const uint8_t *src, size_t n, std::vector<uint8_t> &out) {
  if (n < 2 || (src[0] & 0x0f) != 8 || ((src[0] << 8 | src[1]) % 31) != 0 || (src[1] & 0x20)) return false;  // zlib header
  BitReader br(src + 2, src + n);
  static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  for (;;) {
    const uint32_t last = br.bits(1), type = br.bits(2);
    if (!br.ok) return false;
    if (type == 0) {  // stored
      br.align();
      if (br.end - br.p < 4) return false;
      const uint32_t len = br.p[0] | (br.p[1] << 8), nlen = br.p[2] | (br.p[3] << 8);
      br.p += 4;
      if ((len ^ nlen) != 0xffffu || (size_t)(br.end - br.p) < len) return false;
      out.insert(out.end(), br.p, br.p + len);
      br.p += len;
    } else if (type == 1 || type == 2) {
      Huffman hl, hd;
      uint8_t lens[320];
      if (type == 1) {
        for (int i = 0; i < 144; ++i) lens[i] = 8;
        for (int i = 144; i < 256; ++i) lens[i] = 9;
        for (int i = 256; i < 280; ++i) lens[i] = 7;
        for (int i = 280; i < 288; ++i) lens[i] = 8;
        hl.build(lens, 288);
        for (int i = 0; i < 30; ++i) lens[i] = 5;
        hd.build(lens, 30);
      } else {
        const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
        if (!br.ok || nlen > 286 || ndist > 30) return false;
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t cl[19] = {0};
        for (int i = 0; i < ncode; ++i) cl[order[i]] = (uint8_t)br.bits(3);
        Huffman hc;
        if (!hc.build(cl, 19)) return false;
        int idx = 0;
        while (idx < nlen + ndist) {
          const int sym = hc.decode(br);
          if (sym < 0) return false;
          if (sym < 16) {
            lens[idx++] = (uint8_t)sym;
          } else {
            int rep, val = 0;
            if (sym == 16) {
              if (idx == 0) return false;
              val = lens[idx - 1];
              rep = 3 + (int)br.bits(2);
            } else if (sym == 17) {
              rep = 3 + (int)br.bits(3);
            } else {
              rep = 11 + (int)br.bits(7);
            }
            if (idx + rep > nlen + ndist) return false;
            while (rep--) lens[idx++] = (uint8_t)val;
          }
        }
        if (!hl.build(lens, nlen) || !hd.build(lens + nlen, ndist)) return false;
      }
      for (;;) {
        const int sym = hl.decode(br);
        if (sym < 0 || !br.ok) return false;
        if (sym < 256) {
          out.push_back((uint8_t)sym);
        } else if (sym == 256) {
          break;
        } else {
          if (sym > 285) return false;
          const int len = lbase[sym - 257] + (int)br.bits(lext[sym - 257]);
          const int ds = hd.decode(br);
          if (ds < 0 || ds > 29) return false;
          const size_t dist = dbase[ds] + br.bits(dext[ds]);
          if (!br.ok || dist > out.size()) return false;
          const size_t from = out.size() - dist;
          for (int i = 0; i < len; ++i) out.push_back(out[from + i]);
        }
      }
    } else {
      return false;
    }
    if (last) return true;
  }
}

inline uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

// file bytes -> gray (row-major, width * height); false when the file is not a PNG this decoder handles
inline bool decode_gray(const uint8_t *file, size_t n, std::vector<uint8_t> &gray, int &width, int &height) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (n < 8 + 25 || std::memcmp(file, sig, 8) != 0) return false;
  size_t pos = 8;
  int depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, plte;
  bool have_ihdr = false, done = false;
  while (!done && pos + 12 <= n) {
    const uint32_t len = be32(file + pos);
    const uint8_t *tag = file + pos + 4, *data = file + pos + 8;
    if (pos + 12 + (size_t)len > n) return false;
    if (!std::memcmp(tag, "IHDR", 4)) {
      if (len != 13) return false;
      width = (int)be32(data);
      height = (int)be32(data + 4);
      depth = data[8];
      ctype = data[9];
      interlace = data[12];
      if (data[10] != 0 || data[11] != 0) return false;
      have_ihdr = true;
    } else if (!std::memcmp(tag, "PLTE", 4)) {
      plte.assign(data, data + len);
    } else if (!std::memcmp(tag, "IDAT", 4)) {
      idat.insert(idat.end(), data, data + len);
    } else if (!std::memcmp(tag, "IEND", 4)) {
      done = true;
    }
    pos += 12 + (size_t)len;
  }
  if (!have_ihdr || width <= 0 || height <= 0 || width > 16384 || height > 16384 || interlace != 0) return false;
  int channels;
  switch (ctype) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: return false;
  }
  if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4)))) return false;
  if (ctype == 3 && (depth == 16 || plte.size() < 3)) return false;
  const int bpp_bits = channels * depth;
  const size_t stride = ((size_t)width * bpp_bits + 7) / 8;
  const int bpp = bpp_bits >= 8 ? bpp_bits / 8 : 1;  // filter distance in bytes
  std::vector<uint8_t> raw;
  raw.reserve((stride + 1) * (size_t)height);
  if (!inflate(idat.data(), idat.size(), raw) || raw.size() < (stride + 1) * (size_t)height) return false;
  // unfilter in place
  std::vector<uint8_t> prev(stride, 0);
  gray.assign((size_t)width * height, 0);
  for (int y = 0; y < height; ++y) {
    uint8_t *row = raw.data() + (stride + 1) * (size_t)y;
    const int ft = row[0];
    uint8_t *cur = row + 1;
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
      int pred = 0;
      switch (ft) {
        case 0: pred = 0; break;
        case 1: pred = a; break;
        case 2: pred = b; break;
        case 3: pred = (a + b) >> 1; break;
        case 4: {
          const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
          pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
          break;
        }
        default: return false;
      }
      cur[i] = (uint8_t)(cur[i] + pred);
    }
    std::memcpy(prev.data(), cur, stride);
    uint8_t *g = gray.data() + (size_t)width * y;
    auto sample = [&](int x, int ch) -> int {  // 8-bit value of channel ch of pixel x
      if (depth == 8) return cur[(size_t)x * channels + ch];
      if (depth == 16) return cur[((size_t)x * channels + ch) * 2];  // high byte (OpenCV: >> 8 for 8-bit reads)
      const int per = 8 / depth, byte = cur[x / per], shift = 8 - depth * (x % per + 1);
      return (byte >> shift) & ((1 << depth) - 1);
    };
    // cv::imread(path, 0) lets libpng do the conversion (png_set_rgb_to_gray(png, 1, 0.299, 0.587)): 15-bit
    // coefficients 9797 / 19234 / 3737, truncated for 8-bit samples; 16-bit samples are converted at 16 bits WITH
    // rounding and then stripped to their high byte (verified against OpenCV 4.13: tests/test_host_shim.py)
    auto luma = [](int r, int gr, int b) { return (uint8_t)((r * 9797 + gr * 19234 + b * 3737) >> 15); };
    auto sample16 = [&](int x, int ch) -> long {
      const size_t o = ((size_t)x * channels + ch) * 2;
      return ((long)cur[o] << 8) | cur[o + 1];
    };
    for (int x = 0; x < width; ++x) {
      if (ctype == 0) {
        const int v = sample(x, 0);
        g[x] = (uint8_t)(depth < 8 ? v * 255 / ((1 << depth) - 1) : v);
      } else if (ctype == 4) {
        g[x] = (uint8_t)sample(x, 0);
      } else if (ctype == 3) {
        const size_t i = (size_t)sample(x, 0) * 3;
        g[x] = i + 2 < plte.size() ? luma(plte[i], plte[i + 1], plte[i + 2]) : 0;
      } else if (depth == 16) {
        g[x] = (uint8_t)(((sample16(x, 0) * 9797 + sample16(x, 1) * 19234 + sample16(x, 2) * 3737 + 16384) >> 15) >> 8);
      } else {
        g[x] = luma(sample(x, 0), sample(x, 1), sample(x, 2));
      }
    }
  }
  return true;
}

// 8-bit gray image -> PNG bytes (cv::imwrite("patch.png", patch) of MonoSLAM::SavePatch, monoslam.cpp:1569): filter 0 on
// every scanline, one zlib stream of stored deflate blocks (the patches are 11 x 11 or 15 x 15 pixels; compression is
// not the point, a file every PNG reader accepts is).
inline void encode_gray(const uint8_t *img, int w, int h, size_t stride, std::vector<uint8_t> &out) {
  auto crc32 = [](const uint8_t *p, size_t n) {
    uint32_t c = 0xffffffffu;
    for (size_t i = 0; i < n; ++i) {
      c ^= p[i];
      for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xedb88320u & (0u - (c & 1u)));
    }
    return c ^ 0xffffffffu;
  };
  auto be32 = [](std::vector<uint8_t> &v, uint32_t x) {
    for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s));
  };
  auto chunk = [&](const char *type, const std::vector<uint8_t> &body) {
    be32(out, (uint32_t)body.size());
    const size_t at = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), body.begin(), body.end());
    be32(out, crc32(out.data() + at, out.size() - at));
  };
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  out.assign(sig, sig + 8);
  std::vector<uint8_t> ihdr;
  be32(ihdr, (uint32_t)w);
  be32(ihdr, (uint32_t)h);
  const uint8_t tail[5] = {8, 0, 0, 0, 0};  // depth 8, gray, deflate, adaptive filtering, not interlaced
  ihdr.insert(ihdr.end(), tail, tail + 5);
  chunk("IHDR", ihdr);
  std::vector<uint8_t> raw;  // filter byte 0 + the scanline
  for (int y = 0; y < h; ++y) {
    raw.push_back(0);
    raw.insert(raw.end(), img + (size_t)y * stride, img + (size_t)y * stride + w);
  }
  std::vector<uint8_t> z = {0x78, 0x01};
  uint32_t a1 = 1, a2 = 0;  // adler32
  for (uint8_t b : raw) {
    a1 = (a1 + b) % 65521u;
    a2 = (a2 + a1) % 65521u;
  }
  size_t o = 0;
  do {
    const size_t nb = raw.size() - o < 65535 ? raw.size() - o : 65535;
    z.push_back(o + nb == raw.size() ? 1 : 0);  // BFINAL, BTYPE = 00 (stored)
    z.push_back((uint8_t)(nb & 255));
    z.push_back((uint8_t)(nb >> 8));
    z.push_back((uint8_t)(~nb & 255));
    z.push_back((uint8_t)((~nb >> 8) & 255));
    z.insert(z.end(), raw.begin() + o, raw.begin() + o + nb);
    o += nb;
  } while (o < raw.size());
  be32(z, (a2 << 16) | a1);
  chunk("IDAT", z);
  chunk("IEND", {});
}

}  // namespace sl2png
