// jpeg_decode.h — JPEG -> 8-bit gray, self-contained (no libjpeg / OpenCV in this image).
//
// The reference reads frames and templates with cv::imread(path, 0) (framegrabber/filegrabber.cpp:106-109,
// feature.cpp:119).  For a JPEG, OpenCV asks libjpeg for JCS_GRAYSCALE output, i.e. the luminance component alone,
// decoded with libjpeg's default "slow integer" inverse DCT: 8 column passes and 8 row passes in 13-bit fixed point with
// the published constants (FIX(0.298631336) = 2446 ... FIX(3.072711026) = 25172), rounded right shifts by 11 and 18 bits,
// + 128, clamped through a 10-bit wrap-around range table.  This file restates that algorithm so that the shim's gray
// values equal imread's bit for bit (tests/test_host_shim.py compares with OpenCV on files written by OpenCV and PIL):
// baseline / extended-sequential Huffman JPEG, 8 bits per sample, 1 component (gray) or 3 (YCbCr, any of the usual
// 4:4:4 / 4:2:2 / 4:2:0 / 4:1:1 / 4:4:0 samplings: chroma blocks are entropy-decoded and dropped), restart intervals,
// one interleaved scan.  Progressive, arithmetic-coded, 12-bit, CMYK and multi-scan files are rejected (empty result,
// like a failed imread); EXIF orientation is not applied.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace sl2jpeg {

struct Huff {
  // canonical code tables (ITU T.81 Annex C / F.2.2.3): codes of length l lie in [mincode[l], maxcode[l]]
  int mincode[17], maxcode[18], valptr[17];
  uint8_t vals[256];
  bool present = false;
};

struct Comp {
  int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0;
};

struct Bits {
  const uint8_t *p, *end;
  uint32_t buf = 0;
  int cnt = 0;
  bool ok = true;
  int marker = 0;  // a marker met inside the entropy-coded segment (RSTn / EOI): zero bits follow
  Bits(const uint8_t *b, const uint8_t *e) : p(b), end(e) {}
  void fill() {
    while (cnt <= 24) {
      int c = 0;
      if (!marker && p < end) {
        c = *p++;
        if (c == 0xFF) {
          int m = p < end ? *p : 0xD9;
          if (m == 0) {
            ++p;  // stuffed zero
          } else {
            marker = m;
            --p;  // leave the marker where it is
            c = 0;
          }
        }
      }
      buf |= (uint32_t)c << (24 - cnt);
      cnt += 8;
    }
  }
  int bit() {
    if (cnt == 0) fill();
    const int b = buf >> 31;
    buf <<= 1;
    --cnt;
    return b;
  }
  int bits(int n) {  // n <= 16, MSB first
    if (n == 0) return 0;
    if (cnt < n) fill();
    const int v = (int)(buf >> (32 - n));
    buf <<= n;
    cnt -= n;
    return v;
  }
  void reset() {
    buf = 0;
    cnt = 0;
  }
};

inline bool build_huff(Huff &h, const uint8_t *counts /*16*/, const uint8_t *vals, int nvals) {
  int code = 0, k = 0;
  for (int l = 1; l <= 16; ++l) {
    h.valptr[l] = k;
    h.mincode[l] = code;
    k += counts[l - 1];
    code += counts[l - 1];
    h.maxcode[l] = counts[l - 1] ? code - 1 : -1;
    code <<= 1;
  }
  h.maxcode[17] = 0x7fffffff;
  if (k != nvals || k > 256) return false;
  std::memcpy(h.vals, vals, (size_t)nvals);
  h.present = true;
  return true;
}

inline int decode_sym(Bits &br, const Huff &h) {
  int code = 0;
  for (int l = 1; l <= 16; ++l) {
    code = (code << 1) | br.bit();
    if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
  }
  br.ok = false;
  return 0;
}

inline int extend(int v, int n) { return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v; }  // T.81 F.2.2.1

// libjpeg's jpeg_idct_islow on one dequantised block (natural order), output rows into out[8][stride]
inline void idct_islow(const int *blk, uint8_t *out, int stride) {
  constexpr int CB = 13, P1 = 2;
  constexpr long F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299,
                 F1847 = 15137, F1961 = 16069, F2053 = 16819, F2562 = 20995, F3072 = 25172;
  long ws[64];
  auto descale = [](long x, int n) { return (x + (1L << (n - 1))) >> n; };
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < 8; ++i) {
      long in[8];
      for (int k = 0; k < 8; ++k) in[k] = pass == 0 ? (long)blk[8 * k + i] : ws[8 * i + k];
      // even part
      long z2 = in[2], z3 = in[6];
      long z1 = (z2 + z3) * F0541;
      const long t2e = z1 + z3 * (-F1847), t3e = z1 + z2 * F0765;
      const long t0e = (in[0] + in[4]) * (1L << CB), t1e = (in[0] - in[4]) * (1L << CB);
      const long t10 = t0e + t3e, t13 = t0e - t3e, t11 = t1e + t2e, t12 = t1e - t2e;
      // odd part
      long t0 = in[7], t1 = in[5], t2 = in[3], t3 = in[1];
      z1 = t0 + t3;
      z2 = t1 + t2;
      z3 = t0 + t2;
      long z4 = t1 + t3;
      const long z5 = (z3 + z4) * F1175;
      t0 *= F0298;
      t1 *= F2053;
      t2 *= F3072;
      t3 *= F1501;
      z1 *= -F0899;
      z2 *= -F2562;
      z3 *= -F1961;
      z4 *= -F0390;
      z3 += z5;
      z4 += z5;
      t0 += z1 + z3;
      t1 += z2 + z4;
      t2 += z2 + z3;
      t3 += z1 + z4;
      const long o[8] = {t10 + t3, t11 + t2, t12 + t1, t13 + t0, t13 - t0, t12 - t1, t11 - t2, t10 - t3};
      if (pass == 0) {
        for (int k = 0; k < 8; ++k) ws[8 * k + i] = descale(o[k], CB - P1);
      } else {
        for (int k = 0; k < 8; ++k) {
          const int idx = (int)(descale(o[k], CB + P1 + 3) & 1023);  // libjpeg's range_limit table, RANGE_MASK
          out[i * stride + k] = (uint8_t)(idx < 128 ? idx + 128 : (idx < 512 ? 255 : (idx < 896 ? 0 : idx - 896)));
        }
      }
    }
  }
}

// JPEG bytes -> gray (h x w, row-major).  Returns false for anything this decoder does not cover.
inline bool decode_gray(const uint8_t *d, size_t n, std::vector<uint8_t> &gray, int &W, int &H) {
  static const uint8_t ZZ[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return false;
  uint16_t qt[4][64];
  bool qt_ok[4] = {false, false, false, false};
  Huff hdc[4], hac[4];
  Comp comp[3];
  int ncomp = 0, restart = 0;
  bool have_sof = false;
  size_t pos = 2;
  while (pos + 4 <= n) {
    if (d[pos] != 0xFF) return false;
    while (pos < n && d[pos] == 0xFF) ++pos;  // fill bytes
    if (pos >= n) return false;
    const int m = d[pos++];
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;  // stand-alone markers
    if (m == 0xD9) return false;                                        // EOI before a scan
    if (pos + 2 > n) return false;
    const size_t len = ((size_t)d[pos] << 8) | d[pos + 1];
    if (len < 2 || pos + len > n) return false;
    const uint8_t *s = d + pos + 2;
    const size_t sl = len - 2;
    if (m == 0xDB) {  // DQT
      size_t o = 0;
      while (o < sl) {
        const int pq = s[o] >> 4, tq = s[o] & 15;
        ++o;
        if (tq > 3 || pq > 1 || o + (pq ? 128 : 64) > sl) return false;
        for (int k = 0; k < 64; ++k) {
          qt[tq][ZZ[k]] = pq ? (uint16_t)((s[o] << 8) | s[o + 1]) : s[o];
          o += pq ? 2 : 1;
        }
        qt_ok[tq] = true;
      }
    } else if (m == 0xC4) {  // DHT
      size_t o = 0;
      while (o + 17 <= sl) {
        const int tc = s[o] >> 4, th = s[o] & 15;
        if (tc > 1 || th > 3) return false;
        int nv = 0;
        for (int k = 0; k < 16; ++k) nv += s[o + 1 + k];
        if (o + 17 + nv > sl) return false;
        if (!build_huff(tc ? hac[th] : hdc[th], s + o + 1, s + o + 17, nv)) return false;
        o += 17 + (size_t)nv;
      }
    } else if (m == 0xC0 || m == 0xC1) {  // SOF0 / SOF1: sequential Huffman
      if (sl < 6 || s[0] != 8) return false;
      H = (s[1] << 8) | s[2];
      W = (s[3] << 8) | s[4];
      ncomp = s[5];
      if ((ncomp != 1 && ncomp != 3) || W <= 0 || H <= 0 || sl < (size_t)(6 + 3 * ncomp)) return false;
      for (int c = 0; c < ncomp; ++c) {
        comp[c].id = s[6 + 3 * c];
        comp[c].h = s[7 + 3 * c] >> 4;
        comp[c].v = s[7 + 3 * c] & 15;
        comp[c].tq = s[8 + 3 * c];
        if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4 || comp[c].tq > 3) return false;
      }
      have_sof = true;
    } else if (m == 0xC2 || (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)) {
      return false;  // progressive, lossless, arithmetic, hierarchical
    } else if (m == 0xDD) {  // DRI
      if (sl < 2) return false;
      restart = (s[0] << 8) | s[1];
    } else if (m == 0xDA) {  // SOS: one interleaved scan with every component of the frame
      if (!have_sof || sl < (size_t)(1 + 2 * ncomp + 3) || s[0] != ncomp) return false;
      for (int c = 0; c < ncomp; ++c) {
        if (s[1 + 2 * c] != comp[c].id) return false;
        comp[c].td = s[2 + 2 * c] >> 4;
        comp[c].ta = s[2 + 2 * c] & 15;
        if (comp[c].td > 3 || comp[c].ta > 3 || !hdc[comp[c].td].present || !hac[comp[c].ta].present ||
            !qt_ok[comp[c].tq])
          return false;
      }
      pos += len;
      break;
    }
    pos += len;
  }
  if (!have_sof || pos >= n) return false;
  int hmax = 1, vmax = 1;
  for (int c = 0; c < ncomp; ++c) {
    hmax = comp[c].h > hmax ? comp[c].h : hmax;
    vmax = comp[c].v > vmax ? comp[c].v : vmax;
  }
  if (ncomp == 1) comp[0].h = comp[0].v = hmax = vmax = 1;  // a single-component scan is never interleaved
  if (comp[0].h != hmax || comp[0].v != vmax) return false;  // luminance below full resolution: not covered
  const int mcuw = 8 * hmax, mcuh = 8 * vmax;
  const int mx = (W + mcuw - 1) / mcuw, my = (H + mcuh - 1) / mcuh;
  const int PW = mx * mcuw, PH = my * mcuh;
  if ((size_t)PW * (size_t)PH > ((size_t)1 << 27)) return false;  // a header that claims > 128 Mpixel: not a frame
  std::vector<uint8_t> plane((size_t)PW * PH);
  Bits br(d + pos, d + n);
  int blk[64], todo = restart, next_rst = 0;
  for (int c = 0; c < ncomp; ++c) comp[c].pred = 0;
  for (int y = 0; y < my; ++y) {
    for (int x = 0; x < mx; ++x) {
      if (restart && todo == 0) {  // RSTn between MCUs: byte-align, reset the predictors
        br.reset();            // what is left in the buffer is the padding of the interval's last byte
        if (!br.marker) {      // the refill has not run into the marker yet: it must be next
          while (br.p + 1 < br.end && br.p[0] == 0xFF && br.p[1] == 0xFF) ++br.p;
          if (br.p + 1 >= br.end || br.p[0] != 0xFF) return false;
          br.marker = br.p[1];
        }
        if (br.marker != 0xD0 + next_rst) return false;
        br.p += 2;  // the marker itself (its 0xFF was left in place)
        br.marker = 0;
        next_rst = (next_rst + 1) & 7;
        todo = restart;
        for (int c = 0; c < ncomp; ++c) comp[c].pred = 0;
      }
      for (int c = 0; c < ncomp; ++c) {
        for (int by = 0; by < comp[c].v; ++by)
          for (int bx = 0; bx < comp[c].h; ++bx) {
            std::memset(blk, 0, sizeof blk);
            const int t = decode_sym(br, hdc[comp[c].td]);
            if (t > 11) return false;
            const int diff = t ? extend(br.bits(t), t) : 0;
            comp[c].pred += diff;
            const uint16_t *q = qt[comp[c].tq];
            blk[0] = comp[c].pred * q[0];
            for (int k = 1; k < 64;) {
              const int rs = decode_sym(br, hac[comp[c].ta]);
              const int r = rs >> 4, sz = rs & 15;
              if (sz == 0) {
                if (r != 15) break;  // EOB
                k += 16;
                continue;
              }
              k += r;
              if (k > 63) return false;
              blk[ZZ[k]] = extend(br.bits(sz), sz) * q[ZZ[k]];
              ++k;
            }
            if (!br.ok) return false;
            if (c == 0) idct_islow(blk, plane.data() + (size_t)(y * mcuh + 8 * by) * PW + x * mcuw + 8 * bx, PW);
          }
      }
      if (restart) --todo;
    }
  }
  gray.resize((size_t)W * H);
  for (int r = 0; r < H; ++r) std::memcpy(gray.data() + (size_t)r * W, plane.data() + (size_t)r * PW, (size_t)W);
  return true;
}

}  // namespace sl2jpeg
