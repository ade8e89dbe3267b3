// Hybrid JPEG decode for the real-data ingest (reference dali_dataloader.py:65-72, 140-145:
// fn.decoders.image(_random_crop)(device="mixed") -- nvJPEG's hybrid back end: entropy decoding on the
// host, everything after it on the GPU).  Same split here:
//   host   sib_jpeg_parse                 marker parser: frame / scan geometry, quantisation tables
//          sib_jpeg_decode_coefficients   baseline / extended-sequential Huffman decoding into int16
//                                         DCT coefficients (natural order, NOT dequantised); called
//                                         from a thread pool, one image per call (no global state)
//   device sib_jpeg_idct_rgb              dequantisation + 8x8 inverse DCT -> component planes, then
//                                         chroma upsampling + YCbCr -> RGB straight into the packed
//                                         uint8 [H_i][W_i][3] buffer the ragged augmentation kernels read
// Arithmetic = the published integer algorithms of the IJG / libjpeg-turbo decoder with its default
// settings (what PIL decodes with, and what records.decode_image used on the host until now):
// jidctint.c "islow" IDCT (13-bit constants, 2 extra bits after the column pass), jdsample.c "fancy"
// triangle upsampling for 2x1 / 2x2 chroma, jdcolor.c 16-bit fixed-point YCbCr -> RGB.  All integer:
// the device output is bit-identical to PIL on the same stream (tests/test_jpeg.py, test_gpu_jpeg.py).
// Streams outside this subset (progressive, arithmetic coding, 12-bit, CMYK / Adobe-RGB, multi-scan,
// unusual sampling factors) are reported by sib_jpeg_parse and decoded by the caller's host decoder.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host.h"
#include "../../include/sib200.h"

namespace sib {

// zigzag position -> natural (row-major) position inside an 8x8 block
static const unsigned char kNatural[64 + 16] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
    63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};   // (guard for corrupt runs)

static inline int be16(const unsigned char* p) { return (p[0] << 8) | p[1]; }

// ---------------------------------------------------------------------------------------------
// marker parser
// ---------------------------------------------------------------------------------------------
struct HuffSpec {
  unsigned char bits[17];
  unsigned char vals[256];
  bool present;
};

struct ParsedJpeg {
  sib_jpeg_info info;
  HuffSpec dc[4], ac[4];
  int dc_sel[3], ac_sel[3];
  long scan_begin;       // first byte of entropy-coded data
};

static int parse_jpeg(const unsigned char* d, long n, ParsedJpeg* pj) {
  sib_jpeg_info& o = pj->info;
  memset(&o, 0, sizeof(o));
  for (int i = 0; i < 4; ++i) pj->dc[i].present = pj->ac[i].present = false;
  unsigned short qt[4][64];
  bool have_q[4] = {false, false, false, false};
  int comp_id[3] = {0, 0, 0}, comp_q[3] = {0, 0, 0};
  bool have_sof = false, jfif = false, adobe = false;
  int adobe_transform = -1;
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) { o.status = SIB_JPEG_NOT_JPEG; return 0; }
  long p = 2;
  while (p + 4 <= n) {
    if (d[p] != 0xFF) { o.status = SIB_JPEG_CORRUPT; return 0; }
    int m = d[p + 1];
    if (m == 0xFF) { ++p; continue; }             // fill bytes
    p += 2;
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
    if (m == 0xD9) break;
    if (p + 2 > n) break;
    const int len = be16(d + p);
    if (len < 2 || p + len > n) { o.status = SIB_JPEG_CORRUPT; return 0; }
    const unsigned char* s = d + p + 2;
    const int sl = len - 2;
    if (m == 0xC0 || m == 0xC1) {                 // baseline / extended sequential, Huffman
      if (have_sof || sl < 6) { o.status = SIB_JPEG_CORRUPT; return 0; }
      if (s[0] != 8) { o.status = SIB_JPEG_UNSUPPORTED_PRECISION; return 0; }
      o.height = be16(s + 1);
      o.width = be16(s + 3);
      o.ncomp = s[5];
      if (o.height == 0 || o.width == 0) { o.status = SIB_JPEG_CORRUPT; return 0; }
      if (o.ncomp != 1 && o.ncomp != 3) { o.status = SIB_JPEG_UNSUPPORTED_COLORSPACE; return 0; }
      if (sl < 6 + 3 * o.ncomp) { o.status = SIB_JPEG_CORRUPT; return 0; }
      for (int c = 0; c < o.ncomp; ++c) {
        comp_id[c] = s[6 + 3 * c];
        o.hs[c] = s[7 + 3 * c] >> 4;
        o.vs[c] = s[7 + 3 * c] & 15;
        comp_q[c] = s[8 + 3 * c] & 3;
      }
      have_sof = true;
    } else if ((m >= 0xC2 && m <= 0xCF) && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      o.status = SIB_JPEG_UNSUPPORTED_PROCESS;     // progressive, lossless, arithmetic, hierarchical
      return 0;
    } else if (m == 0xC4) {                        // DHT
      int q = 0;
      while (q + 17 <= sl) {
        const int tc = s[q] >> 4, th = s[q] & 15;
        if (tc > 1 || th > 3) { o.status = SIB_JPEG_CORRUPT; return 0; }
        HuffSpec& h = tc ? pj->ac[th] : pj->dc[th];
        h.bits[0] = 0;
        int cnt = 0;
        for (int i = 1; i <= 16; ++i) { h.bits[i] = s[q + i]; cnt += h.bits[i]; }
        if (cnt > 256 || q + 17 + cnt > sl) { o.status = SIB_JPEG_CORRUPT; return 0; }
        memcpy(h.vals, s + q + 17, cnt);
        h.present = true;
        q += 17 + cnt;
      }
    } else if (m == 0xDB) {                        // DQT (stored in zigzag order)
      int q = 0;
      while (q < sl) {
        const int pq = s[q] >> 4, tq = s[q] & 15;
        if (tq > 3 || pq > 1 || q + 1 + 64 * (pq + 1) > sl) { o.status = SIB_JPEG_CORRUPT; return 0; }
        for (int i = 0; i < 64; ++i)
          qt[tq][kNatural[i]] = pq ? (unsigned short)be16(s + q + 1 + 2 * i) : s[q + 1 + i];
        have_q[tq] = true;
        q += 1 + 64 * (pq + 1);
      }
    } else if (m == 0xDD) {                        // DRI
      if (sl < 2) { o.status = SIB_JPEG_CORRUPT; return 0; }
      o.restart_interval = be16(s);
    } else if (m == 0xE0) {
      if (sl >= 5 && memcmp(s, "JFIF", 5) == 0) jfif = true;
    } else if (m == 0xEE) {
      if (sl >= 12 && memcmp(s, "Adobe", 5) == 0) { adobe = true; adobe_transform = s[11]; }
    } else if (m == 0xDA) {                        // SOS
      if (!have_sof) { o.status = SIB_JPEG_CORRUPT; return 0; }
      const int ns = s[0];
      if (ns != o.ncomp || sl < 1 + 2 * ns + 3) { o.status = SIB_JPEG_UNSUPPORTED_SCANS; return 0; }
      for (int c = 0; c < ns; ++c) {
        if (s[1 + 2 * c] != comp_id[c]) { o.status = SIB_JPEG_UNSUPPORTED_SCANS; return 0; }
        pj->dc_sel[c] = s[2 + 2 * c] >> 4;
        pj->ac_sel[c] = s[2 + 2 * c] & 15;
        if (pj->dc_sel[c] > 3 || pj->ac_sel[c] > 3 || !pj->dc[pj->dc_sel[c]].present ||
            !pj->ac[pj->ac_sel[c]].present) { o.status = SIB_JPEG_CORRUPT; return 0; }
      }
      pj->scan_begin = p + len;
      // colour space as libjpeg's default_decompress_parms decides it for three components
      if (o.ncomp == 3) {
        bool ycc;
        if (jfif) ycc = true;
        else if (adobe) ycc = adobe_transform == 1;
        else ycc = comp_id[0] == 1 && comp_id[1] == 2 && comp_id[2] == 3;
        if (!ycc || (adobe && adobe_transform != 1 && !jfif)) { o.status = SIB_JPEG_UNSUPPORTED_COLORSPACE; return 0; }
      }
      if (o.ncomp == 1) { o.hs[0] = o.vs[0] = 1; }   // a single-component scan ignores the sampling factors
      int hmax = 1, vmax = 1;
      for (int c = 0; c < o.ncomp; ++c) {
        if (o.hs[c] < 1 || o.hs[c] > 2 || o.vs[c] < 1 || o.vs[c] > 2) { o.status = SIB_JPEG_UNSUPPORTED_SAMPLING; return 0; }
        hmax = o.hs[c] > hmax ? o.hs[c] : hmax;
        vmax = o.vs[c] > vmax ? o.vs[c] : vmax;
        if (!have_q[comp_q[c]]) { o.status = SIB_JPEG_CORRUPT; return 0; }
        memcpy(o.quant[c], qt[comp_q[c]], sizeof(qt[0]));
      }
      if (o.ncomp == 3) {
        // luma at full resolution; chroma 1x1 (4:4:4), 2x1 (4:2:2) or 2x2 (4:2:0) subsampled, both alike
        const bool chroma_ok = o.hs[1] == 1 && o.vs[1] == 1 && o.hs[2] == 1 && o.vs[2] == 1;
        const bool luma_ok = o.hs[0] == hmax && o.vs[0] == vmax && !(hmax == 1 && vmax == 2);
        if (!chroma_ok || !luma_ok) { o.status = SIB_JPEG_UNSUPPORTED_SAMPLING; return 0; }
        // libjpeg only takes the fancy (triangle) path when the chroma rows hold more than 2 samples
        if (hmax == 2 && (o.width + 1) / 2 <= 2) { o.status = SIB_JPEG_UNSUPPORTED_SAMPLING; return 0; }
      }
      o.hmax = hmax;
      o.vmax = vmax;
      o.mcus_x = (o.width + 8 * hmax - 1) / (8 * hmax);
      o.mcus_y = (o.height + 8 * vmax - 1) / (8 * vmax);
      o.coef_count = 0;
      for (int c = 0; c < o.ncomp; ++c) {
        o.blocks_w[c] = o.mcus_x * o.hs[c];
        o.blocks_h[c] = o.mcus_y * o.vs[c];
        o.coef_count += (long)o.blocks_w[c] * o.blocks_h[c] * 64;
      }
      o.status = SIB_JPEG_OK;
      return 0;
    }
    p += len;
  }
  o.status = SIB_JPEG_CORRUPT;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Huffman decoding (ITU T.81 F.2.2).  Three tables per Huffman table: canonical (maxcode / valoffset,
// for codes longer than kLook bits), a kLook-bit look-ahead table (code length + symbol), and for AC
// tables a kLook-bit "whole coefficient" table: when the code AND the value bits that follow it fit in
// kLook bits the entry holds (coefficient, run, total bits) and one lookup decodes the coefficient --
// the common case for the small coefficients that make up most of a stream.
// ---------------------------------------------------------------------------------------------
constexpr int kLook = 10;
struct HuffTable {
  int maxcode[18];          // largest code of length l (-1 if none); [17] = sentinel
  int valoffset[17];
  unsigned short look[1 << kLook];   // (length << 8) | symbol, 0 = longer than kLook bits
  short fast_ac[1 << kLook];         // (value << 8) | (run << 4) | bits, 0 = take the slow path
  unsigned char vals[256];
};

static inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

static bool build_table(const HuffSpec& s, HuffTable* t, bool is_ac) {
  int huffsize[257], huffcode[257];
  int k = 0;
  for (int l = 1; l <= 16; ++l)
    for (int i = 0; i < s.bits[l]; ++i) huffsize[k++] = l;
  huffsize[k] = 0;
  int code = 0, si = huffsize[0];
  k = 0;
  while (huffsize[k]) {
    while (huffsize[k] == si) huffcode[k++] = code++;
    if (code > (1 << si)) return false;
    code <<= 1;
    ++si;
  }
  int p = 0;
  for (int l = 1; l <= 16; ++l) {
    if (s.bits[l]) {
      t->valoffset[l] = p - huffcode[p];
      p += s.bits[l];
      t->maxcode[l] = huffcode[p - 1];
    } else {
      t->maxcode[l] = -1;
      t->valoffset[l] = 0;
    }
  }
  t->maxcode[17] = 0xFFFFF;
  memcpy(t->vals, s.vals, 256);
  memset(t->look, 0, sizeof(t->look));
  memset(t->fast_ac, 0, sizeof(t->fast_ac));
  p = 0;
  for (int l = 1; l <= kLook; ++l)
    for (int i = 0; i < s.bits[l]; ++i, ++p) {
      const int first = huffcode[p] << (kLook - l);
      for (int j = 0; j < (1 << (kLook - l)); ++j) t->look[first + j] = (unsigned short)((l << 8) | s.vals[p]);
    }
  if (is_ac) {
    for (int i = 0; i < (1 << kLook); ++i) {
      const int e = t->look[i];
      if (!e) continue;
      const int len = e >> 8, rs = e & 0xFF, run = rs >> 4, size = rs & 15;
      if (size == 0 || len + size > kLook) continue;
      const int v = extend((i >> (kLook - len - size)) & ((1 << size) - 1), size);
      if (v >= -128 && v <= 127) t->fast_ac[i] = (short)((v * 256) | (run << 4) | (len + size));
    }
  }
  return true;
}

struct BitReader {
  const unsigned char* d;
  long pos, end;
  unsigned long long acc;   // the low `count` bits are the unread bits, oldest first
  int count;
  bool hit_marker;
  inline void fill() {
    if (!hit_marker && pos + 8 <= end && count <= 56) {
      // eight bytes at once when none of them is 0xFF (no stuffing, no marker)
      unsigned long long w;
      memcpy(&w, d + pos, 8);
      w = __builtin_bswap64(w);
      const unsigned long long nw = ~w;
      if (!((nw - 0x0101010101010101ull) & ~nw & 0x8080808080808080ull)) {
        const int nbytes = (64 - count) >> 3;
        acc = nbytes == 8 ? w : ((acc << (8 * nbytes)) | (w >> (64 - 8 * nbytes)));
        pos += nbytes;
        count += 8 * nbytes;
        return;
      }
    }
    while (count <= 56) {
      int b = 0;
      if (!hit_marker && pos < end) {
        b = d[pos];
        if (b == 0xFF) {
          const int b2 = pos + 1 < end ? d[pos + 1] : 0xD9;
          if (b2 == 0) pos += 2;                    // stuffed zero
          else { hit_marker = true; b = 0; }         // a marker: feed zeros, leave pos on it
        } else {
          ++pos;
        }
      } else {
        hit_marker = true;
      }
      acc = (acc << 8) | (unsigned)b;
      count += 8;
    }
  }
  inline int peek(int n) { return (int)((acc >> (count - n)) & ((1u << n) - 1)); }
  inline void skip(int n) { count -= n; }
  inline int get(int n) { const int v = peek(n); count -= n; return v; }
};

// (the caller guarantees count >= 16)
static inline int decode_symbol(BitReader& br, const HuffTable& t) {
  const int look = t.look[br.peek(kLook)];
  if (look) {
    br.skip(look >> 8);
    return look & 0xFF;
  }
  int l = kLook + 1;
  int code = br.peek(l);
  while (l <= 16 && code > t.maxcode[l]) { ++l; code = br.peek(l); }
  if (l > 16) { br.skip(16); return 0; }            // corrupt code: behave like libjpeg (symbol 0)
  br.skip(l);
  return t.vals[(code + t.valoffset[l]) & 0xFF];
}

}  // namespace sib

using namespace sib;

extern "C" int sib_jpeg_parse(const unsigned char* data, long size, sib_jpeg_info* info) {
  SIB_CHECK(data != nullptr && info != nullptr, "jpeg_parse: null argument");
  ParsedJpeg pj;
  parse_jpeg(data, size, &pj);
  *info = pj.info;
  return 0;
}

extern "C" int sib_jpeg_decode_coefficients(const unsigned char* data, long size, short* coef) {
  return sib_jpeg_decode_coefficients_rows(data, size, coef, 0);
}

// mcu_rows > 0: decode only the first mcu_rows rows of MCUs (a random crop that ends above the bottom of the
// image does not need the rest: entropy-coded data cannot be skipped, but it need not be read to the end --
// what fn.decoders.image_random_crop's ROI decoding saves).  Coefficients of later rows are left untouched.
extern "C" int sib_jpeg_decode_coefficients_rows(const unsigned char* data, long size, short* coef, int mcu_rows) {
  SIB_CHECK(data != nullptr && coef != nullptr, "jpeg_decode_coefficients: null argument");
  ParsedJpeg pj;
  parse_jpeg(data, size, &pj);
  const sib_jpeg_info& o = pj.info;
  SIB_CHECK(o.status == SIB_JPEG_OK, "jpeg_decode_coefficients: stream not decodable here (status %d)", o.status);
  HuffTable* dc[3];
  HuffTable* ac[3];
  HuffTable tabs[8];
  bool built[8] = {false, false, false, false, false, false, false, false};
  for (int c = 0; c < o.ncomp; ++c) {
    const int di = pj.dc_sel[c], ai = 4 + pj.ac_sel[c];
    if (!built[di]) { SIB_CHECK(build_table(pj.dc[di], &tabs[di], false), "jpeg: bad DC Huffman table"); built[di] = true; }
    if (!built[ai]) { SIB_CHECK(build_table(pj.ac[ai - 4], &tabs[ai], true), "jpeg: bad AC Huffman table"); built[ai] = true; }
    dc[c] = &tabs[di];
    ac[c] = &tabs[ai];
  }
  const int rows_end = (mcu_rows > 0 && mcu_rows < o.mcus_y) ? mcu_rows : o.mcus_y;
  short* base[3];
  long off = 0;
  for (int c = 0; c < o.ncomp; ++c) {
    base[c] = coef + off;
    memset(base[c], 0, sizeof(short) * (long)o.blocks_w[c] * (rows_end * o.vs[c]) * 64);
    off += (long)o.blocks_w[c] * o.blocks_h[c] * 64;
  }
  BitReader br{data, pj.scan_begin, size, 0ull, 0, false};
  int pred[3] = {0, 0, 0};
  int until_restart = o.restart_interval;
  int next_rst = 0;
  for (int my = 0; my < rows_end; ++my) {
    for (int mx = 0; mx < o.mcus_x; ++mx) {
      if (o.restart_interval && until_restart == 0) {
        // byte-align, expect RSTn, reset the predictors
        br.count = 0;
        br.acc = 0;
        br.hit_marker = false;
        while (br.pos + 1 < br.end && !(br.d[br.pos] == 0xFF && br.d[br.pos + 1] >= 0xD0 && br.d[br.pos + 1] <= 0xD7)) {
          if (br.d[br.pos] == 0xFF && br.d[br.pos + 1] != 0 && br.d[br.pos + 1] != 0xFF) break;   // another marker: give up resync
          ++br.pos;
        }
        if (br.pos + 1 < br.end && br.d[br.pos] == 0xFF && br.d[br.pos + 1] == 0xD0 + next_rst) br.pos += 2;
        next_rst = (next_rst + 1) & 7;
        pred[0] = pred[1] = pred[2] = 0;
        until_restart = o.restart_interval;
      }
      for (int c = 0; c < o.ncomp; ++c) {
        for (int v = 0; v < o.vs[c]; ++v) {
          for (int h = 0; h < o.hs[c]; ++h) {
            short* blk = base[c] + ((long)(my * o.vs[c] + v) * o.blocks_w[c] + (mx * o.hs[c] + h)) * 64;
            if (br.count < 32) br.fill();
            int s = decode_symbol(br, *dc[c]);
            if (s) {
              if (br.count < s) br.fill();
              pred[c] += extend(br.get(s), s);
            }
            blk[0] = (short)pred[c];
            const HuffTable& at = *ac[c];
            for (int k = 1; k < 64; ++k) {
              if (br.count < 32) br.fill();            // >= 16 bits for the code + 16 for the value
              const int fa = at.fast_ac[br.peek(kLook)];
              if (fa) {                                 // code and value in one lookup
                k += (fa >> 4) & 15;
                br.skip(fa & 15);
                blk[kNatural[k]] = (short)(fa >> 8);
                continue;
              }
              const int rs = decode_symbol(br, at);
              const int r = rs >> 4;
              s = rs & 15;
              if (s) {
                k += r;
                blk[kNatural[k]] = (short)extend(br.get(s), s);
              } else {
                if (r != 15) break;
                k += 15;
              }
            }
          }
        }
      }
      if (o.restart_interval) --until_restart;
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// device: dequantise + inverse DCT + upsample + colour conversion
// ---------------------------------------------------------------------------------------------
namespace sib {

__device__ __forceinline__ int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// jidctint.c jpeg_idct_islow: CONST_BITS = 13, PASS1_BITS = 2
#define FIX_0_298631336 2446
#define FIX_0_390180644 3196
#define FIX_0_541196100 4433
#define FIX_0_765366865 6270
#define FIX_0_899976223 7373
#define FIX_1_175875602 9633
#define FIX_1_501321110 12299
#define FIX_1_847759065 15137
#define FIX_1_961570560 16069
#define FIX_2_053119869 16819
#define FIX_2_562915447 20995
#define FIX_3_072711026 25172

// one 1-D pass over (d0 .. d7); results descaled by `shift` with rounding
__device__ __forceinline__ void idct_1d(int d0, int d1, int d2, int d3, int d4, int d5, int d6, int d7,
                                        int shift, int* o) {
  // even part
  int z2 = d2, z3 = d6;
  int z1 = (z2 + z3) * FIX_0_541196100;
  int tmp2 = z1 + z3 * (-FIX_1_847759065);
  int tmp3 = z1 + z2 * FIX_0_765366865;
  z2 = d0; z3 = d4;
  int tmp0 = (z2 + z3) << 13;
  int tmp1 = (z2 - z3) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  // odd part
  tmp0 = d7; tmp1 = d5; tmp2 = d3; tmp3 = d1;
  z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * FIX_1_175875602;
  tmp0 *= FIX_0_298631336; tmp1 *= FIX_2_053119869; tmp2 *= FIX_3_072711026; tmp3 *= FIX_1_501321110;
  z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
  z3 += z5; z4 += z5;
  tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
  const int rnd = 1 << (shift - 1);
  o[0] = (tmp10 + tmp3 + rnd) >> shift;
  o[7] = (tmp10 - tmp3 + rnd) >> shift;
  o[1] = (tmp11 + tmp2 + rnd) >> shift;
  o[6] = (tmp11 - tmp2 + rnd) >> shift;
  o[2] = (tmp12 + tmp1 + rnd) >> shift;
  o[5] = (tmp12 - tmp1 + rnd) >> shift;
  o[3] = (tmp13 + tmp0 + rnd) >> shift;
  o[4] = (tmp13 - tmp0 + rnd) >> shift;
}

// One thread per 8x8 block: 64 coefficients (a contiguous 128-byte line) -> 64 samples of its plane.
__global__ void __launch_bounds__(128)
jpeg_idct_kernel(const short* __restrict__ coef, const sib_jpeg_image* __restrict__ images,
                 unsigned char* __restrict__ planes) {
  const sib_jpeg_image& im = images[blockIdx.y];
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  int c = 0;
  for (; c < im.ncomp; ++c) {
    const int nb = im.blocks_w[c] * im.blocks_h[c];
    if (b < nb) break;
    b -= nb;
  }
  if (c >= im.ncomp) return;
  const int by = b / im.blocks_w[c], bx = b - by * im.blocks_w[c];
  // row-limited decode: block rows past the decoded MCU rows hold no coefficients (vs_c = blocks_h[c] * vmax / blocks_h[0])
  if (im.mcu_rows > 0 && by >= im.mcu_rows * (im.blocks_h[c] * im.vmax / im.blocks_h[0])) return;
  const uint4* src = reinterpret_cast<const uint4*>(coef + im.coef_off[c] + (long)b * 64);
  const unsigned short* q = im.quant[c];
  int ws[64];
  // pass 1: columns (all eight rows of the block are loaded as 8 x 16-byte vectors first)
  __align__(16) short cf[64];
#pragma unroll
  for (int r = 0; r < 8; ++r) *reinterpret_cast<uint4*>(&cf[r * 8]) = __ldg(src + r);
#pragma unroll
  for (int col = 0; col < 8; ++col) {
    int o[8];
    idct_1d(cf[col] * q[col], cf[8 + col] * q[8 + col], cf[16 + col] * q[16 + col], cf[24 + col] * q[24 + col],
            cf[32 + col] * q[32 + col], cf[40 + col] * q[40 + col], cf[48 + col] * q[48 + col],
            cf[56 + col] * q[56 + col], 13 - 2, o);
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[r * 8 + col] = o[r];
  }
  // pass 2: rows, + level shift, clamp
  const int stride = im.blocks_w[c] * 8;
  unsigned char* dst = planes + im.plane_off[c] + (long)(by * 8) * stride + bx * 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int o[8];
    idct_1d(ws[r * 8], ws[r * 8 + 1], ws[r * 8 + 2], ws[r * 8 + 3], ws[r * 8 + 4], ws[r * 8 + 5], ws[r * 8 + 6],
            ws[r * 8 + 7], 13 + 2 + 3, o);
    uint2 pk;
    pk.x = clamp255(o[0] + 128) | (clamp255(o[1] + 128) << 8) | (clamp255(o[2] + 128) << 16) | (clamp255(o[3] + 128) << 24);
    pk.y = clamp255(o[4] + 128) | (clamp255(o[5] + 128) << 8) | (clamp255(o[6] + 128) << 16) | (clamp255(o[7] + 128) << 24);
    *reinterpret_cast<uint2*>(dst + (long)r * stride) = pk;
  }
}

// chroma sample at full resolution (jdsample.c): replicate (1x1), h2v1 / h2v2 "fancy" triangle filter.
// (dw, dh) = real extent of the subsampled plane; rows / columns past it replicate the last real one.
__device__ __forceinline__ int chroma_at(const unsigned char* __restrict__ pl, int stride, int dw, int dh,
                                         int hmax, int vmax, int x, int y) {
  if (hmax == 1) return pl[(long)y * stride + x];
  const int cx = x >> 1;
  if (vmax == 1) {
    const unsigned char* row = pl + (long)y * stride;
    const int v = row[cx];
    if (x & 1) return cx == dw - 1 ? v : (3 * v + row[cx + 1] + 2) >> 2;
    return cx == 0 ? v : (3 * v + row[cx - 1] + 1) >> 2;
  }
  const int cy = y >> 1;
  int ny = (y & 1) ? cy + 1 : cy - 1;           // nearer neighbouring row (edges replicate)
  ny = ny < 0 ? 0 : (ny > dh - 1 ? dh - 1 : ny);
  const unsigned char* r0 = pl + (long)cy * stride;
  const unsigned char* r1 = pl + (long)ny * stride;
  const int cur = 3 * r0[cx] + r1[cx];
  if (x & 1) {
    if (cx == dw - 1) return (cur * 4 + 7) >> 4;
    return (3 * cur + 3 * r0[cx + 1] + r1[cx + 1] + 7) >> 4;
  }
  if (cx == 0) return (cur * 4 + 8) >> 4;
  return (3 * cur + 3 * r0[cx - 1] + r1[cx - 1] + 8) >> 4;
}

// Y + upsampled Cb / Cr -> RGB (jdcolor.c, 16-bit fixed point)
__device__ __forceinline__ uint32_t jpeg_pixel(const sib_jpeg_image& im, const unsigned char* __restrict__ planes,
                                               int i, int dw, int dh) {
  const int y = i / im.width, x = i - y * im.width;
  const int yy = planes[im.plane_off[0] + (long)y * (im.blocks_w[0] * 8) + x];
  if (im.ncomp == 1) return (uint32_t)yy * 0x010101u;
  const int cb = chroma_at(planes + im.plane_off[1], im.blocks_w[1] * 8, dw, dh, im.hmax, im.vmax, x, y) - 128;
  const int cr = chroma_at(planes + im.plane_off[2], im.blocks_w[2] * 8, dw, dh, im.hmax, im.vmax, x, y) - 128;
  // FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.71414) = 46802, FIX(0.34414) = 22554, ONE_HALF = 32768
  const int r = yy + ((91881 * cr + 32768) >> 16);
  const int g = yy + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
  const int b = yy + ((116130 * cb + 32768) >> 16);
  return (uint32_t)clamp255(r) | ((uint32_t)clamp255(g) << 8) | ((uint32_t)clamp255(b) << 16);
}

__device__ __forceinline__ uint32_t ycc_pixel(int yy, int cb, int cr) {
  const int r = yy + ((91881 * cr + 32768) >> 16);
  const int g = yy + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
  const int b = yy + ((116130 * cb + 32768) >> 16);
  return (uint32_t)clamp255(r) | ((uint32_t)clamp255(g) << 8) | ((uint32_t)clamp255(b) << 16);
}

// Four consecutive pixels (x0 .. x0+3) of ONE row: the chroma samples they interpolate between are
// cxa .. cxa+3 with cxa = (x0 >> 1) - 1 whatever the parity of x0; the triangle filter's edge cases are
// exactly "the missing neighbour replicates the sample" ((3v + v + 1) >> 2 = v, (4s + 8) >> 4 = the special
// first-column formula, ...), so indices are clamped to the real extent and no case split is needed.
__device__ __forceinline__ void chroma4(const unsigned char* __restrict__ pl, int stride, int dw, int dh, int hmax,
                                        int vmax, int x0, int y, int* out) {
  if (hmax == 1) {
    const unsigned char* row = pl + (long)y * stride + x0;
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = row[j];
    return;
  }
  const int cxa = (x0 >> 1) - 1;
  int cs[4];
  if (vmax == 1) {
    const unsigned char* row = pl + (long)y * stride;
#pragma unroll
    for (int j = 0; j < 4; ++j) cs[j] = row[min(max(cxa + j, 0), dw - 1)];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = x0 + j, k = (x >> 1) - cxa;            // k in 1 .. 3, k + 1 <= 3 for odd x, k - 1 >= 0 for even x
      out[j] = (x & 1) ? (3 * cs[k] + cs[k + 1] + 2) >> 2 : (3 * cs[k] + cs[k - 1] + 1) >> 2;
    }
    return;
  }
  const int cy = y >> 1;
  const int ny = min(max((y & 1) ? cy + 1 : cy - 1, 0), dh - 1);
  const unsigned char* r0 = pl + (long)cy * stride;
  const unsigned char* r1 = pl + (long)ny * stride;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = min(max(cxa + j, 0), dw - 1);
    cs[j] = 3 * r0[c] + r1[c];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = x0 + j, k = (x >> 1) - cxa;
    out[j] = (x & 1) ? (3 * cs[k] + cs[k + 1] + 7) >> 4 : (3 * cs[k] + cs[k - 1] + 8) >> 4;
  }
}

// one thread per FOUR consecutive pixels of the packed [H][W][3] image: 12 output bytes = three aligned
// 32-bit stores (every image starts on a 16-byte boundary).  When the four pixels lie in one row (all but
// one group per row) the chroma interpolation shares its column sums; groups that straddle a row end and
// the last 1-3 pixels of the image take the per-pixel route.
__global__ void __launch_bounds__(256)
jpeg_rgb_kernel(const sib_jpeg_image* __restrict__ images, const unsigned char* __restrict__ planes,
                unsigned char* __restrict__ out) {
  const sib_jpeg_image& im = images[blockIdx.y];
  // row-limited decode: only the decoded luma rows are converted (the last two of them interpolate against
  // an undecoded chroma row and are inexact; the caller asks for two rows more than it reads)
  int rows = im.height;
  if (im.mcu_rows > 0 && im.mcu_rows * 8 * im.vmax < rows) rows = im.mcu_rows * 8 * im.vmax;
  const int W = im.width, ncomp = im.ncomp, hmax = im.hmax, vmax = im.vmax;
  const int npx = W * rows;
  const int dw = (W + hmax - 1) / hmax, dh = (im.height + vmax - 1) / vmax;
  const int ystride = im.blocks_w[0] * 8, cstride = im.blocks_w[1] * 8;
  const unsigned char* py = planes + im.plane_off[0];
  const unsigned char* pcb = planes + im.plane_off[1];
  const unsigned char* pcr = planes + im.plane_off[2];
  unsigned char* base = out + im.out_off;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q * 4 < npx; q += gridDim.x * blockDim.x) {
    const int i0 = q * 4;
    const int y = i0 / W, x0 = i0 - y * W;
    uint32_t p[4];
    if (x0 + 4 <= W) {
      const unsigned char* yrow = py + (long)y * ystride + x0;
      if (ncomp == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = (uint32_t)yrow[j] * 0x010101u;
      } else {
        int cb[4], cr[4];
        chroma4(pcb, cstride, dw, dh, hmax, vmax, x0, y, cb);
        chroma4(pcr, cstride, dw, dh, hmax, vmax, x0, y, cr);
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = ycc_pixel(yrow[j], cb[j] - 128, cr[j] - 128);
      }
    } else if (i0 + 4 <= npx) {
#pragma unroll
      for (int j = 0; j < 4; ++j) p[j] = jpeg_pixel(im, planes, i0 + j, dw, dh);
    } else {
      for (int i = i0; i < npx; ++i) {
        const uint32_t px = jpeg_pixel(im, planes, i, dw, dh);
        base[(long)i * 3] = (unsigned char)px;
        base[(long)i * 3 + 1] = (unsigned char)(px >> 8);
        base[(long)i * 3 + 2] = (unsigned char)(px >> 16);
      }
      continue;
    }
    uint32_t* o = reinterpret_cast<uint32_t*>(base + (long)i0 * 3);
    o[0] = p[0] | (p[1] << 24);
    o[1] = (p[1] >> 8) | (p[2] << 16);
    o[2] = (p[2] >> 16) | (p[3] << 8);
  }
}

}  // namespace sib

extern "C" int sib_jpeg_idct_rgb(const short* coef_dev, const sib_jpeg_image* images_dev, int B, int max_blocks,
                                 int max_pixels, unsigned char* planes_dev, unsigned char* out_dev, void* stream) {
  if (B <= 0) return 0;
  SIB_CHECK(coef_dev && images_dev && planes_dev && out_dev, "jpeg_idct_rgb: null argument");
  SIB_CHECK(max_blocks > 0 && max_pixels > 0, "jpeg_idct_rgb: empty batch geometry");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  jpeg_idct_kernel<<<dim3((max_blocks + 127) / 128, B), 128, 0, st>>>(coef_dev, images_dev, planes_dev);
  SIB_LAUNCH_CHECK();
  int gx = (max_pixels / 4 + 256) / 256;
  if (gx > 1024) gx = 1024;
  jpeg_rgb_kernel<<<dim3(gx, B), 256, 0, st>>>(images_dev, planes_dev, out_dev);
  SIB_LAUNCH_CHECK();
  return 0;
}
