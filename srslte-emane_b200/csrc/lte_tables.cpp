// lte_tables.cpp -- see lte_tables.h
#include "lte_tables.h"

#include <cmath>
#include <cstring>

namespace b200 {

const QppEntry kQpp[kNofCbSizes] = {
  {  40,   3,  10},
  {  48,   7,  12},
  {  56,  19,  42},
  {  64,   7,  16},
  {  72,   7,  18},
  {  80,  11,  20},
  {  88,   5,  22},
  {  96,  11,  24},
  { 104,   7,  26},
  { 112,  41,  84},
  { 120, 103,  90},
  { 128,  15,  32},
  { 136,   9,  34},
  { 144,  17, 108},
  { 152,   9,  38},
  { 160,  21, 120},
  { 168, 101,  84},
  { 176,  21,  44},
  { 184,  57,  46},
  { 192,  23,  48},
  { 200,  13,  50},
  { 208,  27,  52},
  { 216,  11,  36},
  { 224,  27,  56},
  { 232,  85,  58},
  { 240,  29,  60},
  { 248,  33,  62},
  { 256,  15,  32},
  { 264,  17, 198},
  { 272,  33,  68},
  { 280, 103, 210},
  { 288,  19,  36},
  { 296,  19,  74},
  { 304,  37,  76},
  { 312,  19,  78},
  { 320,  21, 120},
  { 328,  21,  82},
  { 336, 115,  84},
  { 344, 193,  86},
  { 352,  21,  44},
  { 360, 133,  90},
  { 368,  81,  46},
  { 376,  45,  94},
  { 384,  23,  48},
  { 392, 243,  98},
  { 400, 151,  40},
  { 408, 155, 102},
  { 416,  25,  52},
  { 424,  51, 106},
  { 432,  47,  72},
  { 440,  91, 110},
  { 448,  29, 168},
  { 456,  29, 114},
  { 464, 247,  58},
  { 472,  29, 118},
  { 480,  89, 180},
  { 488,  91, 122},
  { 496, 157,  62},
  { 504,  55,  84},
  { 512,  31,  64},
  { 528,  17,  66},
  { 544,  35,  68},
  { 560, 227, 420},
  { 576,  65,  96},
  { 592,  19,  74},
  { 608,  37,  76},
  { 624,  41, 234},
  { 640,  39,  80},
  { 656, 185,  82},
  { 672,  43, 252},
  { 688,  21,  86},
  { 704, 155,  44},
  { 720,  79, 120},
  { 736, 139,  92},
  { 752,  23,  94},
  { 768, 217,  48},
  { 784,  25,  98},
  { 800,  17,  80},
  { 816, 127, 102},
  { 832,  25,  52},
  { 848, 239, 106},
  { 864,  17,  48},
  { 880, 137, 110},
  { 896, 215, 112},
  { 912,  29, 114},
  { 928,  15,  58},
  { 944, 147, 118},
  { 960,  29,  60},
  { 976,  59, 122},
  { 992,  65, 124},
  {1008,  55,  84},
  {1024,  31,  64},
  {1056,  17,  66},
  {1088, 171, 204},
  {1120,  67, 140},
  {1152,  35,  72},
  {1184,  19,  74},
  {1216,  39,  76},
  {1248,  19,  78},
  {1280, 199, 240},
  {1312,  21,  82},
  {1344, 211, 252},
  {1376,  21,  86},
  {1408,  43,  88},
  {1440, 149,  60},
  {1472,  45,  92},
  {1504,  49, 846},
  {1536,  71,  48},
  {1568,  13,  28},
  {1600,  17,  80},
  {1632,  25, 102},
  {1664, 183, 104},
  {1696,  55, 954},
  {1728, 127,  96},
  {1760,  27, 110},
  {1792,  29, 112},
  {1824,  29, 114},
  {1856,  57, 116},
  {1888,  45, 354},
  {1920,  31, 120},
  {1952,  59, 610},
  {1984, 185, 124},
  {2016, 113, 420},
  {2048,  31,  64},
  {2112,  17,  66},
  {2176, 171, 136},
  {2240, 209, 420},
  {2304, 253, 216},
  {2368, 367, 444},
  {2432, 265, 456},
  {2496, 181, 468},
  {2560,  39,  80},
  {2624,  27, 164},
  {2688, 127, 504},
  {2752, 143, 172},
  {2816,  43,  88},
  {2880,  29, 300},
  {2944,  45,  92},
  {3008, 157, 188},
  {3072,  47,  96},
  {3136,  13,  28},
  {3200, 111, 240},
  {3264, 443, 204},
  {3328,  51, 104},
  {3392,  51, 212},
  {3456, 451, 192},
  {3520, 257, 220},
  {3584,  57, 336},
  {3648, 313, 228},
  {3712, 271, 232},
  {3776, 179, 236},
  {3840, 331, 120},
  {3904, 363, 244},
  {3968, 375, 248},
  {4032, 127, 168},
  {4096,  31,  64},
  {4160,  33, 130},
  {4224,  43, 264},
  {4288,  33, 134},
  {4352, 477, 408},
  {4416,  35, 138},
  {4480, 233, 280},
  {4544, 357, 142},
  {4608, 337, 480},
  {4672,  37, 146},
  {4736,  71, 444},
  {4800,  71, 120},
  {4864,  37, 152},
  {4928,  39, 462},
  {4992, 127, 234},
  {5056,  39, 158},
  {5120,  39,  80},
  {5184,  31,  96},
  {5248, 113, 902},
  {5312,  41, 166},
  {5376, 251, 336},
  {5440,  43, 170},
  {5504,  21,  86},
  {5568,  43, 174},
  {5632,  45, 176},
  {5696,  45, 178},
  {5760, 161, 120},
  {5824,  89, 182},
  {5888, 323, 184},
  {5952,  47, 186},
  {6016,  23,  94},
  {6080,  47, 190},
  {6144, 263, 480},
};

int cb_index_ceil(uint32_t long_cb)
{
  for (int i = 0; i < kNofCbSizes; i++)
    if (kQpp[i].K >= long_cb) return i;
  return -1;
}

int cb_index_exact(uint32_t K)
{
  // sizes are 40..512 step 8, 528..1024 step 16, 1056..2048 step 32, 2112..6144 step 64
  if (K < 40 || K > kMaxK) return -1;
  int i;
  if (K <= 512) {
    if (K % 8) return -1;
    i = (int)(K - 40) / 8;
  } else if (K <= 1024) {
    if (K % 16) return -1;
    i = 60 + (int)(K - 528) / 16;
  } else if (K <= 2048) {
    if (K % 32) return -1;
    i = 92 + (int)(K - 1056) / 32;
  } else {
    if (K % 64) return -1;
    i = 124 + (int)(K - 2112) / 64;
  }
  return (i >= 0 && i < kNofCbSizes && kQpp[i].K == K) ? i : -1;
}

int cb_size(uint32_t idx) { return idx < (uint32_t)kNofCbSizes ? (int)kQpp[idx].K : -1; }

int nof_windows(uint32_t K)
{
  if (K % 16 == 0 && K > 800) return 16;
  if (K % 8 == 0 && K > 400) return 8;
  return 0;
}

uint32_t working_len(uint32_t K) { return nof_windows(K) ? 3 * (K + 32) + 12 : 3 * K + 12; }

int cbsegm(CbSegm* s, uint32_t tbs)
{
  std::memset(s, 0, sizeof(*s));
  if (tbs == 0) return 0;
  const uint32_t B = tbs + 24;
  uint32_t       Bp;
  s->tbs = tbs;
  if (B <= kMaxK) {
    s->C = 1;
    Bp   = B;
  } else {
    s->C = (uint32_t)std::ceil((float)B / (float)(kMaxK - 24));
    Bp   = B + 24 * s->C;
  }
  const int i1 = cb_index_ceil((Bp - 1) / s->C + 1);
  if (i1 < 0) return -1;
  s->K1     = kQpp[i1].K;
  s->K1_idx = (uint32_t)i1;
  if (s->C == 1) {
    s->C1 = 1;
  } else {
    if (i1 == 0) return -1;
    s->K2     = kQpp[i1 - 1].K;
    s->K2_idx = (uint32_t)i1 - 1;
    s->C2     = (s->C * s->K1 - Bp) / (s->K1 - s->K2);
    s->C1     = s->C - s->C2;
  }
  s->F = s->C1 * s->K1 + s->C2 * s->K2 - Bp;
  return 0;
}

// ---- rate de-matching ---------------------------------------------------------------------------
// 36.212 5.1.4.1.1 inter-column permutation of the sub-block interleaver
static const uint8_t kColPerm[32] = {0, 16, 8,  24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                                     1, 17, 9,  25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};

void rm_rx_table(uint32_t K, uint32_t rv, bool sb_layout, std::vector<uint16_t>& table)
{
  const int D   = (int)K + 4;          // length of each of the three d-streams
  const int R   = (D - 1) / 32 + 1;    // interleaver rows
  const int KP  = 32 * R;
  const int ND  = KP - D;              // dummy (<NULL>) entries at the head of every stream
  const int Ncb = 3 * KP;
  const int N   = 3 * (int)K + 12;
  const int W   = sb_layout ? nof_windows(K) : 0;
  const int L   = W ? (int)K / W : 0;
  const int k0  = R * (2 * (int)std::ceil((float)Ncb / (float)(8 * R)) * (int)rv + 2);

  table.resize(N);
  int filled = 0;
  for (int j = 0; filled < N; j++) {
    const int pos = (k0 + j) % Ncb;  // position in the virtual circular buffer
    int       stream, d;
    if (pos < KP) {
      stream = 0;
      d      = (pos % R) * 32 + kColPerm[pos / R] - ND;
    } else if (((pos - KP) & 1) == 0) {
      const int q = (pos - KP) >> 1;
      stream      = 1;
      d           = (q % R) * 32 + kColPerm[q / R] - ND;
    } else {
      const int q = (pos - KP - 1) >> 1;
      stream      = 2;
      d           = (kColPerm[q / R] + 32 * (q % R) + 1) % KP - ND;
    }
    if (d < 0) continue;  // dummy: nothing was transmitted for it
    uint32_t idx;
    if (W == 0) {
      idx = 3 * (uint32_t)d + (uint32_t)stream;
    } else if (d >= (int)K) {
      idx = 3 * (K + 32) + 3 * (uint32_t)(d - (int)K) + (uint32_t)stream;  // 12 tail samples keep their order
    } else {
      idx = (uint32_t)stream * (K + 32) + (uint32_t)(d % L) * (uint32_t)W + (uint32_t)(d / L);
    }
    table[filled++] = (uint16_t)idx;
  }
}

// ---- CRC ------------------------------------------------------------------------------------------
void crc24_table(uint32_t poly, uint32_t table[256])
{
  for (uint32_t b = 0; b < 256; b++) {
    uint32_t r = b << 16;
    for (int i = 0; i < 8; i++) r = (r & 0x800000u) ? ((r << 1) ^ poly) : (r << 1);
    table[b] = r & 0xFFFFFFu;
  }
}

// Slicing-by-8: eight table look-ups advance the register by eight bytes at once (the byte-serial form is a chain
// of dependent look-ups, 3 ns per byte; transport blocks of 75 kbit are checked on the host after every decode).
uint32_t crc24_bytes(uint32_t poly, const uint8_t* data, uint32_t nbytes)
{
  struct Tables {
    uint32_t poly = 0;
    uint32_t t[8][256];
  };
  static thread_local Tables cache[2];
  Tables* T = nullptr;
  for (auto& c : cache)
    if (c.poly == poly) T = &c;
  if (!T) {
    T = cache[0].poly == 0 ? &cache[0] : &cache[1];
    T->poly = poly;
    crc24_table(poly, T->t[0]);
    // t[k][b] = register after byte b followed by k zero bytes
    for (int k = 1; k < 8; k++)
      for (uint32_t b = 0; b < 256; b++) {
        const uint32_t r = T->t[k - 1][b];
        T->t[k][b] = ((r << 8) ^ T->t[0][(r >> 16) & 0xFFu]) & 0xFFFFFFu;
      }
  }
  uint32_t crc = 0, i = 0;
  for (; i + 8 <= nbytes; i += 8) {
    // the 24-bit register meets the first three bytes, the other five start from zero
    const uint32_t b0 = ((crc >> 16) & 0xFFu) ^ data[i], b1 = ((crc >> 8) & 0xFFu) ^ data[i + 1],
                   b2 = (crc & 0xFFu) ^ data[i + 2];
    crc = T->t[7][b0] ^ T->t[6][b1] ^ T->t[5][b2] ^ T->t[4][data[i + 3]] ^ T->t[3][data[i + 4]] ^ T->t[2][data[i + 5]] ^
          T->t[1][data[i + 6]] ^ T->t[0][data[i + 7]];
  }
  for (; i < nbytes; i++) crc = ((crc << 8) ^ T->t[0][((crc >> 16) & 0xFFu) ^ data[i]]) & 0xFFFFFFu;
  return crc;
}

void crc_pos_tables(uint32_t K, std::vector<uint32_t>& out)
{
  const int      ki = cb_index_exact(K);
  const uint32_t W = (uint32_t)nof_windows(K), L = W ? K / W : 0;
  out.clear();
  if (ki < 0 || W == 0) return;
  const uint64_t f1 = kQpp[ki].f1, f2 = kQpp[ki].f2;
  out.resize(4 * (size_t)K);
  const uint32_t polys[2] = {kCrc24A, kCrc24B};
  std::vector<uint32_t> pw(K + 24);
  for (uint32_t which = 0; which < 2; which++) {
    uint32_t v = 1;
    for (uint32_t n = 0; n < K + 24; n++) {  // pw[n] = x^n mod P
      pw[n] = v;
      v <<= 1;
      if (v & 0x1000000u) v ^= polys[which];
    }
    for (uint32_t k = 0; k < L; k++)
      for (uint32_t d = 0; d < W; d++) {
        const uint64_t i = (uint64_t)d * L + k;
        const uint32_t pi = (uint32_t)((f1 * i + f2 * i * i) % K);
        out[((which * 2 + 0) * (size_t)L + k) * W + d] = pw[K + 23 - (uint32_t)i];
        out[((which * 2 + 1) * (size_t)L + k) * W + d] = pw[K + 23 - pi];
      }
  }
}

void gold_tables(uint32_t len, std::vector<uint32_t>& x1_packed, std::vector<uint32_t>& x2_mask)
{
  const uint32_t Nc = 1600, n_tot = Nc + len;
  std::vector<uint8_t>  x1(n_tot + 31, 0);
  std::vector<uint32_t> m2(n_tot + 31, 0);
  x1[0] = 1;
  for (uint32_t n = 0; n < 31; n++) m2[n] = 1u << n;
  for (uint32_t n = 0; n < n_tot; n++) {
    x1[n + 31] = (uint8_t)(x1[n + 3] ^ x1[n]);
    m2[n + 31] = m2[n + 3] ^ m2[n + 2] ^ m2[n + 1] ^ m2[n];
  }
  x1_packed.assign((len + 31) / 32, 0);
  x2_mask.resize(len);
  for (uint32_t n = 0; n < len; n++) {
    x1_packed[n >> 5] |= (uint32_t)x1[n + Nc] << (n & 31);
    x2_mask[n] = m2[n + Nc];
  }
}

}  // namespace b200
