/*
 * oracle/tdec_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see tdec_port.h).
 *
 * Scalar restatement of the reference's 16-bit turbo-decode receive tail.  All arrays
 * are kept in NATURAL trellis order; the reference's sub-block storage layout is a
 * pure relabelling (storage index W*k+d <-> trellis position d*L+k) that is undone
 * when an input in that layout is loaded.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py runs this file side by side
 * with the reference's own compiled objects (oracle/_ref/libsrslte_ref.so) and
 * tests/test_oracle_golden.py checks it against the reference's known-answer data
 * (crc_test.h CRC words, turbodecoder_test.h K=504 code word) and against fixtures
 * produced by the compiled reference (tests/golden/, script tests/golden/make_golden.py).
 */
#define _GNU_SOURCE
#include "tdec_port.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF (-10000) /* turbodecoder_win.h:58,89 ; turbodecoder_gen.c INF */
#define WARMUP 40        /* win_overlap_len, turbodecoder_win.h:54,87         */

/* ------------------------------------------------------------------------------------
 * 36.212 Table 5.1.3-3: K, f1, f2.  (cbsegm.c:30-44, tc_interl_lte.c:38-62)
 * ---------------------------------------------------------------------------------- */
static const struct { uint16_t K, f1, f2; } qpp_tab[PORT_NOF_K] = {
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

int port_cb_size(uint32_t idx) { return idx < PORT_NOF_K ? (int)qpp_tab[idx].K : -1; }

int port_cb_index(uint32_t long_cb)
{
  /* cbsegm.c:109-120: first table entry that is >= long_cb */
  for (int j = 0; j < PORT_NOF_K; j++)
    if (qpp_tab[j].K >= long_cb) return j;
  return -1;
}

int port_cb_is_valid(uint32_t K)
{
  int i = port_cb_index(K);
  return i >= 0 && qpp_tab[i].K == K;
}

int port_nof_subblocks(uint32_t K)
{
  /* turbodecoder.c:394-406, AVX2 build */
  if (K % 16 == 0 && K > 800) return 16;
  if (K % 8 == 0 && K > 400) return 8;
  return 0;
}

int port_qpp_params(uint32_t K, uint32_t* f1, uint32_t* f2)
{
  int i = port_cb_index(K);
  if (i < 0 || qpp_tab[i].K != K) return -1;
  *f1 = qpp_tab[i].f1;
  *f2 = qpp_tab[i].f2;
  return 0;
}

uint32_t port_qpp(uint32_t K, uint32_t i)
{
  uint32_t f1 = 0, f2 = 0;
  port_qpp_params(K, &f1, &f2);
  return (uint32_t)(((uint64_t)f1 * i + (uint64_t)f2 * i * i) % K);
}

/* ------------------------------------------------------------------------------------
 * CRC (MSB first, init 0, no reflection, no final xor)          crc.c:38-153
 * ---------------------------------------------------------------------------------- */
static uint32_t crc24_step_bit(uint32_t crc, uint32_t poly, unsigned bit)
{
  unsigned top = ((crc >> 23) & 1u) ^ (bit & 1u);
  crc = (crc << 1) & 0xFFFFFFu;
  if (top) crc ^= (poly & 0xFFFFFFu);
  return crc;
}

uint32_t port_crc_bytes(uint32_t poly, const uint8_t* data, uint32_t nbits)
{
  uint32_t crc = 0;
  for (uint32_t i = 0; i < nbits / 8; i++)
    for (int b = 7; b >= 0; b--) crc = crc24_step_bit(crc, poly, (data[i] >> b) & 1u);
  return crc;
}

uint32_t port_crc_bits(uint32_t poly, const uint8_t* bits, uint32_t nbits)
{
  /* crc.c:98-136 pads the last partial byte with zeros and then runs the register
   * backwards by the number of pad bits; for an init-0 MSB-first CRC that equals the
   * bit-serial CRC over exactly nbits bits.                                           */
  uint32_t crc = 0;
  for (uint32_t i = 0; i < nbits; i++) crc = crc24_step_bit(crc, poly, bits[i] & 1u);
  return crc;
}

/* ------------------------------------------------------------------------------------
 * Code-block segmentation                                         cbsegm.c:53-104
 * ---------------------------------------------------------------------------------- */
int port_cbsegm(port_cbsegm_t* s, uint32_t tbs)
{
  memset(s, 0, sizeof(*s));
  if (tbs == 0) return 0;
  uint32_t B = tbs + 24, Bp;
  s->tbs = tbs;
  if (B <= PORT_MAX_K) {
    s->C = 1;
    Bp   = B;
  } else {
    s->C = (uint32_t)ceilf((float)B / (float)(PORT_MAX_K - 24));
    Bp   = B + 24 * s->C;
  }
  int idx1 = port_cb_index((Bp - 1) / s->C + 1);
  if (idx1 < 0) return -1;
  s->K1     = (uint32_t)port_cb_size((uint32_t)idx1);
  s->K1_idx = (uint32_t)idx1;
  if (s->C == 1) {
    s->C1 = 1;
  } else {
    if (idx1 == 0) return -1;
    s->K2     = (uint32_t)port_cb_size((uint32_t)idx1 - 1);
    s->K2_idx = (uint32_t)idx1 - 1;
    s->C2     = (s->C * s->K1 - Bp) / (s->K1 - s->K2);
    s->C1     = s->C - s->C2;
  }
  s->F = s->C1 * s->K1 + s->C2 * s->K2 - Bp;
  return 0;
}

/* ------------------------------------------------------------------------------------
 * Rate de-matching                                   rm_turbo.c:160-260, 374-426
 * ---------------------------------------------------------------------------------- */
static const uint8_t col_perm[32] = {0, 16, 8,  24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                                     1, 17, 9,  25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};

/* natural output index (3*d + stream) held by circular-buffer position jp, -1 = dummy */
static int circ_to_natural(int jp, int R, int ND)
{
  int KP = 32 * R;
  int stream, d;
  if (jp < KP) {
    stream = 0;
    d      = (jp % R) * 32 + col_perm[jp / R] - ND;
  } else if (((jp - KP) & 1) == 0) {
    int q  = (jp - KP) / 2;
    stream = 1;
    d      = (q % R) * 32 + col_perm[q / R] - ND;
  } else {
    int q  = (jp - KP - 1) / 2;
    stream = 2;
    d      = (col_perm[q / R] + 32 * (q % R) + 1) % KP - ND;
  }
  return d < 0 ? -1 : 3 * d + stream;
}

static uint32_t natural_to_storage(uint32_t nat, uint32_t K, int W)
{
  /* rm_turbo.c:246-260: stream j at j*(K+32); position n at (n mod L)*W + n/L; tail kept */
  if (W <= 0) return nat;
  if (nat >= 3 * K) return (nat - 3 * K) + 3 * (K + 32);
  uint32_t L = K / (uint32_t)W, n = nat / 3, j = nat % 3;
  return j * (K + 32) + (n % L) * (uint32_t)W + n / L;
}

int port_rm_rx_table(uint32_t K, uint32_t rv, int sb_layout, uint16_t* table)
{
  if (!port_cb_is_valid(K) || rv > 3) return -2;
  int N   = 3 * (int)K + 12;
  int D   = (int)K + 4;
  int R   = (D - 1) / 32 + 1;
  int KP  = 32 * R;
  int ND  = KP - D;
  int Ncb = 3 * KP;
  int k0  = R * (2 * (int)ceilf((float)Ncb / (float)(8 * R)) * (int)rv + 2);
  int W   = sb_layout ? port_nof_subblocks(K) : 0;
  int k = 0, j = 0;
  while (k < N) {
    int nat = circ_to_natural((k0 + j) % Ncb, R, ND);
    if (nat >= 0) table[k++] = (uint16_t)natural_to_storage((uint32_t)nat, K, W);
    j++;
  }
  return 0;
}

int port_rm_turbo_rx(const int16_t* e, uint32_t E, int16_t* softbuf, uint32_t K, uint32_t rv,
                     int sb_layout)
{
  uint32_t  N   = 3 * K + 12;
  uint16_t* tab = (uint16_t*)malloc(sizeof(uint16_t) * N);
  if (!tab) return -1;
  int r = port_rm_rx_table(K, rv, sb_layout, tab);
  if (r == 0)
    for (uint32_t i = 0; i < E; i++) {
      uint16_t o = tab[i % N];
      softbuf[o] = (int16_t)(softbuf[o] + e[i]); /* wrapping +=, rm_turbo.c:417 */
    }
  free(tab);
  return r;
}

/* ------------------------------------------------------------------------------------
 * max-log-MAP
 * ---------------------------------------------------------------------------------- */
struct port_tdec {
  uint32_t K;
  int      W;
  int      n_iter;
  int      cb_ok;
  uint64_t clamps;
  int16_t  sys[PORT_MAX_K + 4], par0[PORT_MAX_K + 4], par1[PORT_MAX_K + 4];
  int16_t  app1[PORT_MAX_K + 4], app2[PORT_MAX_K + 4], ext1[PORT_MAX_K + 4], ext2[PORT_MAX_K + 4];
  int16_t  beta[8 * (PORT_MAX_K + 4)];
  uint32_t perm[PORT_MAX_K];
};

static inline int16_t sat_add(port_tdec_t* h, int a, int b)
{
  int v = a + b;
  if (v > 32767) { h->clamps++; return 32767; }
  if (v < -32768) { h->clamps++; return -32768; }
  return (int16_t)v;
}
static inline int16_t sat_sub(port_tdec_t* h, int a, int b) { return sat_add(h, a, -b); }
static inline int16_t wrap16(int v) { return (int16_t)(uint16_t)(unsigned)v; }
static inline int16_t max16(int16_t a, int16_t b) { return a > b ? a : b; }

/* sat != 0: saturating adds (window decoders); sat == 0: wrapping adds (generic, tail) */
#define ADD(a, b) (sat ? sat_add(h, (a), (b)) : wrap16((a) + (b)))

static inline void beta_step(port_tdec_t* h, int sat, int16_t b[8], int16_t x, int16_t y, int16_t xy)
{
  /* turbodecoder_win.h:488-511 / turbodecoder_gen.c:76-99 */
  int16_t m[8], n[8];
  m[0] = ADD(b[4], xy); m[1] = b[4];          m[2] = ADD(b[5], y);  m[3] = ADD(b[5], x);
  m[4] = ADD(b[6], x);  m[5] = ADD(b[6], y);  m[6] = b[7];          m[7] = ADD(b[7], xy);
  n[0] = b[0];          n[1] = ADD(b[0], xy); n[2] = ADD(b[1], x);  n[3] = ADD(b[1], y);
  n[4] = ADD(b[2], y);  n[5] = ADD(b[2], x);  n[6] = ADD(b[3], xy); n[7] = b[3];
  for (int i = 0; i < 8; i++) b[i] = max16(m[i], n[i]);
}

static inline void alpha_branches(port_tdec_t* h, int sat, const int16_t a[8], int16_t x, int16_t y,
                                  int16_t xy, int16_t m[8], int16_t n[8])
{
  /* turbodecoder_win.h:614-632 / turbodecoder_gen.c:142-158 */
  m[0] = a[0];          m[1] = ADD(a[3], y);  m[2] = ADD(a[4], y);  m[3] = a[7];
  m[4] = a[1];          m[5] = ADD(a[2], y);  m[6] = ADD(a[5], y);  m[7] = a[6];
  n[0] = ADD(a[1], xy); n[1] = ADD(a[2], x);  n[2] = ADD(a[5], x);  n[3] = ADD(a[6], xy);
  n[4] = ADD(a[0], xy); n[5] = ADD(a[3], x);  n[6] = ADD(a[4], x);  n[7] = ADD(a[7], xy);
}

static inline void win_normalize(port_tdec_t* h, int k, int16_t s[8])
{
  /* turbodecoder_win.h:332-349, normalize_period 2 */
  if ((k % 2) == 0 && k != 0) {
    for (int i = 1; i < 8; i++) s[i] = sat_sub(h, s[i], s[0]);
    s[0] = 0;
  }
}

static void map_window(port_tdec_t* h, int W, const int16_t* sys, const int16_t* app,
                       const int16_t* par, int16_t* out)
{
  const int K = (int)h->K, L = K / W, sat = 1;
  for (int d = 0; d < W; d++) {
    int16_t  b[8], a[8], m[8], n[8];
    int16_t* B = h->beta; /* B[8*k+i], k = 0..L */

    /* ---- beta boundary (turbodecoder_win.h:414-477, 351-395) ---- */
    if (d < W - 1) {
      for (int i = 0; i < 8; i++) b[i] = NEG_INF;
      for (int k = WARMUP - 1; k >= 0; k--) {
        int     p  = (d + 1) * L + k;
        int16_t x  = app ? sat_add(h, app[p], sys[p]) : sys[p];
        int16_t y  = par[p];
        int16_t xy = sat_add(h, x, y);
        beta_step(h, 1, b, x, y, xy);
        win_normalize(h, k, b);
      }
    } else {
      b[0] = 0;
      for (int i = 1; i < 8; i++) b[i] = NEG_INF;
      for (int k = K + 2; k >= K; k--) { /* tail: plain adds, no normalisation */
        int16_t x = sys[k], y = par[k], xy = wrap16(x + y);
        beta_step(h, 0, b, x, y, xy);
      }
    }
    memcpy(&B[8 * L], b, sizeof(b));

    /* ---- beta over the window (turbodecoder_win.h:479-526) ---- */
    for (int k = L - 1; k >= 0; k--) {
      int     p  = d * L + k;
      int16_t x  = app ? sat_add(h, app[p], sys[p]) : sys[p];
      int16_t y  = par[p];
      int16_t xy = sat_add(h, x, y);
      beta_step(h, 1, b, x, y, xy);
      memcpy(&B[8 * k], b, sizeof(b));
      win_normalize(h, k, b);
    }

    /* ---- alpha boundary (turbodecoder_win.h:552-603) ---- */
    if (d == 0) {
      a[0] = 0;
      for (int i = 1; i < 8; i++) a[i] = NEG_INF;
    } else {
      for (int i = 0; i < 8; i++) a[i] = NEG_INF;
      for (int k = 0; k < WARMUP; k++) {
        int     p  = (d - 1) * L + (L - WARMUP) + k;
        int16_t x  = app ? sat_add(h, app[p], sys[p]) : sys[p];
        int16_t y  = par[p];
        int16_t xy = sat_add(h, x, y);
        alpha_branches(h, 1, a, x, y, xy, m, n);
        for (int i = 0; i < 8; i++) a[i] = max16(m[i], n[i]);
        win_normalize(h, k, a);
      }
    }

    /* ---- alpha + output (turbodecoder_win.h:605-677) ---- */
    for (int k = 0; k < L; k++) {
      int     p  = d * L + k;
      int16_t x  = app ? sat_add(h, app[p], sys[p]) : sys[p];
      int16_t y  = par[p];
      int16_t xy = sat_add(h, x, y);
      alpha_branches(h, 1, a, x, y, xy, m, n);
      const int16_t* bk = &B[8 * (k + 1)];
      int16_t        M0 = sat_add(h, bk[0], m[0]), M1 = sat_add(h, bk[0], n[0]);
      for (int i = 1; i < 8; i++) {
        M0 = max16(M0, sat_add(h, bk[i], m[i]));
        M1 = max16(M1, sat_add(h, bk[i], n[i]));
      }
      int16_t o = sat_sub(h, M1, M0);
      if (W == 8) o = (int16_t)(o >> 1); /* divide_output, sse16 only: turbodecoder_win.h:56,658 */
      out[p] = o;
      for (int i = 0; i < 8; i++) a[i] = max16(m[i], n[i]);
      win_normalize(h, k, a);
    }
    (void)sat;
  }
}

static void map_generic(port_tdec_t* h, const int16_t* sys, const int16_t* app, const int16_t* par,
                        int16_t* out)
{
  /* turbodecoder_gen.c:54-231: no windows, wrapping arithmetic, normalise every 4 */
  const int K = (int)h->K, sat = 0;
  int16_t   b[8], a[8], m[8], n[8];
  int16_t*  B = h->beta;
  b[0]        = 0;
  for (int i = 1; i < 8; i++) b[i] = NEG_INF;
  memcpy(&B[8 * (K + 3)], b, sizeof(b));
  for (int k = K + 2; k >= 0; k--) {
    int16_t x = sys[k];
    if (app && k < K) x = wrap16(x + app[k]);
    int16_t y = par[k], xy = wrap16(x + y);
    beta_step(h, 0, b, x, y, xy);
    memcpy(&B[8 * k], b, sizeof(b));
    if ((k % 4) == 0 && k < K) {
      for (int i = 1; i < 8; i++) b[i] = wrap16(b[i] - b[0]);
      b[0] = 0;
    }
  }
  a[0] = 0;
  for (int i = 1; i < 8; i++) a[i] = NEG_INF;
  for (int k = 1; k <= K; k++) {
    int16_t x = sys[k - 1];
    if (app) x = wrap16(x + app[k - 1]);
    int16_t y = par[k - 1], xy = wrap16(x + y);
    alpha_branches(h, 0, a, x, y, xy, m, n);
    const int16_t* bk = &B[8 * k];
    int16_t        M0 = wrap16(m[0] + bk[0]), M1 = wrap16(n[0] + bk[0]);
    for (int i = 1; i < 8; i++) {
      M0 = max16(M0, wrap16(m[i] + bk[i]));
      M1 = max16(M1, wrap16(n[i] + bk[i]));
    }
    for (int i = 0; i < 8; i++) a[i] = max16(m[i], n[i]);
    if ((k % 4) == 0) {
      for (int i = 1; i < 8; i++) a[i] = wrap16(a[i] - a[0]);
      a[0] = 0;
    }
    out[k - 1] = wrap16(M1 - M0);
  }
  (void)sat;
}

static void map_dec(port_tdec_t* h, const int16_t* sys, const int16_t* app, const int16_t* par,
                    int16_t* out)
{
  if (h->W)
    map_window(h, h->W, sys, app, par, out);
  else
    map_generic(h, sys, app, par, out);
}

/* ------------------------------------------------------------------------------------
 * handle + half-iteration controller              turbodecoder_iter.h:68-142
 * ---------------------------------------------------------------------------------- */
port_tdec_t* port_tdec_new(void)
{
  port_tdec_t* h = (port_tdec_t*)calloc(1, sizeof(port_tdec_t));
  if (h) h->cb_ok = 0;
  return h;
}
void port_tdec_free(port_tdec_t* h) { free(h); }

int port_tdec_new_cb(port_tdec_t* h, uint32_t K)
{
  /* turbodecoder.c:522-537.  NB the reference accepts any long_cb <= max and rounds the
   * table index up; every caller passes a valid K, and so must callers of the port.  */
  if (!h || !port_cb_is_valid(K)) {
    if (h) h->cb_ok = 0;
    return -1;
  }
  if (K != h->K) {
    uint32_t f1, f2;
    port_qpp_params(K, &f1, &f2);
    for (uint32_t i = 0; i < K; i++)
      h->perm[i] = (uint32_t)(((uint64_t)f1 * i + (uint64_t)f2 * i * i) % K);
  }
  h->K      = K;
  h->W      = port_nof_subblocks(K);
  h->n_iter = 0;
  h->cb_ok  = 1;
  h->clamps = 0;
  return 0;
}

static void load_input(port_tdec_t* h, const int16_t* in, int natural)
{
  const uint32_t K = h->K;
  if (natural || h->W == 0) {
    /* turbodecoder_gen.c:233-250 / turbodecoder_win.h:727-769 (the generic decoder always
     * takes natural order: turbodecoder_iter.h:45 input_is_interleaved = current_dec > 0) */
    for (uint32_t i = 0; i < K; i++) {
      h->sys[i]  = in[3 * i];
      h->par0[i] = in[3 * i + 1];
      h->par1[i] = in[3 * i + 2];
    }
    for (uint32_t t = 0; t < 3; t++) {
      h->sys[K + t]  = in[3 * K + 2 * t];
      h->par0[K + t] = in[3 * K + 2 * t + 1];
      h->app2[K + t] = in[3 * K + 6 + 2 * t];
      h->par1[K + t] = in[3 * K + 6 + 2 * t + 1];
    }
  } else {
    /* sub-block layout: turbodecoder_iter.h:56-65,86-93 ; rm_turbo.c:246-260 */
    const uint32_t W = (uint32_t)h->W, L = K / W, S = K + 32;
    for (uint32_t n = 0; n < K; n++) {
      uint32_t s = (n % L) * W + n / L;
      h->sys[n]  = in[s];
      h->par0[n] = in[S + s];
      h->par1[n] = in[2 * S + s];
    }
    for (uint32_t t = 0; t < 3; t++) {
      h->sys[K + t]  = in[3 * S + 2 * t];
      h->par0[K + t] = in[3 * S + 2 * t + 1];
      h->app2[K + t] = in[3 * S + 6 + 2 * t];
      h->par1[K + t] = in[3 * S + 6 + 2 * t + 1];
    }
  }
}

static void decide(const port_tdec_t* h, uint8_t* out)
{
  /* turbodecoder.c:383-390: app1 when n_iter is even, ext1 when odd; bit = LLR > 0, MSB first */
  const int16_t* src = (h->n_iter % 2) == 0 ? h->app1 : h->ext1;
  for (uint32_t i = 0; i < h->K / 8; i++) {
    uint8_t v = 0;
    for (int b = 0; b < 8; b++) v = (uint8_t)((v << 1) | (src[8 * i + b] > 0));
    out[i] = v;
  }
}

static void half_iteration(port_tdec_t* h, const int16_t* input, int natural)
{
  const uint32_t K = h->K;
  const int      n = h->n_iter;
  if (n == 0) load_input(h, input, natural);
  if ((n % 2) == 0) {
    if (n)
      for (uint32_t i = 0; i < K; i++) h->app1[i] = wrap16(h->app1[i] - h->ext1[i]);
    map_dec(h, h->sys, n ? h->app1 : NULL, h->par0, h->ext1);
  } else {
    if (n > 1)
      for (uint32_t i = 0; i < K; i++) h->ext1[i] = wrap16(h->ext1[i] - h->app1[i]);
    for (uint32_t i = 0; i < K; i++) h->app2[i] = h->ext1[h->perm[i]];
    map_dec(h, h->app2, NULL, h->par1, h->ext2);
    for (uint32_t i = 0; i < K; i++) h->app1[h->perm[i]] = h->ext2[i];
  }
  h->n_iter++;
}

void port_tdec_iteration(port_tdec_t* h, const int16_t* input, int natural, uint8_t* out)
{
  if (!h || !h->cb_ok) return; /* turbodecoder.c:541 */
  half_iteration(h, input, natural);
  decide(h, out);
}

int port_tdec_run_all(port_tdec_t* h, const int16_t* input, int natural, uint8_t* out,
                      uint32_t nof_iterations, uint32_t K)
{
  if (port_tdec_new_cb(h, K)) return -1;
  do {
    half_iteration(h, input, natural);
  } while ((uint32_t)h->n_iter < nof_iterations); /* turbodecoder.c:555-557 */
  decide(h, out);
  return 0;
}

int port_tdec_get_nof_iterations(const port_tdec_t* h) { return h->n_iter; }
uint64_t port_tdec_clamp_count(const port_tdec_t* h) { return h->clamps; }
const int16_t* port_tdec_last_llr(const port_tdec_t* h)
{
  return (h->n_iter % 2) == 0 ? h->app1 : h->ext1;
}

/* ------------------------------------------------------------------------------------
 * soft buffer + transport block                 softbuffer.c:41-150, sch.c:299-500
 * ---------------------------------------------------------------------------------- */
int port_softbuffer_init(port_softbuffer_t* q, uint32_t max_cb)
{
  memset(q, 0, sizeof(*q));
  q->max_cb   = max_cb;
  q->buffer_f = (int16_t*)calloc((size_t)max_cb * PORT_SOFTBUF_LEN, sizeof(int16_t));
  q->data     = (uint8_t*)calloc((size_t)max_cb * (PORT_MAX_K / 8), 1);
  q->cb_crc   = (uint8_t*)calloc(max_cb, 1);
  if (!q->buffer_f || !q->data || !q->cb_crc) {
    port_softbuffer_free(q);
    return -1;
  }
  return 0;
}

void port_softbuffer_reset(port_softbuffer_t* q)
{
  memset(q->buffer_f, 0, (size_t)q->max_cb * PORT_SOFTBUF_LEN * sizeof(int16_t));
  memset(q->data, 0, (size_t)q->max_cb * (PORT_MAX_K / 8));
  memset(q->cb_crc, 0, q->max_cb);
  q->tb_crc = 0;
}

void port_softbuffer_free(port_softbuffer_t* q)
{
  free(q->buffer_f);
  free(q->data);
  free(q->cb_crc);
  memset(q, 0, sizeof(*q));
}

int port_decode_tb(port_tdec_t* dec, port_softbuffer_t* sb, uint32_t tbs, uint32_t Qm, uint32_t rv,
                   uint32_t nof_e_bits, const int16_t* e_bits, uint8_t* data,
                   uint32_t max_iterations, float* avg_iterations, uint32_t* cb_noi)
{
  if (!dec || !sb || !e_bits || !data) return -2;
  port_cbsegm_t seg;
  if (port_cbsegm(&seg, tbs)) return -1;
  if (seg.tbs == 0 || seg.C == 0) return 0;
  if (seg.F) return -2;
  if (seg.C > sb->max_cb) return -2;

  data[tbs / 8 + 0] = 0;
  data[tbs / 8 + 1] = 0;
  data[tbs / 8 + 2] = 0;

  float total_it = 0;
  for (uint32_t cb = 0; cb < seg.C; cb++) {
    uint32_t K    = cb < seg.C1 ? seg.K1 : seg.K2;
    uint32_t rlen = seg.C == 1 ? K : K - 24;
    if (cb_noi) cb_noi[cb] = 0;
    if (sb->cb_crc[cb]) {
      memcpy(&data[cb * rlen / 8], &sb->data[(size_t)cb * (PORT_MAX_K / 8)], rlen / 8);
      continue;
    }
    /* sch.c:324-334, incl. the `cb_idx > C - gamma` quirk */
    uint32_t Gp = nof_e_bits / Qm, gamma = Gp % seg.C, n_e = Qm * (Gp / seg.C);
    uint32_t rp = cb * n_e, n_e2 = n_e;
    if (cb > seg.C - gamma) {
      n_e2 = n_e + Qm;
      rp   = (seg.C - gamma) * n_e + (cb - (seg.C - gamma)) * n_e2;
    }
    int16_t* buf = &sb->buffer_f[(size_t)cb * PORT_SOFTBUF_LEN];
    if (port_rm_turbo_rx(&e_bits[rp], n_e2, buf, K, rv, 1)) return -1;

    port_tdec_new_cb(dec, K);
    uint32_t noi = 0;
    int      ok  = 0;
    do {
      port_tdec_iteration(dec, buf, 0, &data[cb * rlen / 8]);
      noi++;
      total_it += 1.0f;
      uint32_t crc = seg.C > 1 ? port_crc_bytes(PORT_CRC24B, &data[cb * rlen / 8], K)
                               : port_crc_bytes(PORT_CRC24A, &data[cb * rlen / 8], tbs + 24);
      if (crc == 0) {
        sb->cb_crc[cb] = 1;
        ok             = 1;
      }
    } while (noi < max_iterations && !ok);
    if (cb_noi) cb_noi[cb] = noi;
  }

  sb->tb_crc = 1;
  for (uint32_t i = 0; i < seg.C && sb->tb_crc; i++) sb->tb_crc = sb->cb_crc[i];
  if (!sb->tb_crc) {
    for (uint32_t i = 0; i < seg.C; i++)
      if (sb->cb_crc[i]) {
        uint32_t K    = i < seg.C1 ? seg.K1 : seg.K2;
        uint32_t rlen = seg.C == 1 ? K : K - 24;
        memcpy(&sb->data[(size_t)i * (PORT_MAX_K / 8)], &data[i * rlen / 8], rlen / 8);
      }
  }
  if (avg_iterations) *avg_iterations = total_it / (float)seg.C;
  if (!sb->tb_crc) return -1;

  uint32_t par_rx = port_crc_bytes(PORT_CRC24A, data, tbs);
  uint32_t par_tx = ((uint32_t)data[tbs / 8] << 16) | ((uint32_t)data[tbs / 8 + 1] << 8) |
                    (uint32_t)data[tbs / 8 + 2];
  return (par_rx == par_tx && par_rx) ? 0 : -1; /* sch.c:481 */
}

/* ------------------------------------------------------------------------------------
 * batch helper (CPU baseline)
 * ---------------------------------------------------------------------------------- */
typedef struct {
  const int16_t* in;
  uint8_t*       out;
  uint32_t       in_stride, out_stride, first, last, K, nit;
  int            natural, rc;
} batch_job_t;

static void* batch_worker(void* arg)
{
  batch_job_t* j = (batch_job_t*)arg;
  port_tdec_t* h = port_tdec_new();
  if (!h) {
    j->rc = -1;
    return NULL;
  }
  for (uint32_t i = j->first; i < j->last; i++)
    if (port_tdec_run_all(h, j->in + (size_t)i * j->in_stride, j->natural,
                          j->out + (size_t)i * j->out_stride, j->nit, j->K))
      j->rc = -1;
  port_tdec_free(h);
  return NULL;
}

int port_batch_run_all(const int16_t* in, uint32_t in_stride, int natural, uint8_t* out,
                       uint32_t out_stride, uint32_t n, uint32_t K, uint32_t nof_iterations,
                       uint32_t threads)
{
  if (threads == 0) threads = 1;
  if (threads > n) threads = n ? n : 1;
  pthread_t*   th   = (pthread_t*)calloc(threads, sizeof(pthread_t));
  batch_job_t* jobs = (batch_job_t*)calloc(threads, sizeof(batch_job_t));
  int          rc   = 0;
  for (uint32_t t = 0; t < threads; t++) {
    jobs[t] = (batch_job_t){in, out, in_stride, out_stride, (uint32_t)((uint64_t)n * t / threads),
                            (uint32_t)((uint64_t)n * (t + 1) / threads), K, nof_iterations, natural, 0};
    pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  for (uint32_t t = 0; t < threads; t++) {
    pthread_join(th[t], NULL);
    if (jobs[t].rc) rc = -1;
  }
  free(th);
  free(jobs);
  return rc;
}

/* =====================================================================================================
 * Front end of the path: soft demodulation to int16 LLRs + descrambling.
 * TEST INFRASTRUCTURE like the rest of this file.
 *
 * Reference (AVX2/SSE build, as compiled by oracle/Makefile):
 *   srslte_demod_soft_demodulate_s      lib/src/phy/modem/demod_soft.c:503-525
 *     QPSK    demod_qpsk_lte_s          :68-70  -> srslte_vec_convert_fi_simd, lib/src/phy/utils/vector_simd.c:392-427
 *     16QAM   demod_16qam_lte_s(_sse)   :90-133 (groups of 4 symbols), scalar remainder :121-131
 *     64QAM   demod_64qam_lte_s(_sse)   :240-302, scalar remainder :290-300
 *     256QAM  demod_256qam_lte_s        :457-477 (scalar float)
 *   srslte_sequence_set_LTE_pr          lib/src/phy/common/sequence.c:46-75 (36.211 7.2 Gold sequence, Nc = 1600)
 *   srslte_scrambling_s_offset          lib/src/phy/scrambling/scrambling.c:44-47 -> srslte_vec_neg_sss (sign flip)
 * The SSE bodies of 16QAM / 64QAM round to nearest even (cvtps2dq) and saturate (packssdw); QPSK's AVX2 body
 * truncates (cvttps2dq) and saturates; the scalar remainders truncate and wrap (out-of-range conversions are
 * undefined behaviour in the reference: parity is claimed for |scale * x| < 32768 only).
 * ===================================================================================================== */
#include <math.h>

static int16_t sat16_i32(int32_t v) { return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); }
static int16_t rint_sat16(float v)
{ /* cvtps2dq + packssdw */
  if (!(v > -2147483648.0f && v < 2147483648.0f)) return (int16_t)-32768; /* integer indefinite */
  return sat16_i32((int32_t)lrintf(v));
}
static int16_t trunc16_f(float v) { return (int16_t)(int32_t)v; }
static int16_t trunc16_d(double v) { return (int16_t)(int32_t)v; }

void port_gold_sequence(uint32_t seed, uint32_t len, uint8_t* c)
{
  const uint32_t n_tot = 1600 + len;
  uint8_t* x1 = (uint8_t*)calloc(n_tot + 31, 1);
  uint8_t* x2 = (uint8_t*)calloc(n_tot + 31, 1);
  for (uint32_t n = 0; n < 31; n++) x2[n] = (uint8_t)((seed >> n) & 1u);
  x1[0] = 1;
  for (uint32_t n = 0; n < n_tot; n++) {
    x1[n + 31] = (uint8_t)((x1[n + 3] + x1[n]) & 1);
    x2[n + 31] = (uint8_t)((x2[n + 3] + x2[n + 2] + x2[n + 1] + x2[n]) & 1);
  }
  for (uint32_t n = 0; n < len; n++) c[n] = (uint8_t)((x1[n + 1600] + x2[n + 1600]) & 1);
  free(x1);
  free(x2);
}

/* qm = bits per symbol: 2 (QPSK), 4 (16QAM), 6 (64QAM), 8 (256QAM); sym = nsym complex floats (re, im) */
int port_demod_s(int qm, const float* sym, int16_t* llr, uint32_t nsym)
{
  if (qm == 2) {
    const float    scale = (float)(-100 * sqrt(2));
    const uint32_t len = 2 * nsym, simd = len & ~15u; /* 16 int16 per AVX2 iteration */
    for (uint32_t i = 0; i < simd; i++) { /* cvttps2dq + packssdw: truncate, then saturate (simd.h:1681-1686) */
      const float v = sym[i] * scale;
      llr[i] = (v > -2147483648.0f && v < 2147483648.0f) ? sat16_i32((int32_t)v) : (int16_t)-32768;
    }
    for (uint32_t i = simd; i < len; i++) llr[i] = trunc16_f(sym[i] * scale);
    return 0;
  }
  if (qm == 4) {
    const int16_t  off  = (int16_t)(2 * 400 / sqrt(10));
    const uint32_t simd = nsym & ~3u;
    for (uint32_t i = 0; i < simd; i++) {
      const int16_t re = rint_sat16(sym[2 * i] * -400.0f), im = rint_sat16(sym[2 * i + 1] * -400.0f);
      const int16_t are = (int16_t)(re < 0 ? -re : re), aim = (int16_t)(im < 0 ? -im : im); /* pabsw: -32768 stays */
      llr[4 * i + 0] = re;
      llr[4 * i + 1] = im;
      llr[4 * i + 2] = (int16_t)(are - off);
      llr[4 * i + 3] = (int16_t)(aim - off);
    }
    for (uint32_t i = simd; i < nsym; i++) {
      const int16_t yre = trunc16_f(400 * sym[2 * i]), yim = trunc16_f(400 * sym[2 * i + 1]);
      llr[4 * i + 0] = (int16_t)-yre;
      llr[4 * i + 1] = (int16_t)-yim;
      llr[4 * i + 2] = trunc16_d(abs(yre) - 2 * 400 / sqrt(10));
      llr[4 * i + 3] = trunc16_d(abs(yim) - 2 * 400 / sqrt(10));
    }
    return 0;
  }
  if (qm == 6) {
    const int16_t  off1 = (int16_t)(4 * 700 / sqrt(42)), off2 = (int16_t)(2 * 700 / sqrt(42));
    const uint32_t simd = nsym & ~3u;
    for (uint32_t i = 0; i < simd; i++) {
      for (int c = 0; c < 2; c++) {
        const int16_t v  = rint_sat16(sym[2 * i + c] * -700.0f);
        const int16_t a1 = (int16_t)((int16_t)(v < 0 ? -v : v) - off1);
        const int16_t a2 = (int16_t)((int16_t)(a1 < 0 ? -a1 : a1) - off2);
        llr[6 * i + c]     = v;
        llr[6 * i + 2 + c] = a1;
        llr[6 * i + 4 + c] = a2;
      }
    }
    for (uint32_t i = simd; i < nsym; i++) {
      for (int c = 0; c < 2; c++) {
        const float y = (float)trunc16_f(700 * sym[2 * i + c]);
        llr[6 * i + c]     = trunc16_f(-y);
        llr[6 * i + 2 + c] = trunc16_d(abs((int)y) - 4 * 700 / sqrt(42));
        llr[6 * i + 4 + c] = trunc16_d(abs(llr[6 * i + 2 + c]) - 2 * 700 / sqrt(42));
      }
    }
    return 0;
  }
  if (qm == 8) {
    const float c8 = 8.0f / sqrtf(170.0f), c4 = 4.0f / sqrtf(170.0f), c2 = 2.0f / sqrtf(170.0f);
    for (uint32_t i = 0; i < nsym; i++) {
      for (int c = 0; c < 2; c++) {
        float v = -sym[2 * i + c];
        llr[8 * i + c] = trunc16_f(1000 * v);
        v = fabsf(v) - c8;
        llr[8 * i + 2 + c] = trunc16_f(1000 * v);
        v = fabsf(v) - c4;
        llr[8 * i + 4 + c] = trunc16_f(1000 * v);
        v = fabsf(v) - c2;
        llr[8 * i + 6 + c] = trunc16_f(1000 * v);
      }
    }
    return 0;
  }
  return -1;
}

/* llr[i] = c[i] ? -llr[i] : llr[i] (wrapping negation: -(-32768) = -32768) */
void port_descramble_s(int16_t* llr, const uint8_t* c, uint32_t len)
{
  for (uint32_t i = 0; i < len; i++)
    if (c[i]) llr[i] = (int16_t)(0u - (uint16_t)llr[i]);
}

/* pdsch.c:760-779 / pusch.c:482-500: demodulate nsym symbols, descramble the first nof_bits LLRs with seed c_init */
int port_demod_descramble(int qm, const float* sym, uint32_t nsym, uint32_t c_init, uint32_t nof_bits, int16_t* llr)
{
  if (port_demod_s(qm, sym, llr, nsym)) return -1;
  if (nof_bits > (uint32_t)qm * nsym) return -1;
  uint8_t* c = (uint8_t*)malloc(nof_bits + 1);
  port_gold_sequence(c_init, nof_bits, c);
  port_descramble_s(llr, c, nof_bits);
  free(c);
  return 0;
}

/* ---- UL-SCH with multiplexed UCI (36.212 5.2.2.6 - 5.2.2.8): what srslte_ulsch_decode does to the LLRs before
 * decode_tb sees them (lib/src/phy/phch/sch.c:920-1064).  Only the DATA path is restated: the values of the ACK /
 * RI / CQI bits themselves are control information and stay with the reference's uci.c. ---- */

/* number of coded ACK or RI symbols: Q_prime_ri_ack, lib/src/phy/phch/uci.c:547-571, with K = cfg->K_segm > 0 */
uint32_t port_uci_q_prime_ri_ack(uint32_t O, uint32_t K_segm, uint32_t L_prb, uint32_t nof_symb, float beta)
{
  const uint32_t x = (uint32_t)ceilf((float)O * L_prb * 12 * nof_symb * beta / K_segm);
  const uint32_t m = 4 * L_prb * 12;
  return x < m ? x : m;
}

/* number of coded CQI symbols: Q_prime_cqi, lib/src/phy/phch/uci.c:266-283 (L = 8 CRC bits from 11 payload bits on) */
uint32_t port_uci_q_prime_cqi(uint32_t O, uint32_t K_segm, uint32_t L_prb, uint32_t nof_symb, float beta, uint32_t q_prime_ri)
{
  const uint32_t L = O < 11 ? 0 : 8;
  uint32_t       x = 999999;
  if (K_segm > 0) x = (uint32_t)ceilf((float)(O + L) * L_prb * 12 * nof_symb * beta / K_segm);
  const uint32_t m = L_prb * 12 * nof_symb - q_prime_ri;
  return x < m ? x : m;
}

/* position in q of bit k of coded ACK / RI symbol number idx: uci_ulsch_interleave_ack_gen / _ri_gen,
 * lib/src/phy/phch/uci.c:497-545 (bottom rows of the interleaver matrix, columns next to the reference symbols).
 * Returns -1 where the reference reports an error (more symbols than 4 per matrix row allow). */
static int port_uci_position(int is_ri, uint32_t idx, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs,
                             uint32_t k, uint32_t* pos)
{
  static const uint32_t ack_norm[4] = {2, 3, 8, 9}, ack_ext[4] = {1, 2, 6, 7};
  static const uint32_t ri_norm[4] = {1, 4, 7, 10}, ri_ext[4] = {0, 3, 5, 8};
  const uint32_t rows = H_prime_total / N_pusch_symbs;
  if (rows < 1 + idx / 4) return -1;
  const uint32_t row = rows - 1 - idx / 4, colidx = (3 * idx) % 4;
  const uint32_t col = N_pusch_symbs > 10 ? (is_ri ? ri_norm : ack_norm)[colidx] : (is_ri ? ri_ext : ack_ext)[colidx];
  *pos = row * Qm + rows * col * Qm + k;
  return 0;
}

/* q (descrambled LLRs, channel order, H' * Qm of them; NOT modified here) -> g (UL-SCH order), the sequence of
 * lib/src/phy/phch/sch.c:940-1035:
 *   1. the Q'_ack * Qm ACK positions are erased (q = 0, :961-964);
 *   2. a 1-bit RI decode flips q[p1] of every RI symbol back where c[p1] = 1 (decode_ri_ack_1bit, uci.c:622-634);
 *      the 2-bit decode does not write (uci.c:636-649);
 *   3. ulsch_deinterleave (:891-918): lut[x] = running index over the non-RI positions in row-major order, 0 for the
 *      RI positions (ulsch_interleave_gen :580-598), then srslte_vec_lut_sis g[lut[x]] = q[x] for x ascending -- so
 *      every RI sample is written to g[0] and the last writer in x order wins.
 * g has (H' - Q'_ri) * Qm defined entries; the rest is left untouched (stale in the reference).
 * Returns 0, or -1 where the reference fails.                                                                  */
int port_ulsch_demux(const int16_t* q_in, const uint8_t* c_seq, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs,
                     uint32_t q_prime_ack, uint32_t q_prime_ri, uint32_t ri_len, int16_t* g)
{
  const uint32_t n = H_prime_total * Qm, rows = H_prime_total / N_pusch_symbs, cols = N_pusch_symbs;
  int16_t*       q = (int16_t*)malloc(sizeof(int16_t) * (n + 1));
  uint8_t*       ri_present = (uint8_t*)calloc(n + 1, 1);
  uint32_t*      lut = (uint32_t*)calloc(n + 1, sizeof(uint32_t));
  int            rc = 0;
  memcpy(q, q_in, sizeof(int16_t) * n);
  for (uint32_t i = 0; i < q_prime_ack && !rc; i++)
    for (uint32_t k = 0; k < Qm; k++) {
      uint32_t pos;
      if (port_uci_position(0, i, Qm, H_prime_total, N_pusch_symbs, k, &pos)) { rc = -1; break; }
      q[pos] = 0;
    }
  for (uint32_t i = 0; i < q_prime_ri && !rc; i++)
    for (uint32_t k = 0; k < Qm; k++) {
      uint32_t pos;
      if (port_uci_position(1, i, Qm, H_prime_total, N_pusch_symbs, k, &pos)) { rc = -1; break; }
      ri_present[pos] = 1;
      if (ri_len == 1 && k == 1 && c_seq[pos]) q[pos] = (int16_t)(-q[pos]);
    }
  if (!rc) {
    uint32_t idx = 0;
    for (uint32_t j = 0; j < rows; j++)
      for (uint32_t i = 0; i < cols; i++)
        for (uint32_t k = 0; k < Qm; k++) {
          const uint32_t x = j * Qm + i * rows * Qm + k;
          lut[x] = ri_present[x] ? 0 : idx++;
        }
    for (uint32_t x = 0; x < n; x++) g[lut[x]] = q[x];
  }
  free(q);
  free(ri_present);
  free(lut);
  return rc;
}

/* UL-SCH channel de-interleaver of 36.212 5.2.2.8 without multiplexed UCI (no RI / ACK / CQI):
 * reference ulsch_deinterleave + ulsch_interleave_gen, lib/src/phy/phch/sch.c:580-598, 891-918:
 * lut[(i*rows + j)*Qm + k] = (j*cols + i)*Qm + k,  g[lut[x]] = q[x].                                         */
void port_ulsch_deinterleave(const int16_t* q, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs, int16_t* g)
{
  const uint32_t rows = H_prime_total / N_pusch_symbs, cols = N_pusch_symbs;
  uint32_t idx = 0;
  for (uint32_t j = 0; j < rows; j++)
    for (uint32_t i = 0; i < cols; i++)
      for (uint32_t k = 0; k < Qm; k++) g[idx++] = q[j * Qm + i * rows * Qm + k];
}
