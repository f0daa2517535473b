// tdec_kernels.cu -- sm_100a kernels of the LTE turbo-decode receive tail.
//
// What is computed (bit-exact with the reference's 16-bit AUTO/AVX2 decoder):
//   window max-log-MAP, W = 16 (K > 800) / W = 8 (400 < K <= 800)
//       reference: lib/include/srslte/phy/fec/turbodecoder_win.h:332-679
//   generic max-log-MAP (K <= 400)      reference: lib/src/phy/fec/turbodecoder_gen.c:54-231
//   half-iteration controller           reference: lib/include/srslte/phy/fec/turbodecoder_iter.h:68-142
//   hard decision                       reference: lib/src/phy/fec/turbodecoder.c:383-390
//   CRC24A/B with early termination     reference: lib/src/phy/phch/sch.c:353-383, lib/src/phy/fec/crc.c:139-153
//   rate de-matching                    reference: lib/src/phy/fec/rm_turbo.c:374-426
//
// How (B200 mapping):
//   * a code block is W independent trellis windows; one thread owns TWO adjacent windows as the
//     two int16 halves of every 32-bit register (the reference's storage layout puts windows
//     d, d+1 of trellis row k next to each other, so one 32-bit load feeds both).  W/2 threads per
//     block, 4 (W=16) or 8 (W=8) code blocks per warp, the 8 state metrics live in registers.
//   * the reference keeps all beta metrics of a half iteration (98 KB per K=6144 block).  Here the
//     backward pass keeps one checkpoint per 16 trellis rows; the forward pass rebuilds the 16 rows
//     of beta it needs in shared memory (the recursion and its normalisation points depend only
//     on the row index, so the rebuilt values are identical).
//   * the two extrinsic arrays are stored already differenced (what the reference computes with
//     srslte_vec_sub_sss at the start of the next half iteration), so the a-posteriori value of
//     every bit is always A + E (wrapping) and no third array is needed.
//   * QPP addresses are computed on the fly from (f1, f2): pi(d*L + k) shares its row
//     pi(k) mod L across all windows (contention-free property), only the window index differs.
#include "tdec_kernels.h"

#include <cstdio>

namespace b200 {

namespace {

constexpr int      kWarm    = 40;  // win_overlap_len
constexpr int      kChunk   = 16;  // rows of beta rebuilt at a time (must be even)
constexpr int      kThreads = 128;
constexpr int      kMaxChunks = 24;  // ceil(384 / 16)
constexpr int      kNegInf  = -10000;
constexpr uint32_t kNegInf2 = 0xD8F0D8F0u;  // (-10000, -10000)

__constant__ uint32_t c_crc_tab[2][256];

// ---- packed int16x2 arithmetic -------------------------------------------------------------------
__device__ __forceinline__ uint32_t sadd2(uint32_t a, uint32_t b) { return __vaddss2(a, b); }
__device__ __forceinline__ uint32_t ssub2(uint32_t a, uint32_t b) { return __vsubss2(a, b); }
__device__ __forceinline__ uint32_t wsub2(uint32_t a, uint32_t b) { return __vsub2(a, b); }
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
__device__ __forceinline__ uint32_t sra1_2(uint32_t v) { return ((v >> 1) & 0x7FFF7FFFu) | (v & 0x80008000u); }

__device__ __forceinline__ void beta_step(uint32_t b[8], uint32_t x, uint32_t y, uint32_t xy)
{
  const uint32_t n0 = max2(sadd2(b[4], xy), b[0]);
  const uint32_t n1 = max2(b[4], sadd2(b[0], xy));
  const uint32_t n2 = max2(sadd2(b[5], y), sadd2(b[1], x));
  const uint32_t n3 = max2(sadd2(b[5], x), sadd2(b[1], y));
  const uint32_t n4 = max2(sadd2(b[6], x), sadd2(b[2], y));
  const uint32_t n5 = max2(sadd2(b[6], y), sadd2(b[2], x));
  const uint32_t n6 = max2(b[7], sadd2(b[3], xy));
  const uint32_t n7 = max2(sadd2(b[7], xy), b[3]);
  b[0] = n0; b[1] = n1; b[2] = n2; b[3] = n3; b[4] = n4; b[5] = n5; b[6] = n6; b[7] = n7;
}

// branch metrics into each state: m = bit-0 branches, n = bit-1 branches
__device__ __forceinline__ void alpha_branches(const uint32_t a[8], uint32_t x, uint32_t y, uint32_t xy,
                                               uint32_t m[8], uint32_t n[8])
{
  m[0] = a[0];            m[1] = sadd2(a[3], y);  m[2] = sadd2(a[4], y);  m[3] = a[7];
  m[4] = a[1];            m[5] = sadd2(a[2], y);  m[6] = sadd2(a[5], y);  m[7] = a[6];
  n[0] = sadd2(a[1], xy); n[1] = sadd2(a[2], x);  n[2] = sadd2(a[5], x);  n[3] = sadd2(a[6], xy);
  n[4] = sadd2(a[0], xy); n[5] = sadd2(a[3], x);  n[6] = sadd2(a[4], x);  n[7] = sadd2(a[7], xy);
}

__device__ __forceinline__ void normalize(uint32_t s[8])
{
#pragma unroll
  for (int i = 1; i < 8; i++) s[i] = ssub2(s[i], s[0]);
  s[0] = 0;
}

// ---- per-warp decode state -----------------------------------------------------------------------
template <int W>
struct WinCtx {
  static constexpr int WH = W / 2;
  // geometry of this code-block size
  uint32_t K, L, f1, f2, mK, mL;
  // thread position
  int t;  // 0..WH-1: owns windows 2t (low half) and 2t+1 (high half)
  // QPP window offsets of the two windows (mod W)
  uint32_t base_lo, base_hi, inc_lo, inc_hi;
  // data
  const uint32_t* sys;   // [L][WH]
  const uint32_t* par0;  // [L][WH]
  const uint32_t* par1;  // [L][WH]
  const int16_t*  tail;  // 12 tail samples
  uint32_t*       A32;   // [L][WH]  extrinsic of DEC2 minus E (a-priori of DEC1)
  uint32_t*       E32;   // [L][WH]  a-posteriori of DEC1 minus A
  uint32_t*       chk;   // checkpoints: [(c*8 + i)*32], already offset by lane
  uint4*          sm;    // chunk of beta: [(s*2 + h)*kThreads], already offset by tid
};

template <int W>
__device__ __forceinline__ void qpp_at(const WinCtx<W>& c, uint32_t k, uint32_t& row, uint32_t& w_lo, uint32_t& w_hi)
{
  const uint32_t v = (c.f1 + c.f2 * k) * k;  // pi(k) before reduction; k < L <= 384 keeps it below 2^32
  uint32_t       p = v - __umulhi(v, c.mK) * c.K;
  if (p >= c.K) p -= c.K;
  if (p >= c.K) p -= c.K;
  const uint32_t w0 = __umulhi(p, c.mL);
  row               = p - w0 * c.L;
  w_lo              = (w0 + c.base_lo + c.inc_lo * k) & (W - 1);
  w_hi              = (w0 + c.base_hi + c.inc_hi * k) & (W - 1);
}

// systematic (+ a-priori) and parity of trellis row k for this thread's two windows.
// aux returns what the output stage has to subtract: the a-priori (DEC1) or x itself (DEC2).
template <int W>
__device__ __forceinline__ void load_xy(const WinCtx<W>& c, bool dec2, bool apriori, uint32_t k, uint32_t& x,
                                        uint32_t& y, uint32_t& aux)
{
  constexpr int WH = W / 2;
  if (!dec2) {
    x   = __ldg(c.sys + k * WH + c.t);
    y   = __ldg(c.par0 + k * WH + c.t);
    aux = 0;
    if (apriori) {
      aux = c.A32[k * WH + c.t];
      x   = sadd2(aux, x);
    }
  } else {
    uint32_t row, w_lo, w_hi;
    qpp_at<W>(c, k, row, w_lo, w_hi);
    const uint16_t* E16 = reinterpret_cast<const uint16_t*>(c.E32);
    x   = (uint32_t)E16[row * W + w_lo] | ((uint32_t)E16[row * W + w_hi] << 16);
    y   = __ldg(c.par1 + k * WH + c.t);
    aux = x;
  }
}

template <int W>
__device__ __forceinline__ void store_out(const WinCtx<W>& c, bool dec2, uint32_t k, uint32_t o, uint32_t aux)
{
  constexpr int  WH = W / 2;
  const uint32_t d  = wsub2(o, aux);
  if (!dec2) {
    c.E32[k * WH + c.t] = d;
  } else {
    uint32_t row, w_lo, w_hi;
    qpp_at<W>(c, k, row, w_lo, w_hi);
    uint16_t* A16       = reinterpret_cast<uint16_t*>(c.A32);
    A16[row * W + w_lo] = (uint16_t)(d & 0xFFFFu);
    A16[row * W + w_hi] = (uint16_t)(d >> 16);
  }
}

// beta of the terminated last window from the 3 tail rows: plain (wrapping) int16 arithmetic.
__device__ __forceinline__ void tail_beta(const int16_t* tl, int16_t b[8])
{
  b[0] = 0;
#pragma unroll
  for (int i = 1; i < 8; i++) b[i] = (int16_t)kNegInf;
  for (int r = 2; r >= 0; r--) {
    const int x = tl[2 * r], y = tl[2 * r + 1];
    const int xy = (int16_t)(x + y);
    int16_t   m[8], n[8];
    m[0] = (int16_t)(b[4] + xy); m[1] = b[4];               m[2] = (int16_t)(b[5] + y);  m[3] = (int16_t)(b[5] + x);
    m[4] = (int16_t)(b[6] + x);  m[5] = (int16_t)(b[6] + y);  m[6] = b[7];               m[7] = (int16_t)(b[7] + xy);
    n[0] = b[0];               n[1] = (int16_t)(b[0] + xy); n[2] = (int16_t)(b[1] + x);  n[3] = (int16_t)(b[1] + y);
    n[4] = (int16_t)(b[2] + y);  n[5] = (int16_t)(b[2] + x);  n[6] = (int16_t)(b[3] + xy); n[7] = b[3];
#pragma unroll
    for (int i = 0; i < 8; i++) b[i] = m[i] > n[i] ? m[i] : n[i];
  }
}

// One constituent MAP decoder run (one srsLTE "iteration") for this thread's two windows.
template <int W>
__device__ void half_iteration(const WinCtx<W>& c, bool dec2, bool apriori)
{
  constexpr int  WH = W / 2;
  const uint32_t L  = c.L;
  const int      nchunks = (int)((L + kChunk - 1) / kChunk);
  uint32_t       s[8], x, y, aux;

  // ---------------- backward pass: boundary metrics, then checkpoints ----------------
#pragma unroll
  for (int i = 0; i < 8; i++) s[i] = kNegInf2;
  for (int k = kWarm - 1; k >= 0; k--) {
    load_xy<W>(c, dec2, apriori, (uint32_t)k, x, y, aux);
    beta_step(s, x, y, sadd2(x, y));
    if ((k & 1) == 0 && k != 0) normalize(s);
  }
  {
    // window d starts from what window d+1 estimated; the last window from the tail
    int16_t tb[8];
    tail_beta(c.tail + (dec2 ? 6 : 0), tb);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint32_t nxt = __shfl_down_sync(0xFFFFFFFFu, s[i], 1, WH);
      uint32_t       v   = __byte_perm(s[i], nxt, 0x5432);
      if (c.t == WH - 1) v = (v & 0xFFFFu) | ((uint32_t)(uint16_t)tb[i] << 16);
      s[i] = v;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) c.chk[((nchunks - 1) * 8 + i) * 32] = s[i];
  for (int k = (int)L - 1; k >= 0; k--) {
    load_xy<W>(c, dec2, apriori, (uint32_t)k, x, y, aux);
    beta_step(s, x, y, sadd2(x, y));
    if ((k % kChunk) == 0 && k != 0) {
#pragma unroll
      for (int i = 0; i < 8; i++) c.chk[((k / kChunk - 1) * 8 + i) * 32] = s[i];
    }
    if ((k & 1) == 0 && k != 0) normalize(s);
  }

  // ---------------- forward pass ----------------
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = kNegInf2;
  for (int k = 0; k < kWarm; k++) {
    uint32_t m[8], n[8];
    load_xy<W>(c, dec2, apriori, L - kWarm + (uint32_t)k, x, y, aux);
    alpha_branches(a, x, y, sadd2(x, y), m, n);
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = max2(m[i], n[i]);
    if ((k & 1) == 0 && k != 0) normalize(a);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t prv = __shfl_up_sync(0xFFFFFFFFu, a[i], 1, WH);
    uint32_t       v   = __byte_perm(a[i], prv, 0x1076);
    if (c.t == 0) v = (v & 0xFFFF0000u) | (i == 0 ? 0u : (uint32_t)(uint16_t)kNegInf);
    a[i] = v;
  }

  for (int ch = 0; ch < nchunks; ch++) {
    const int lo = ch * kChunk;
    const int hi = min(lo + kChunk, (int)L);
    // rebuild B[lo+1 .. hi] into shared memory, slot (k - lo - 1) holds B[k]
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = c.chk[(ch * 8 + i) * 32];
    c.sm[((hi - lo - 1) * 2 + 0) * kThreads] = make_uint4(s[0], s[1], s[2], s[3]);
    c.sm[((hi - lo - 1) * 2 + 1) * kThreads] = make_uint4(s[4], s[5], s[6], s[7]);
    if (hi != (int)L) normalize(s);  // hi is even and non-zero: the backward pass normalised after storing
    for (int k = hi - 1; k > lo; k--) {
      load_xy<W>(c, dec2, apriori, (uint32_t)k, x, y, aux);
      beta_step(s, x, y, sadd2(x, y));
      c.sm[((k - lo - 1) * 2 + 0) * kThreads] = make_uint4(s[0], s[1], s[2], s[3]);
      c.sm[((k - lo - 1) * 2 + 1) * kThreads] = make_uint4(s[4], s[5], s[6], s[7]);
      if ((k & 1) == 0) normalize(s);
    }
    // alpha recursion + a-posteriori output over the chunk
    for (int k = lo; k < hi; k++) {
      uint32_t m[8], n[8];
      load_xy<W>(c, dec2, apriori, (uint32_t)k, x, y, aux);
      alpha_branches(a, x, y, sadd2(x, y), m, n);
      const uint4 b0 = c.sm[((k - lo) * 2 + 0) * kThreads];
      const uint4 b1 = c.sm[((k - lo) * 2 + 1) * kThreads];
      const uint32_t bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint32_t       M0 = sadd2(bb[0], m[0]), M1 = sadd2(bb[0], n[0]);
#pragma unroll
      for (int i = 1; i < 8; i++) {
        M0 = max2(M0, sadd2(bb[i], m[i]));
        M1 = max2(M1, sadd2(bb[i], n[i]));
      }
      uint32_t o = ssub2(M1, M0);
      if (W == 8) o = sra1_2(o);  // the 8-window (sse16) decoder halves its output
      store_out<W>(c, dec2, (uint32_t)k, o, aux);
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = max2(m[i], n[i]);
      if ((k & 1) == 0 && k != 0) normalize(a);
    }
  }
  __syncwarp();
}

// hard decision of this code block: bit n = (A[n] + E[n] > 0), MSB first.
template <int W>
__device__ void decide(const WinCtx<W>& c, uint8_t* out)
{
  constexpr int   WH  = W / 2;
  const uint16_t* A16 = reinterpret_cast<const uint16_t*>(c.A32);
  const uint16_t* E16 = reinterpret_cast<const uint16_t*>(c.E32);
  for (uint32_t j = (uint32_t)c.t; j < c.K / 8; j += WH) {
    uint32_t n = 8 * j;
    uint32_t d = __umulhi(n, c.mL);
    uint32_t k = n - d * c.L;
    uint32_t byte = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const int16_t v = (int16_t)(uint16_t)(A16[k * W + d] + E16[k * W + d]);
      byte            = (byte << 1) | (v > 0 ? 1u : 0u);
      if (++k == c.L) {
        k = 0;
        d++;
      }
    }
    out[j] = (uint8_t)byte;
  }
}

__device__ __forceinline__ uint32_t crc24_bytes_dev(int which, const uint8_t* p, uint32_t nbytes)
{
  uint32_t crc = 0;
  for (uint32_t i = 0; i < nbytes; i++) crc = ((crc << 8) ^ c_crc_tab[which][((crc >> 16) & 0xFFu) ^ p[i]]) & 0xFFFFFFu;
  return crc;
}

template <int W>
__global__ void __launch_bounds__(kThreads, 3) tdec_win_kernel(const TdecLaunch a)
{
  constexpr int WH  = W / 2;
  constexpr int CBW = 32 / WH;
  constexpr uint32_t KMAX = (W == 16) ? 6144u : 800u;
  extern __shared__ uint4 smem[];

  const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int      grp = lane / WH, t = lane % WH;
  const uint32_t slot = blockIdx.x * (kThreads / 32) + warp;

  for (;;) {
    uint32_t it = 0;
    if (lane == 0) it = atomicAdd(a.counter, 1u);
    it = __shfl_sync(0xFFFFFFFFu, it, 0);
    if (it >= a.n_items) break;
    const WorkItem wi     = a.items[it];
    const bool     active = grp < (int)wi.count;
    const uint32_t cb     = a.order[wi.first + (active ? grp : 0)];

    WinCtx<W> c;
    c.K  = wi.K;
    c.L  = c.K / W;
    c.f1 = wi.f1;
    c.f2 = wi.f2;
    c.mK = (uint32_t)(0x100000000ull / c.K);
    c.mL = (uint32_t)((0x100000000ull + c.L - 1) / c.L);
    c.t  = t;
    {
      const uint32_t d0 = 2 * (uint32_t)t, d1 = d0 + 1;
      c.base_lo = (d0 * (c.f1 + c.f2 * d0 * c.L)) & (W - 1);
      c.base_hi = (d1 * (c.f1 + c.f2 * d1 * c.L)) & (W - 1);
      c.inc_lo  = (2 * c.f2 * d0) & (W - 1);
      c.inc_hi  = (2 * c.f2 * d1) & (W - 1);
    }
    const int16_t* in = a.in + (size_t)cb * a.in_stride;
    c.sys  = reinterpret_cast<const uint32_t*>(in);
    c.par0 = reinterpret_cast<const uint32_t*>(in + (c.K + 32));
    c.par1 = reinterpret_cast<const uint32_t*>(in + 2 * (c.K + 32));
    c.tail = in + 3 * (c.K + 32);
    int16_t* ae = a.ws_ae + ((size_t)slot * CBW + grp) * 2 * KMAX;
    c.A32 = reinterpret_cast<uint32_t*>(ae);
    c.E32 = reinterpret_cast<uint32_t*>(ae + KMAX);
    c.chk = a.ws_chk + (size_t)slot * kMaxChunks * 8 * 32 + lane;
    c.sm  = smem + tid;

    for (uint32_t k = 0; k < c.L; k++) c.A32[k * WH + t] = 0;
    __syncwarp();

    uint8_t* out  = a.out + (size_t)cb * a.out_stride;
    uint32_t n    = 0, iters = 0;
    bool     done = false, ok = false;
    const int which = a.crc_mode == CRC_24A ? 0 : 1;
    do {
      half_iteration<W>(c, (n & 1) != 0, n > 0);
      n++;
      if (a.crc_mode != CRC_NONE) {
        if (!done && active) decide<W>(c, out);
        __syncwarp();
        uint32_t crc = 1;
        if (!done && active && t == 0) crc = crc24_bytes_dev(which, out, c.K / 8);
        crc = __shfl_sync(0xFFFFFFFFu, crc, grp * WH);
        if (!done) {
          iters = n;
          if (crc == 0) {
            ok   = true;
            done = true;
          }
        }
      }
    } while (n < a.max_iter && !__all_sync(0xFFFFFFFFu, done || !active));
    if (a.crc_mode == CRC_NONE) {
      if (active) decide<W>(c, out);
      iters = n;
    }
    if (active && t == 0) {
      if (a.n_iter) a.n_iter[cb] = (uint8_t)iters;
      if (a.crc_ok) a.crc_ok[cb] = ok ? 1 : 0;
    }
    __syncwarp();
  }
}

// ---- generic decoder (K <= 400): one thread per code block, natural order, wrapping arithmetic ----
struct GenCtx {
  uint32_t       K, f1, f2;
  const int16_t* in;   // natural 3i+j, tails at 3K
  int16_t*       A;    // [K]
  int16_t*       E;    // [K]
  uint4*         beta; // [(k)*32], offset by lane; k = 0..K+3
};

__device__ __forceinline__ int16_t w16(int v) { return (int16_t)v; }

__device__ void gen_half_iteration(const GenCtx& c, bool dec2, bool apriori)
{
  const int K = (int)c.K;
  int16_t   s[8];
  s[0] = 0;
  for (int i = 1; i < 8; i++) s[i] = (int16_t)kNegInf;
  auto pi = [&](int k) { return (uint32_t)(((c.f1 + c.f2 * (uint32_t)k) * (uint32_t)k) % c.K); };
  auto load = [&](int k, int& x, int& y, int& aux, uint32_t& p) {
    p = 0;
    if (k >= K) {  // tail rows: no a-priori
      const int r = k - K;
      x   = c.in[3 * K + (dec2 ? 6 : 0) + 2 * r];
      y   = c.in[3 * K + (dec2 ? 6 : 0) + 2 * r + 1];
      aux = 0;
    } else if (!dec2) {
      x   = c.in[3 * k];
      y   = c.in[3 * k + 1];
      aux = 0;
      if (apriori) {
        aux = c.A[k];
        x   = w16(x + aux);
      }
    } else {
      p   = pi(k);
      x   = c.E[p];
      y   = c.in[3 * k + 2];
      aux = x;
    }
  };
  for (int k = K + 2; k >= 0; k--) {
    int      x, y, aux;
    uint32_t p;
    load(k, x, y, aux, p);
    const int xy = w16(x + y);
    int16_t   m[8], n[8];
    m[0] = w16(s[4] + xy); m[1] = s[4];           m[2] = w16(s[5] + y);  m[3] = w16(s[5] + x);
    m[4] = w16(s[6] + x);  m[5] = w16(s[6] + y);  m[6] = s[7];           m[7] = w16(s[7] + xy);
    n[0] = s[0];           n[1] = w16(s[0] + xy); n[2] = w16(s[1] + x);  n[3] = w16(s[1] + y);
    n[4] = w16(s[2] + y);  n[5] = w16(s[2] + x);  n[6] = w16(s[3] + xy); n[7] = s[3];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = m[i] > n[i] ? m[i] : n[i];
    uint4 v;
    v.x = (uint16_t)s[0] | ((uint32_t)(uint16_t)s[1] << 16);
    v.y = (uint16_t)s[2] | ((uint32_t)(uint16_t)s[3] << 16);
    v.z = (uint16_t)s[4] | ((uint32_t)(uint16_t)s[5] << 16);
    v.w = (uint16_t)s[6] | ((uint32_t)(uint16_t)s[7] << 16);
    c.beta[(size_t)k * 32] = v;
    if ((k % 4) == 0 && k < K) {
      for (int i = 1; i < 8; i++) s[i] = w16(s[i] - s[0]);
      s[0] = 0;
    }
  }
  s[0] = 0;
  for (int i = 1; i < 8; i++) s[i] = (int16_t)kNegInf;
  for (int k = 1; k <= K; k++) {
    int      x, y, aux;
    uint32_t p;
    load(k - 1, x, y, aux, p);
    const int xy = w16(x + y);
    int16_t   m[8], n[8], b[8];
    m[0] = s[0];           m[1] = w16(s[3] + y);  m[2] = w16(s[4] + y);  m[3] = s[7];
    m[4] = s[1];           m[5] = w16(s[2] + y);  m[6] = w16(s[5] + y);  m[7] = s[6];
    n[0] = w16(s[1] + xy); n[1] = w16(s[2] + x);  n[2] = w16(s[5] + x);  n[3] = w16(s[6] + xy);
    n[4] = w16(s[0] + xy); n[5] = w16(s[3] + x);  n[6] = w16(s[4] + x);  n[7] = w16(s[7] + xy);
    const uint4 v = c.beta[(size_t)k * 32];
    b[0] = (int16_t)(v.x & 0xFFFF); b[1] = (int16_t)(v.x >> 16); b[2] = (int16_t)(v.y & 0xFFFF); b[3] = (int16_t)(v.y >> 16);
    b[4] = (int16_t)(v.z & 0xFFFF); b[5] = (int16_t)(v.z >> 16); b[6] = (int16_t)(v.w & 0xFFFF); b[7] = (int16_t)(v.w >> 16);
    int16_t M0 = w16(m[0] + b[0]), M1 = w16(n[0] + b[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) {
      const int16_t c0 = w16(m[i] + b[i]), c1 = w16(n[i] + b[i]);
      if (c0 > M0) M0 = c0;
      if (c1 > M1) M1 = c1;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = m[i] > n[i] ? m[i] : n[i];
    if ((k % 4) == 0) {
      for (int i = 1; i < 8; i++) s[i] = w16(s[i] - s[0]);
      s[0] = 0;
    }
    const int16_t o = w16(M1 - M0);
    if (!dec2)
      c.E[k - 1] = w16(o - aux);
    else
      c.A[p] = w16(o - aux);
  }
}

__global__ void __launch_bounds__(kThreads) tdec_gen_kernel(const TdecLaunch a)
{
  constexpr uint32_t KMAX = 400;
  const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t slot = blockIdx.x * (kThreads / 32) + warp;
  for (;;) {
    uint32_t it = 0;
    if (lane == 0) it = atomicAdd(a.counter, 1u);
    it = __shfl_sync(0xFFFFFFFFu, it, 0);
    if (it >= a.n_items) break;
    const WorkItem wi     = a.items[it];
    const bool     active = lane < (int)wi.count;
    const uint32_t cb     = a.order[wi.first + (active ? lane : 0)];
    GenCtx c;
    c.K    = wi.K;
    c.f1   = wi.f1;
    c.f2   = wi.f2;
    c.in   = a.in + (size_t)cb * a.in_stride;
    c.A    = a.ws_ae + ((size_t)slot * 32 + lane) * 2 * KMAX;
    c.E    = c.A + KMAX;
    c.beta = reinterpret_cast<uint4*>(a.ws_chk) + (size_t)slot * (KMAX + 4) * 32 + lane;
    for (uint32_t k = 0; k < c.K; k++) c.A[k] = 0;

    uint8_t* out  = a.out + (size_t)cb * a.out_stride;
    uint32_t n    = 0, iters = 0;
    bool     done = false, ok = false;
    const int which = a.crc_mode == CRC_24A ? 0 : 1;
    auto decide_gen = [&]() {
      for (uint32_t j = 0; j < c.K / 8; j++) {
        uint32_t byte = 0;
        for (int b = 0; b < 8; b++) byte = (byte << 1) | (w16(c.A[8 * j + b] + c.E[8 * j + b]) > 0 ? 1u : 0u);
        out[j] = (uint8_t)byte;
      }
    };
    do {
      gen_half_iteration(c, (n & 1) != 0, n > 0);
      n++;
      if (a.crc_mode != CRC_NONE && !done && active) {
        decide_gen();
        iters = n;
        if (crc24_bytes_dev(which, out, c.K / 8) == 0) {
          ok   = true;
          done = true;
        }
      }
    } while (n < a.max_iter && !__all_sync(0xFFFFFFFFu, done || !active));
    if (a.crc_mode == CRC_NONE) {
      if (active) decide_gen();
      iters = n;
    }
    if (active) {
      if (a.n_iter) a.n_iter[cb] = (uint8_t)iters;
      if (a.crc_ok) a.crc_ok[cb] = ok ? 1 : 0;
    }
    __syncwarp();
  }
}

// ---- layout conversion: natural 3i+j -> working layout -----------------------------------------
__global__ void natural_to_working_kernel(const int16_t* __restrict__ nat, uint32_t nat_stride,
                                          int16_t* __restrict__ work, uint32_t work_stride,
                                          const uint32_t* __restrict__ cb_K, uint32_t uniform_K)
{
  const uint32_t cb = blockIdx.x;
  const uint32_t K  = cb_K ? cb_K[cb] : uniform_K;
  const int16_t* src = nat + (size_t)cb * nat_stride;
  int16_t*       dst = work + (size_t)cb * work_stride;
  const uint32_t W = (K % 16 == 0 && K > 800) ? 16u : (K % 8 == 0 && K > 400) ? 8u : 0u;
  if (W == 0) {
    for (uint32_t i = threadIdx.x; i < 3 * K + 12; i += blockDim.x) dst[i] = src[i];
    return;
  }
  const uint32_t L = K / W;
  for (uint32_t o = threadIdx.x; o < 3 * K; o += blockDim.x) {
    const uint32_t j = o / K, s = o - j * K;
    const uint32_t k = s / W, d = s - k * W;
    dst[j * (K + 32) + s] = src[3 * (d * L + k) + j];
  }
  for (uint32_t i = threadIdx.x; i < 12; i += blockDim.x) dst[3 * (K + 32) + i] = src[3 * K + i];
}

// ---- rate de-matching ----------------------------------------------------------------------------
__global__ void rm_rx_kernel(const int16_t* __restrict__ e, int16_t* __restrict__ work,
                             const uint16_t* __restrict__ tab_pool, const RmItem* __restrict__ items)
{
  const RmItem    it  = items[blockIdx.x];
  const int16_t*  src = e + it.e_off;
  int16_t*        dst = work + it.work_off;
  const uint16_t* tab = tab_pool + it.tab_off;
  for (uint32_t i = threadIdx.x; i < it.N && i < it.E; i += blockDim.x) {
    int acc = 0;
    for (uint32_t p = i; p < it.E; p += it.N) acc += src[p];  // wrap-around repeats hit the same cell
    const uint32_t o = tab[i];
    dst[o] = (int16_t)(dst[o] + acc);  // wrapping int16, like the reference's `+=`
  }
}

}  // namespace

// ---- host side -----------------------------------------------------------------------------------
void upload_crc_tables()
{
  uint32_t tab[2][256];
  const uint32_t polys[2] = {0x1864CFBu, 0x1800063u};
  for (int w = 0; w < 2; w++)
    for (uint32_t b = 0; b < 256; b++) {
      uint32_t r = b << 16;
      for (int i = 0; i < 8; i++) r = (r & 0x800000u) ? ((r << 1) ^ polys[w]) : (r << 1);
      tab[w][b] = r & 0xFFFFFFu;
    }
  cudaMemcpyToSymbol(c_crc_tab, tab, sizeof(tab));
}

int tdec_blocks_per_warp(int W) { return W == 16 ? 4 : W == 8 ? 8 : 32; }

cudaError_t tdec_geometry(int W, int device, TdecGeometry* g)
{
  int sms = 0;
  cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) return e;
  g->threads = kThreads;
  const size_t warps_per_block = kThreads / 32;
  if (W == 0) {
    g->smem   = 0;
    g->blocks = sms * 4;
    const size_t slots = (size_t)g->blocks * warps_per_block;
    g->ws_ae_bytes  = slots * 32 * 2 * 400 * sizeof(int16_t);
    g->ws_chk_bytes = slots * (400 + 4) * 32 * sizeof(uint4);
    return cudaSuccess;
  }
  g->smem = (size_t)kChunk * 2 * kThreads * sizeof(uint4);
  int per_sm = 0;
  if (W == 16) {
    e = cudaFuncSetAttribute(tdec_win_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tdec_win_kernel<16>, kThreads, g->smem);
  } else {
    e = cudaFuncSetAttribute(tdec_win_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tdec_win_kernel<8>, kThreads, g->smem);
  }
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  g->blocks = sms * per_sm;
  const size_t slots = (size_t)g->blocks * warps_per_block;
  const size_t kmax  = W == 16 ? 6144 : 800;
  g->ws_ae_bytes  = slots * (size_t)tdec_blocks_per_warp(W) * 2 * kmax * sizeof(int16_t);
  g->ws_chk_bytes = slots * kMaxChunks * 8 * 32 * sizeof(uint32_t);
  return cudaSuccess;
}

cudaError_t tdec_launch(int W, const TdecGeometry& g, const TdecLaunch& a, cudaStream_t s)
{
  if (a.n_items == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(a.counter, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  const int blocks = (int)((a.n_items + (kThreads / 32) - 1) / (kThreads / 32)) < g.blocks
                         ? (int)((a.n_items + (kThreads / 32) - 1) / (kThreads / 32))
                         : g.blocks;
  if (W == 16)
    tdec_win_kernel<16><<<blocks, g.threads, g.smem, s>>>(a);
  else if (W == 8)
    tdec_win_kernel<8><<<blocks, g.threads, g.smem, s>>>(a);
  else
    tdec_gen_kernel<<<blocks, g.threads, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t natural_to_working_launch(const int16_t* nat, uint32_t nat_stride, int16_t* work, uint32_t work_stride,
                                      const uint32_t* cb_K, uint32_t uniform_K, uint32_t n_cb, cudaStream_t s)
{
  if (n_cb == 0) return cudaSuccess;
  natural_to_working_kernel<<<n_cb, 256, 0, s>>>(nat, nat_stride, work, work_stride, cb_K, uniform_K);
  return cudaGetLastError();
}

cudaError_t rm_rx_launch(const int16_t* e, int16_t* work, const uint16_t* tab_pool, const RmItem* items,
                         uint32_t n_items, cudaStream_t s)
{
  if (n_items == 0) return cudaSuccess;
  rm_rx_kernel<<<n_items, 256, 0, s>>>(e, work, tab_pool, items);
  return cudaGetLastError();
}

}  // namespace b200
