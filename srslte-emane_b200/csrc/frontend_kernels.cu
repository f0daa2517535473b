// frontend_kernels.cu -- the stage in front of rate de-matching: equalised symbols ->
// int16 LLRs (soft demodulation) -> descrambling, alone or fused into the rate-dematching scatter-add so that the
// e_bits array of the reference (pdsch.c:760-779, pusch.c:482-500) is never written to HBM.
//
// What is computed (bit-exact with the reference's AVX2/SSE build for |scale * x| < 32768):
//   srslte_demod_soft_demodulate_s   lib/src/phy/modem/demod_soft.c:503-525
//       QPSK   :68-70 -> srslte_vec_convert_fi_simd (vector_simd.c:392-427): x * (float)(-100 sqrt 2), truncated;
//              saturating in the 16-wide AVX2 body, plain cast in the remainder
//       16QAM  :90-133   groups of 4 symbols: round-to-nearest-even of x * -400, saturate, |.| - 252 (wrapping);
//              remainder :121-131: truncation and a double-precision offset
//       64QAM  :240-302  the same with -700 and offsets 432 / 216; remainder :290-300
//       256QAM :457-477  scalar float chain, truncation
//   srslte_scrambling_s_offset       lib/src/phy/scrambling/scrambling.c:44-47 (sign flip where c(n) = 1)
//   ulsch_deinterleave (no UCI)      lib/src/phy/phch/sch.c:580-598, 891-918: g[(j*cols + i)*Qm + k] = q[(i*rows + j)*Qm + k];
//       applied as an index map in front of the LLR computation, so the permuted array never exists
//   the scrambling sequence c(n)     lib/src/phy/common/sequence.c:46-75 (36.211 7.2, Nc = 1600); here
//       c(n) = x1(n + 1600) xor parity(mask[n] & c_init): x2 is linear in its seed, so one table of 31-bit masks
//       (lte_tables.cpp:gold_tables) serves every seed and every LLR is computed independently.
#include "tdec_kernels.h"

namespace b200 {

namespace {

constexpr uint32_t kLlrPerThread = 8;

__device__ __forceinline__ int sat16(int v) { return max(-32768, min(32767, v)); }
__device__ __forceinline__ int wrap16(int v) { return (int)(int16_t)v; }

// UL-SCH order -> channel order for a codeword with multiplexed UCI (data path of srslte_ulsch_decode,
// lib/src/phy/phch/sch.c:920-1064).  The reference marks the Q'_ri * Qm RI positions (uci_ulsch_interleave_ri_gen,
// uci.c:522-545: symbol idx sits in row rows-1-idx/4, column set[(3 idx) % 4]), numbers the remaining matrix cells in
// row-major order (ulsch_interleave_gen, sch.c:580-598) and scatters g[lut[x]] = q[x]; here the inverse is closed
// form: rows above r0 = rows - ceil(Q'_ri / 4) are complete, row r0 misses Q'_ri - 4 (rows-1-r0) cells, the rows below
// miss 4.  ACK symbols (uci_ulsch_interleave_ack_gen, uci.c:497-520) use the same row rule on their own column set
// and are erased (sch.c:961-964).  Returns the channel position, kUciErased for an LLR the reference sets to 0 (or
// never defines: j >= (H' - Q'_ri) Qm), with kUciRaw set when the sample is used without descrambling (g0_raw).
constexpr uint32_t kUciErased = 0xFFFFFFFFu, kUciRaw = 0x80000000u;
__device__ __noinline__ uint32_t ul_uci_map(const FeCodeword& cw, uint32_t j, uint32_t qm)
{
  const uint32_t rows = cw.ul_rows, cols = cw.ul_cols;
  const bool     norm = cols > 10;
  if (j == 0 && cw.g0_src != kNoG0) return cw.g0_src | (cw.g0_raw ? kUciRaw : 0u);
  const uint32_t v = j / qm, k = j - v * qm;
  if (v >= rows * cols - cw.q_ri) return kUciErased;
  const uint32_t ri_rows = (cw.q_ri + 3) / 4, r0 = rows - ri_rows;
  uint32_t       row, col;
  if (v < r0 * cols) {
    row = v / cols;
    col = v - row * cols;
  } else {
    uint32_t       t = v - r0 * cols;
    const uint32_t nri0 = cw.q_ri - 4 * (ri_rows - 1);  // 1..4 RI cells in the first RI row
    uint32_t       nri = nri0;
    row = r0;
    if (t >= cols - nri0) {
      t -= cols - nri0;
      const uint32_t d = t / (cols - 4);
      row = r0 + 1 + d;
      t -= d * (cols - 4);
      nri = 4;
    }
    col = t;  // t-th column of the row that carries no RI; symbol 4g + m of a row uses set[(3m) % 4]: m = 0, 1, 2, 3
              // -> set[0], set[3], set[2], set[1]
    if (col >= uci_col(true, norm, 0)) col++;
    if (nri >= 4 && col >= uci_col(true, norm, 1)) col++;
    if (nri >= 3 && col >= uci_col(true, norm, 2)) col++;
    if (nri >= 2 && col >= uci_col(true, norm, 3)) col++;
  }
  if (cw.q_ack) {
#pragma unroll
    for (uint32_t c = 0; c < 4; c++)
      if (col == uci_col(false, norm, c) && 4 * (rows - 1 - row) + ((4 - c) & 3u) < cw.q_ack) return kUciErased;
  }
  return (col * rows + row) * qm + k;
}

// The modulation's level chain for channel position j (symbol j / QM, level (j % QM) / 2, x = its real or imaginary
// part), not yet descrambled.  Which conversion rule applies depends on where the symbol sits relative to the reference's
// SIMD bodies and scalar remainders (see the file header).
// BODY: the caller knows that the symbol lies in the reference's SIMD body (no per-LLR test, no remainder code)
template <uint32_t QM, bool BODY = false>
__device__ __forceinline__ int fe_value(const FeCodeword& cw, float x, uint32_t j, uint32_t s, uint32_t lvl)
{
  constexpr uint32_t qm = QM;
  int                v;
  if (qm == 2) {
    constexpr float kScale = (float)(-100.0 * 1.4142135623730951);
    const int       t = __float2int_rz(__fmul_rn(x, kScale));
    v = BODY || j < ((2 * cw.nsym) & ~15u) ? sat16(t) : wrap16(t);
  } else if (qm == 4) {
    if (BODY || s < (cw.nsym & ~3u)) {
      const int v0 = sat16(__float2int_rn(__fmul_rn(x, -400.0f)));
      v = lvl == 0 ? v0 : wrap16(abs(v0) - 252);
    } else {
      const int y = wrap16(__float2int_rz(__fmul_rn(400.0f, x)));
      v = lvl == 0 ? wrap16(-y) : wrap16(__double2int_rz((double)abs(y) - 800.0 / 3.1622776601683795));
    }
  } else if (qm == 6) {
    if (BODY || s < (cw.nsym & ~3u)) {
      const int v0 = sat16(__float2int_rn(__fmul_rn(x, -700.0f)));
      const int a1 = wrap16(abs(v0) - 432);
      v = lvl == 0 ? v0 : lvl == 1 ? a1 : wrap16(abs(a1) - 216);
    } else {
      const int y  = wrap16(__float2int_rz(__fmul_rn(700.0f, x)));
      const int l2 = wrap16(__double2int_rz((double)abs(y) - 2800.0 / 6.48074069840786));
      v = lvl == 0 ? wrap16(-y) : lvl == 1 ? l2 : wrap16(__double2int_rz((double)abs(l2) - 1400.0 / 6.48074069840786));
    }
  } else {
    // 8 / sqrtf(170), 4 / sqrtf(170), 2 / sqrtf(170) in single precision (demod_soft.c:457-477), as constants:
    // sqrtf(170.0f) = 0x1.a13a9cp+3, the three quotients rounded to nearest differ in the exponent only
    constexpr float k8 = 0x1.3a261cp-1f, k4 = 0x1.3a261cp-2f, k2 = 0x1.3a261cp-3f;
    float           f = -x;
    if (lvl >= 1) f = __fsub_rn(fabsf(f), k8);
    if (lvl >= 2) f = __fsub_rn(fabsf(f), k4);
    if (lvl >= 3) f = __fsub_rn(fabsf(f), k2);
    v = wrap16(__float2int_rz(__fmul_rn(1000.0f, f)));
  }
  return v;
}

// LLR number j of a codeword (before rate de-matching), descrambled.  QM = bits per symbol, a compile-time constant:
// the kernels branch once per CTA on the codeword's modulation (no division by a run-time Qm, no per-LLR dispatch).
template <uint32_t QM>
__device__ __forceinline__ int fe_llr(const FeCodeword& cw, const float* __restrict__ sym, uint32_t j,
                                      const uint32_t* __restrict__ x1, const uint32_t* __restrict__ x2mask)
{
  constexpr uint32_t qm = QM;
  bool               raw = false;  // g[0] quirk with a 1-bit RI: the sample is used as demodulated (not descrambled)
  if (cw.ul_cols) {  // j counts in UL-SCH order: vector v = j / qm sits at row v / cols, column v % cols of the
                     // interleaver matrix and was sent as vector column * rows + row
    if (cw.q_ack | cw.q_ri) {
      const uint32_t x = ul_uci_map(cw, j, qm);
      if (x == kUciErased) return 0;
      raw = (x & kUciRaw) != 0;
      j   = x & ~kUciRaw;
    } else {
      const uint32_t v = j / qm, k = j - v * qm, row = v / cw.ul_cols, col = v - row * cw.ul_cols;
      j = (col * cw.ul_rows + row) * qm + k;
    }
  }
  const uint32_t s = j / qm, r = j - s * qm;
  int            v = fe_value<QM>(cw, __ldg(sym + 2 * (size_t)s + (r & 1u)), j, s, r >> 1);
  if (j < cw.nof_bits) {
    const uint32_t c = ((__ldg(x1 + (j >> 5)) >> (j & 31u)) ^ (uint32_t)__popc(__ldg(x2mask + j) & cw.c_init)) & 1u;
    if (c && !raw) v = wrap16(-v);
  }
  return v;
}

// G consecutive LLRs of a PDSCH codeword (channel order = output order) by one thread: whole symbols in (one or two
// 128-bit loads), whole 128-bit stores out, the x1 bits from one or two words and the x2 masks as 128-bit loads.
// G = 8 (4 QPSK / 2 16QAM / 1 256QAM symbols) or 24 (4 64QAM symbols); j0 is a multiple of G.
template <uint32_t QM>
__device__ __forceinline__ void fe_group(const FeCodeword& cw, const float* __restrict__ sym, int (&v)[QM == 6 ? 24 : 8],
                                         uint32_t j0, const uint32_t* __restrict__ x1, const uint32_t* __restrict__ x2mask)
{
  constexpr uint32_t G = QM == 6 ? 24 : 8, NF = 2 * G / QM;
  float              x[NF];
  const float*       p = sym + 2 * (size_t)(j0 / QM);
  if constexpr (NF == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    x[0] = t.x;
    x[1] = t.y;
  } else {
#pragma unroll
    for (uint32_t i = 0; i < NF / 4; i++) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
      x[4 * i] = t.x, x[4 * i + 1] = t.y, x[4 * i + 2] = t.z, x[4 * i + 3] = t.w;
    }
  }
  const uint32_t s0 = j0 / QM;  // j0 is a multiple of G, G of QM
  // the whole group lies in the reference's SIMD body (QPSK: 16 LLRs at a time, 16QAM / 64QAM: 4 symbols at a time; the
  // 256QAM chain has no remainder rule) or it is the one group per codeword that touches the scalar remainder
  const bool body = QM == 8 || (QM == 2 ? j0 + G <= ((2 * cw.nsym) & ~15u) : s0 + G / QM <= (cw.nsym & ~3u));
  if (body) {
#pragma unroll
    for (uint32_t i = 0; i < G; i++)
      v[i] = fe_value<QM, true>(cw, x[2 * (i / QM) + ((i % QM) & 1u)], j0 + i, s0 + i / QM, (i % QM) >> 1);
  } else {
#pragma unroll
    for (uint32_t i = 0; i < G; i++)
      v[i] = fe_value<QM, false>(cw, x[2 * (i / QM) + ((i % QM) & 1u)], j0 + i, s0 + i / QM, (i % QM) >> 1);
  }
  if (j0 + G <= cw.nof_bits) {
    const uint32_t sh = j0 & 31u, w0 = __ldg(x1 + (j0 >> 5));
    const uint32_t w1 = sh + G > 32 ? __ldg(x1 + (j0 >> 5) + 1) : 0u;
    const uint32_t w  = __funnelshift_r(w0, w1, sh);
#pragma unroll
    for (uint32_t q = 0; q < G / 4; q++) {
      const uint4    m    = __ldg(reinterpret_cast<const uint4*>(x2mask + j0) + q);
      const uint32_t mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (uint32_t k = 0; k < 4; k++) {
        const uint32_t i = 4 * q + k;
        if (((w >> i) ^ (uint32_t)__popc(mm[k] & cw.c_init)) & 1u) v[i] = wrap16(-v[i]);
      }
    }
  } else {
#pragma unroll
    for (uint32_t i = 0; i < G; i++) {
      const uint32_t j = j0 + i;
      if (j < cw.nof_bits && (((__ldg(x1 + (j >> 5)) >> (j & 31u)) ^ (uint32_t)__popc(__ldg(x2mask + j) & cw.c_init)) & 1u))
        v[i] = wrap16(-v[i]);
    }
  }
}

// The same for a PUSCH codeword without multiplexed control information: the G output LLRs are G / QM consecutive
// vectors of the UL-SCH order; vector v sits at row v / cols, column v % cols of the channel interleaver matrix and was
// sent as symbol column * rows + row (ulsch_deinterleave, sch.c:891-918), where it was demodulated and descrambled.
template <uint32_t QM>
__device__ __forceinline__ void fe_group_ul(const FeCodeword& cw, const float* __restrict__ sym, int (&v)[QM == 6 ? 24 : 8],
                                            uint32_t j0, const uint32_t* __restrict__ x1, const uint32_t* __restrict__ x2mask)
{
  constexpr uint32_t G = QM == 6 ? 24 : 8, NV = G / QM;
  const uint32_t     v0 = j0 / QM;
  uint32_t           row = v0 / cw.ul_cols, col = v0 - row * cw.ul_cols;
#pragma unroll
  for (uint32_t u = 0; u < NV; u++) {
    const uint32_t s = col * cw.ul_rows + row, jc = s * QM;
    if (++col == cw.ul_cols) col = 0, row++;
    const float2 t = __ldg(reinterpret_cast<const float2*>(sym) + s);
    const bool   body = QM == 8 || (QM == 2 ? jc + QM <= ((2 * cw.nsym) & ~15u) : s < (cw.nsym & ~3u));
    if (body) {
#pragma unroll
      for (uint32_t k = 0; k < QM; k++) v[u * QM + k] = fe_value<QM, true>(cw, (k & 1u) ? t.y : t.x, jc + k, s, k >> 1);
    } else {
#pragma unroll
      for (uint32_t k = 0; k < QM; k++) v[u * QM + k] = fe_value<QM, false>(cw, (k & 1u) ? t.y : t.x, jc + k, s, k >> 1);
    }
    if (jc + QM <= cw.nof_bits) {
      const uint32_t sh = jc & 31u, w0 = __ldg(x1 + (jc >> 5));
      const uint32_t w1 = sh + QM > 32 ? __ldg(x1 + (jc >> 5) + 1) : 0u;
      const uint32_t w  = __funnelshift_r(w0, w1, sh);
#pragma unroll
      for (uint32_t k = 0; k < QM; k += 2) {  // jc is a multiple of QM, QM is even: 64-bit loads are aligned
        const uint2 m = __ldg(reinterpret_cast<const uint2*>(x2mask + jc + k));
        if (((w >> k) ^ (uint32_t)__popc(m.x & cw.c_init)) & 1u) v[u * QM + k] = wrap16(-v[u * QM + k]);
        if (((w >> (k + 1)) ^ (uint32_t)__popc(m.y & cw.c_init)) & 1u) v[u * QM + k + 1] = wrap16(-v[u * QM + k + 1]);
      }
    } else {
#pragma unroll
      for (uint32_t k = 0; k < QM; k++) {
        const uint32_t j = jc + k;
        if (j < cw.nof_bits && (((__ldg(x1 + (j >> 5)) >> (j & 31u)) ^ (uint32_t)__popc(__ldg(x2mask + j) & cw.c_init)) & 1u))
          v[u * QM + k] = wrap16(-v[u * QM + k]);
      }
    }
  }
}

// G LLRs of a codeword starting at j0 (a multiple of G), PDSCH or PUSCH without control information
template <uint32_t QM>
__device__ __forceinline__ void fe_group_any(const FeCodeword& cw, const float* __restrict__ sym, int (&v)[QM == 6 ? 24 : 8],
                                             uint32_t j0, const uint32_t* __restrict__ x1, const uint32_t* __restrict__ x2mask)
{
  if (cw.ul_cols)
    fe_group_ul<QM>(cw, sym, v, j0, x1, x2mask);
  else
    fe_group<QM>(cw, sym, v, j0, x1, x2mask);
}

// which codewords the group functions handle: PDSCH with its symbols on a 128-bit boundary (64 bits for 256QAM), PUSCH
// without multiplexed control information (with it: the general index map ul_uci_map, one LLR at a time)
template <uint32_t QM>
__device__ __forceinline__ bool fe_group_ok(const FeCodeword& cw, const float* sym)
{
  return cw.ul_cols ? (cw.q_ack | cw.q_ri) == 0 && cw.g0_src == kNoG0 && (reinterpret_cast<uintptr_t>(sym) & 7u) == 0
                    : (reinterpret_cast<uintptr_t>(sym) & (QM == 8 ? 7u : 15u)) == 0;
}

template <uint32_t QM>
__device__ __forceinline__ void demod_descramble_cw(const FeCodeword& cw, const float* __restrict__ sym,
                                                    int16_t* __restrict__ out, const uint32_t* __restrict__ x1,
                                                    const uint32_t* __restrict__ x2mask)
{
  constexpr uint32_t G = QM == 6 ? 24 : 8;
  const uint32_t     n = QM * cw.nsym;
  // codewords whose LLRs (and, for PDSCH, symbols) sit on 128-bit boundaries: a thread per G LLRs, 128-bit stores
  const bool     fast   = (reinterpret_cast<uintptr_t>(out) & 15u) == 0 && fe_group_ok<QM>(cw, sym);
  const uint32_t groups = fast ? n / G : 0;
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
    int v[G];
    fe_group_any<QM>(cw, sym, v, g * G, x1, x2mask);
#pragma unroll
    for (uint32_t q = 0; q < G / 8; q++) {
      uint4 o;
      o.x = (uint32_t)(v[8 * q + 0] & 0xFFFF) | ((uint32_t)v[8 * q + 1] << 16);
      o.y = (uint32_t)(v[8 * q + 2] & 0xFFFF) | ((uint32_t)v[8 * q + 3] << 16);
      o.z = (uint32_t)(v[8 * q + 4] & 0xFFFF) | ((uint32_t)v[8 * q + 5] << 16);
      o.w = (uint32_t)(v[8 * q + 6] & 0xFFFF) | ((uint32_t)v[8 * q + 7] << 16);
      reinterpret_cast<uint4*>(out + g * G)[q] = o;
    }
  }
  // what is left (fewer than G LLRs) and unaligned buffers: one LLR at a time, consecutive threads write consecutive
  // LLRs
  const uint32_t first = groups * G, st = blockDim.x, per = blockDim.x * kLlrPerThread;
  for (uint32_t base = first + blockIdx.x * per; base < n; base += gridDim.x * per) {
    const uint32_t j1 = min(n, base + per);
    for (uint32_t j = base + threadIdx.x; j < j1; j += st) out[j] = (int16_t)fe_llr<QM>(cw, sym, j, x1, x2mask);
  }
}

// grid: (chunks of 256 * kLlrPerThread LLRs, codewords)
__global__ void __launch_bounds__(256) demod_descramble_kernel(const FeCodeword* __restrict__ cws,
                                                               const float* __restrict__ symbols,
                                                               int16_t* __restrict__ e, const uint32_t* __restrict__ x1,
                                                               const uint32_t* __restrict__ x2mask)
{
  const FeCodeword cw = cws[blockIdx.y];
  const float*     sym = symbols + 2 * cw.sym_off;
  int16_t*         out = e + cw.llr_off;
  switch (cw.qm) {
    case 2: demod_descramble_cw<2>(cw, sym, out, x1, x2mask); break;
    case 4: demod_descramble_cw<4>(cw, sym, out, x1, x2mask); break;
    case 6: demod_descramble_cw<6>(cw, sym, out, x1, x2mask); break;
    default: demod_descramble_cw<8>(cw, sym, out, x1, x2mask); break;
  }
}

// rate de-matching straight from the symbols: work[tab[i]] += sum over the wrap-around repeats of LLR(e_off + p).
// A block without wrap-around (E <= N, the usual case) of a codeword the group functions handle is produced G LLRs per
// thread, in groups aligned to the codeword (the first and the last group of a block are shared with its neighbours and
// computed by both); the table is one-to-one, so every thread owns the cells it writes.
template <uint32_t QM>
__device__ __forceinline__ void rm_rx_sym_cw(const RmSymItem& it, const FeCodeword& cw, const float* __restrict__ sym,
                                             int16_t* __restrict__ dst, const uint16_t* __restrict__ tab, int16_t* img,
                                             const uint32_t* __restrict__ x1, const uint32_t* __restrict__ x2mask)
{
  constexpr uint32_t G = QM == 6 ? 24 : 8;
  const uint32_t     e0 = it.e_off;
  if (it.E <= it.N && e0 + it.E <= QM * cw.nsym && fe_group_ok<QM>(cw, sym)) {
    const uint32_t wl8 = (it.wl + 7) & ~7u;
    for (uint32_t j = threadIdx.x; j < wl8 / 2; j += blockDim.x) reinterpret_cast<uint32_t*>(img)[j] = 0;
    __syncthreads();
    const uint32_t g0 = e0 / G, g1 = (e0 + it.E + G - 1) / G;
    for (uint32_t g = g0 + threadIdx.x; g < g1; g += blockDim.x) {
      if (g * G + G <= QM * cw.nsym) {
        int v[G];
        fe_group_any<QM>(cw, sym, v, g * G, x1, x2mask);
#pragma unroll
        for (uint32_t i = 0; i < G; i++) {
          const uint32_t p = g * G + i - e0;  // wraps below e0
          if (p < it.E) img[tab[p]] = (int16_t)v[i];
        }
      } else {  // the codeword ends inside this group
        for (uint32_t i = 0; i < G; i++) {
          const uint32_t p = g * G + i - e0;
          if (p < it.E) img[tab[p]] = (int16_t)fe_llr<QM>(cw, sym, g * G + i, x1, x2mask);
        }
      }
    }
    __syncthreads();
    rm_rx_add_image(img, dst, it.wl, it.overwrite != 0);
    return;
  }
  rm_rx_body([&](uint32_t p) { return fe_llr<QM>(cw, sym, e0 + p, x1, x2mask); }, it.E, it.N, it.wl, tab, dst, img,
             it.overwrite != 0);
}

__global__ void rm_rx_sym_kernel(const FeCodeword* __restrict__ cws, const float* __restrict__ symbols,
                                 int16_t* __restrict__ work, const uint16_t* __restrict__ tab_pool,
                                 const RmSymItem* __restrict__ items, const uint32_t* __restrict__ x1,
                                 const uint32_t* __restrict__ x2mask)
{
  extern __shared__ int16_t rm_img[];
  const RmSymItem  it  = items[blockIdx.x];
  const FeCodeword cw  = cws[it.cw];
  const float*     sym = symbols + 2 * cw.sym_off;
  const uint16_t*  tab = tab_pool + it.tab_off;
  int16_t*         dst = work + it.work_off;
  switch (cw.qm) {
    case 2: rm_rx_sym_cw<2>(it, cw, sym, dst, tab, rm_img, x1, x2mask); break;
    case 4: rm_rx_sym_cw<4>(it, cw, sym, dst, tab, rm_img, x1, x2mask); break;
    case 6: rm_rx_sym_cw<6>(it, cw, sym, dst, tab, rm_img, x1, x2mask); break;
    default: rm_rx_sym_cw<8>(it, cw, sym, dst, tab, rm_img, x1, x2mask); break;
  }
}

// ---- TX mirror: turbo encoder + rate matching, for vector generation on the device ----------
// reference: srslte_tcod_encode (lib/src/phy/fec/turbocoder.c:95-187: two RSC encoders g0 = 1 + D^2 + D^3,
// g1 = 1 + D + D^3, the second one on the QPP-interleaved bits, output 3i+j then the 12 tail bits) and
// srslte_rm_turbo_tx (lib/src/phy/fec/rm_turbo.c:303-372: e[i] = d[table[i mod N]], the same selection table the
// receive side scatters through).  One CTA per code block.  The encoder's register is a linear function of its
// start state and of the input bits, so the K-step chain is cut into 128 chunks per constituent encoder (threads 0-127:
// natural order, threads 128-255: QPP order): every thread runs its chunk from the zero state, one thread per encoder
// chains the 128 end states through the chunk's zero-input transition (an 8-entry table), and every thread runs its
// chunk again from its true start state and writes the parity bits.  The rate-matched output is gathered by everybody.
__device__ __forceinline__ uint32_t rsc_step(uint32_t& st, uint32_t in)
{
  // st = r0 | r1 << 1 | r2 << 2;  feedback g0 = 1 + D^2 + D^3, output g1 = 1 + D + D^3
  const uint32_t r0 = st & 1u, r1 = (st >> 1) & 1u, r2 = st >> 2;
  const uint32_t fb = in ^ r2 ^ r1;
  st = fb | (r0 << 1) | (r1 << 2);
  return r2 ^ r0 ^ fb;
}

constexpr uint32_t kTxThreads = 256, kTxChunks = kTxThreads / 2;

__global__ void __launch_bounds__(kTxThreads) tcod_rm_tx_kernel(const TxItem* __restrict__ items,
                                                                const uint8_t* __restrict__ bits_all,
                                                                uint8_t* __restrict__ e_all, const uint16_t* __restrict__ tab_pool)
{
  extern __shared__ uint8_t tx_d[];  // the 3K + 12 coded bits
  __shared__ uint8_t s_end[2][kTxChunks], s_start[2][kTxChunks], s_zero[8];
  const TxItem   it   = items[blockIdx.x];
  const uint32_t K    = it.K;
  const uint8_t* bits = bits_all + it.bits_off;
  for (uint32_t k = threadIdx.x; k < K; k += blockDim.x) tx_d[3 * k] = bits[k] & 1u;
  const uint32_t L = (K + kTxChunks - 1) / kTxChunks, nch = (K + L - 1) / L;  // L <= 48 steps per chunk
  const uint32_t second = threadIdx.x / kTxChunks, c = threadIdx.x % kTxChunks;
  const uint32_t k0 = c * L, k1 = min(K, k0 + L);
  uint64_t       in = 0;  // the chunk's input bits
  if (c < nch) {
    // pi(k) = (f1 k + f2 k^2) mod K and its increment pi(k+1) - pi(k) = f1 + f2 (2k + 1), both mod K
    uint32_t       pk = (uint32_t)(((uint64_t)it.f1 * k0 + (uint64_t)it.f2 * k0 % K * k0) % K);
    uint32_t       g  = (uint32_t)(((uint64_t)it.f1 + (uint64_t)it.f2 * (2 * k0 + 1)) % K);
    const uint32_t g2 = (2 * it.f2) % K;
    uint32_t       st = 0;
    for (uint32_t k = k0; k < k1; k++) {
      const uint32_t b = bits[second ? pk : k] & 1u;
      in |= (uint64_t)b << (k - k0);
      rsc_step(st, b);
      pk += g; if (pk >= K) pk -= K;
      g += g2; if (g >= K) g -= K;
    }
    s_end[second][c] = (uint8_t)st;
  }
  if (threadIdx.x < 8) {  // where L steps without input take each of the 8 states
    uint32_t st = threadIdx.x;
    for (uint32_t k = 0; k < L; k++) rsc_step(st, 0);
    s_zero[threadIdx.x] = (uint8_t)st;
  }
  __syncthreads();
  if (c == 0) {  // start states of the chunks (every chunk before the last one has L steps)
    uint32_t st = 0;
    for (uint32_t i = 0; i < nch; i++) {
      s_start[second][i] = (uint8_t)st;
      st = s_zero[st] ^ s_end[second][i];
    }
  }
  __syncthreads();
  if (c < nch) {
    uint32_t st = s_start[second][c];
    for (uint32_t k = k0; k < k1; k++) tx_d[3 * k + 1 + second] = (uint8_t)rsc_step(st, (uint32_t)(in >> (k - k0)) & 1u);
    if (c == nch - 1) {  // flush: the input equals the feedback, so 0 enters the register
      uint32_t r0 = st & 1u, r1 = (st >> 1) & 1u, r2 = st >> 2;
      uint8_t* tail = tx_d + 3 * K + (second ? 6 : 0);
      for (int j = 0; j < 3; j++) {
        tail[2 * j]     = (uint8_t)(r2 ^ r1);
        tail[2 * j + 1] = (uint8_t)(r2 ^ r0);
        r2 = r1; r1 = r0; r0 = 0;
      }
    }
  }
  __syncthreads();
  const uint16_t* tab = tab_pool + it.tab_off;
  uint8_t*        e   = e_all + it.e_off;
  const uint32_t  N   = 3 * K + 12;
  for (uint32_t i = threadIdx.x; i < it.E; i += blockDim.x) e[i] = tx_d[tab[i % N]];
}

}  // namespace

cudaError_t tcod_rm_tx_launch(const TxItem* items, uint32_t n_items, const uint8_t* bits, uint8_t* e,
                              const uint16_t* tab_pool, cudaStream_t s)
{
  if (n_items == 0) return cudaSuccess;
  tcod_rm_tx_kernel<<<n_items, kTxThreads, 3 * 6144 + 16, s>>>(items, bits, e, tab_pool);
  return cudaGetLastError();
}

cudaError_t demod_descramble_launch(const FeCodeword* cws, uint32_t n_cw, uint32_t max_llr, const float* symbols,
                                    int16_t* e, const uint32_t* x1, const uint32_t* x2mask, cudaStream_t s)
{
  if (n_cw == 0 || max_llr == 0) return cudaSuccess;
  // enough CTAs for a few waves of the 148 SMs; a CTA strides over its codeword (empty CTAs cost a launch slot each)
  const uint32_t chunks = (max_llr + 256 * kLlrPerThread - 1) / (256 * kLlrPerThread);
  const uint32_t cap    = n_cw >= 4736 ? 1u : 4736u / n_cw;
  dim3           grid(chunks < cap ? chunks : cap, n_cw);
  demod_descramble_kernel<<<grid, 256, 0, s>>>(cws, symbols, e, x1, x2mask);
  return cudaGetLastError();
}

cudaError_t rm_rx_sym_launch(const FeCodeword* cws, const float* symbols, int16_t* work, const uint16_t* tab_pool,
                             const RmSymItem* items, uint32_t n_items, const uint32_t* x1, const uint32_t* x2mask,
                             cudaStream_t s)
{
  if (n_items == 0) return cudaSuccess;
  rm_rx_sym_kernel<<<n_items, 256, (kRmMaxWorkLen + 8) * sizeof(int16_t), s>>>(cws, symbols, work, tab_pool, items, x1,
                                                                               x2mask);
  return cudaGetLastError();
}

}  // namespace b200
