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
//     two int16 halves of every 32-bit register.  W/2 threads per block, 4 (W=16) or 8 (W=8) code
//     blocks per warp, the 8 state metrics live in registers.
//   * inputs are held in an item-interleaved "pair-major" layout: the words a thread needs for 4 consecutive
//     trellis rows are contiguous (one 128-bit shared-memory load feeds 4 steps), and the row groups of all the
//     code blocks of a warp are adjacent, so ONE TMA bulk copy per stream stages an 8-row chunk for the warp.
//   * the reference keeps all beta metrics of a half iteration (98 KB per K=6144 block).  Here the
//     backward pass keeps one checkpoint per 16 trellis rows; the forward pass rebuilds the 16 rows
//     of beta it needs in shared memory (the recursion and its normalisation points depend only on
//     the row index, so the rebuilt values are identical).
//   * the two extrinsic arrays are stored already differenced (what the reference computes with
//     srslte_vec_sub_sss at the start of the next half iteration), so the a-posteriori value of
//     every bit is always A + E (wrapping) and no third array is needed.
//   * QPP addresses are computed in the kernel from (f1, f2): pi(d*L + k) shares its row
//     pi(k) mod L across all windows (contention-free property), only the window index differs.  Each
//     extrinsic array is stored in the order its CONSUMER reads it: the producer scatters on its single
//     write (pi for DEC2's output, pi^-1 for DEC1's), every read is linear and goes through TMA.
//   * saturating int16x2 adds are 6 instructions on sm_100a, wrapping adds and fused add-max are 1.
//     Every half iteration first runs a FAST variant (wrapping VIADD.16x2 on the FMA pipe, fused
//     VIADDMNMX.S16x2 / VIMNMX.S16x2 on the ALU pipe) that also records the extremes of its state
//     metrics; from them it PROVES that no saturating operation of the reference could have
//     clamped (then wrapping == saturating, bit for bit).  If the proof fails, the half iteration
//     is re-run with the EXACT saturating variant.  Both run on the GPU; there is no CPU path.
#include "tdec_kernels.h"

#include <cstdio>
#include <cstdlib>

namespace b200 {

namespace {

constexpr int      kWarm      = 40;  // win_overlap_len
constexpr int      kThreads   = 256; // one CTA per SM; its warps decode the work items of one CTA round in step
constexpr int      kWarps     = kThreads / 32;
constexpr int      kGenThreads = 384; // generic decoder (K <= 400)
constexpr int      kBlocksPerSm = 1;
constexpr int      kMaxL      = 384;
constexpr int      kMaxGroups = kMaxL / 4;  // row groups of 4 trellis rows
constexpr uint32_t kGroupBytes = 512;       // one row group of a warp in every stream / array: [lane][4 rows] 32-bit words
// ---- main fast path: beta checkpoints every kCkRows rows in shared memory, an 8-row chunk of beta in registers ----
constexpr int      kCkRows    = 16;
constexpr int      kMaxCk     = kMaxL / kCkRows;      // 24 checkpoints of 32 B per thread
constexpr int      kWarpSmem  = (kMaxCk + 1) * 1024;  // shared memory of one warp: its checkpoints [ck][half][lane] 128-bit
                                                      // + one scratch vector (the beta between the two units of a chunk)
// ---- general path (L % 4 != 0, tracked tier) and exact variant: staged chunks of 8 rows (TMA ring), an 8-row chunk of
// ---- beta in shared memory, checkpoints every 8 rows in HBM.  They live in the same per-warp shared memory (a warp runs
// ---- one variant at a time): [ring kStages x kStageBytes | beta chunk 8 rows x 2 x 32 lanes x 16 B]
constexpr int      kChunk     = 8;
constexpr int      kMaxChunks = 48;  // ceil(384 / 8)
constexpr int      kChkAhead  = 4;   // forward pass: checkpoints are prefetched into L2 this many chunks ahead
constexpr int      kStages    = 3;   // chunks in flight per warp (TMA bulk copies, one mbarrier per stage)
constexpr int      kStageBytes = 3072;  // [sys | par | A or E] x 1 KB: 8 rows of all the code blocks of a warp
constexpr int      kSmLanes   = 32;     // stride (in 128-bit words) between the two halves / the rows of the beta chunk
static_assert(kStages * kStageBytes + kChunk * 2 * 32 * 16 <= kWarpSmem, "general path must fit the warp's shared memory");
// per-warp-slot workspace strides are not powers of two: warps run in near lock step, and power-of-two
// strides would send all of them to the same L2 slices / HBM channels at once
constexpr size_t   kXArrayBytes16 = (size_t)(kMaxL + 1) * 128;  // one A or E array: rows of 32 words (all blocks of a warp)
constexpr size_t   kXArrayBytes8  = (size_t)(100 + 1) * 128;    // W = 8: K <= 800, L <= 100
constexpr size_t   kChkSlotBytes  = (size_t)kMaxChunks * 1024 + 128;
constexpr int      kStagePad  = 4;   // int16 of padding per window in to_internal_kernel's staged copy
constexpr int      kExactRows = 4;   // rows next to a known-state boundary always use exact arithmetic
constexpr int      kPureFastG = 2529;   // 10000 + 9 * G <= 32768: fast arithmetic is exact even next to the known start state
constexpr int      kStaticFastG = 2978; // 11 * G <= 32767: the fast variant needs no bookkeeping at all
constexpr int      kMaxFastG  = 5461; // largest per-step metric change the fast variant accepts (6 * G <= 32767)
#ifndef B200_PF
#define B200_PF 3  // development probe: bit 0 = L2 prefetch in the backward pass, 1 = forward pass, 2 = next half iteration
#endif
constexpr int      kNegInf    = -10000;
constexpr uint32_t kNegInf2   = 0xD8F0D8F0u;  // (-10000, -10000)
constexpr uint32_t kMax2      = 0x7FFF7FFFu;
constexpr uint32_t kMin2      = 0x80008000u;

// a page of zeros with the geometry of a stream: what a half iteration reads where it has no input (DEC2 has no
// separate systematic stream, the first half iteration has no a-priori values) -- keeps the loops branch free
__device__ uint4 g_zero_page[(kMaxGroups + 2) * 32];

__constant__ uint32_t c_crc_tab[2][256];

// ---- packed int16x2 arithmetic -------------------------------------------------------------------
__device__ __forceinline__ uint32_t sadd2(uint32_t a, uint32_t b) { return __vaddss2(a, b); }
__device__ __forceinline__ uint32_t ssub2(uint32_t a, uint32_t b) { return __vsubss2(a, b); }
__device__ __forceinline__ uint32_t wadd2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
__device__ __forceinline__ uint32_t wsub2(uint32_t a, uint32_t b) { return __vsub2(a, b); }
__device__ __forceinline__ uint32_t wneg2(uint32_t a) { return __vneg2(a); }
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
__device__ __forceinline__ uint32_t min2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
__device__ __forceinline__ uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
__device__ __forceinline__ uint32_t min3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_s16x2(a, b, c); }
// max(a + b, c) with a wrapping add: only used where a + b is proven not to overflow
__device__ __forceinline__ uint32_t addmax2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
// prmt with the sign-replication bit of the selector nibbles set: every result byte is 0xFF / 0x00 by the top bit of the
// selected byte (0x9999: byte 1, the sign of the low half; 0xBBBB: byte 3).  (__byte_perm masks that selector bit off.)
template <uint32_t SEL>
__device__ __forceinline__ uint32_t sign_mask(uint32_t v)
{
  uint32_t r;
  asm("prmt.b32 %0, %1, %1, %2;" : "=r"(r) : "r"(v), "n"(SEL));
  return r;
}
__device__ __forceinline__ uint32_t sra1_2(uint32_t v) { return ((v >> 1) & 0x7FFF7FFFu) | (v & 0x80008000u); }

// extremes of the values a thread has seen, per int16 lane
struct Range {
  uint32_t hi, lo;
  __device__ __forceinline__ void reset() { hi = 0; lo = 0; }
  __device__ __forceinline__ void add8(const uint32_t s[8])
  {  // s[0] is 0 after a normalisation and hi/lo already straddle 0
    const uint32_t h1 = max3(s[1], s[2], s[3]), h2 = max3(s[4], s[5], s[6]);
    const uint32_t l1 = min3(s[1], s[2], s[3]), l2 = min3(s[4], s[5], s[6]);
    hi = max3(hi, h1, h2);
    lo = min3(lo, l1, l2);
    hi = max2(hi, s[7]);
    lo = min2(lo, s[7]);
  }
  __device__ __forceinline__ void add8full(const uint32_t s[8])
  {  // a state vector that has not just been normalised: s[0] counts too
    add8(s);
    add1(s[0]);
  }
  __device__ __forceinline__ void add2v(uint32_t a, uint32_t b)
  {
    hi = max3(hi, a, b);
    lo = min3(lo, a, b);
  }
  __device__ __forceinline__ void add1(uint32_t a)
  {
    hi = max2(hi, a);
    lo = min2(lo, a);
  }
};

__device__ __forceinline__ int lo16(uint32_t v) { return (int)(int16_t)(v & 0xFFFFu); }
__device__ __forceinline__ int hi16(uint32_t v) { return (int)(int16_t)(v >> 16); }

// ---- trellis steps ---------------------------------------------------------------------------------
// FAST = wrapping adds fused with the max; !FAST = saturating adds exactly as the reference.
template <bool FAST>
__device__ __forceinline__ void beta_step(uint32_t b[8], uint32_t x, uint32_t y, uint32_t xy)
{
  uint32_t n0, n1, n2, n3, n4, n5, n6, n7;
  if (FAST) {
    n0 = addmax2(b[4], xy, b[0]);
    n1 = addmax2(b[0], xy, b[4]);
    n2 = addmax2(b[5], y, wadd2(b[1], x));
    n3 = addmax2(b[5], x, wadd2(b[1], y));
    n4 = addmax2(b[6], x, wadd2(b[2], y));
    n5 = addmax2(b[6], y, wadd2(b[2], x));
    n6 = addmax2(b[3], xy, b[7]);
    n7 = addmax2(b[7], xy, b[3]);
  } else {
    n0 = max2(sadd2(b[4], xy), b[0]);
    n1 = max2(b[4], sadd2(b[0], xy));
    n2 = max2(sadd2(b[5], y), sadd2(b[1], x));
    n3 = max2(sadd2(b[5], x), sadd2(b[1], y));
    n4 = max2(sadd2(b[6], x), sadd2(b[2], y));
    n5 = max2(sadd2(b[6], y), sadd2(b[2], x));
    n6 = max2(b[7], sadd2(b[3], xy));
    n7 = max2(sadd2(b[7], xy), b[3]);
  }
  b[0] = n0; b[1] = n1; b[2] = n2; b[3] = n3; b[4] = n4; b[5] = n5; b[6] = n6; b[7] = n7;
}

// alpha recursion without output (warm-up rows)
template <bool FAST>
__device__ __forceinline__ void alpha_step(uint32_t a[8], uint32_t x, uint32_t y, uint32_t xy)
{
  uint32_t n0, n1, n2, n3, n4, n5, n6, n7;
  if (FAST) {
    n0 = addmax2(a[1], xy, a[0]);
    n1 = addmax2(a[2], x, wadd2(a[3], y));
    n2 = addmax2(a[5], x, wadd2(a[4], y));
    n3 = addmax2(a[6], xy, a[7]);
    n4 = addmax2(a[0], xy, a[1]);
    n5 = addmax2(a[3], x, wadd2(a[2], y));
    n6 = addmax2(a[4], x, wadd2(a[5], y));
    n7 = addmax2(a[7], xy, a[6]);
  } else {
    n0 = max2(a[0], sadd2(a[1], xy));
    n1 = max2(sadd2(a[3], y), sadd2(a[2], x));
    n2 = max2(sadd2(a[4], y), sadd2(a[5], x));
    n3 = max2(a[7], sadd2(a[6], xy));
    n4 = max2(a[1], sadd2(a[0], xy));
    n5 = max2(sadd2(a[2], y), sadd2(a[3], x));
    n6 = max2(sadd2(a[5], y), sadd2(a[4], x));
    n7 = max2(a[6], sadd2(a[7], xy));
  }
  a[0] = n0; a[1] = n1; a[2] = n2; a[3] = n3; a[4] = n4; a[5] = n5; a[6] = n6; a[7] = n7;
}

// alpha recursion + a-posteriori output (max over bit-1 branches minus max over bit-0 branches)
// aux: FAST only -- returns out - aux (the differenced value that is stored), saving one packed subtraction
template <bool FAST, bool TRACK = true>
__device__ __forceinline__ uint32_t alpha_out_step(uint32_t a[8], const uint32_t bb[8], uint32_t x, uint32_t y,
                                                   uint32_t xy, Range& rm, uint32_t aux = 0)
{
  uint32_t m[8], n[8], M0, M1, o;
  if (FAST) {
    m[0] = a[0];            m[1] = wadd2(a[3], y);  m[2] = wadd2(a[4], y);  m[3] = a[7];
    m[4] = a[1];            m[5] = wadd2(a[2], y);  m[6] = wadd2(a[5], y);  m[7] = a[6];
    n[0] = wadd2(a[1], xy); n[1] = wadd2(a[2], x);  n[2] = wadd2(a[5], x);  n[3] = wadd2(a[6], xy);
    n[4] = wadd2(a[0], xy); n[5] = wadd2(a[3], x);  n[6] = wadd2(a[4], x);  n[7] = wadd2(a[7], xy);
    M0 = wadd2(bb[0], m[0]);
    M1 = wadd2(bb[0], n[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) {
      M0 = addmax2(bb[i], m[i], M0);
      M1 = addmax2(bb[i], n[i], M1);
    }
    if (TRACK) rm.add2v(M0, M1);
    o = wsub2(M1, wadd2(M0, aux));
  } else {
    m[0] = a[0];            m[1] = sadd2(a[3], y);  m[2] = sadd2(a[4], y);  m[3] = a[7];
    m[4] = a[1];            m[5] = sadd2(a[2], y);  m[6] = sadd2(a[5], y);  m[7] = a[6];
    n[0] = sadd2(a[1], xy); n[1] = sadd2(a[2], x);  n[2] = sadd2(a[5], x);  n[3] = sadd2(a[6], xy);
    n[4] = sadd2(a[0], xy); n[5] = sadd2(a[3], x);  n[6] = sadd2(a[4], x);  n[7] = sadd2(a[7], xy);
    M0 = sadd2(bb[0], m[0]);
    M1 = sadd2(bb[0], n[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) {
      M0 = max2(M0, sadd2(bb[i], m[i]));
      M1 = max2(M1, sadd2(bb[i], n[i]));
    }
    o = ssub2(M1, M0);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = max2(m[i], n[i]);
  return o;
}

template <bool FAST>
__device__ __forceinline__ void normalize(uint32_t s[8])
{
  if (FAST) {
    const uint32_t neg = wneg2(s[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) s[i] = wadd2(s[i], neg);
  } else {
#pragma unroll
    for (int i = 1; i < 8; i++) s[i] = ssub2(s[i], s[0]);
  }
  s[0] = 0;
}

// ---- per-thread decode context ---------------------------------------------------------------------
// A work item is up to 32/(W/2) code blocks of equal K decoded by one warp.  Its input streams are laid out
// [row group of 4][lane of the warp][4 rows] 32-bit words (two windows each), 512 bytes per row group: lane = block *
// W/2 + thread, so the words a thread needs for 4 consecutive trellis rows are one 128-bit load, and a warp reads
// whole 512-byte pieces.  to_internal_kernel writes them that way (a partial item leaves the lanes of its missing
// blocks unused).  The two extrinsic arrays are [row][lane] words: the 16 windows of a row of a code block are one
// 32-byte sector, which the producer's scatter writes completely at once (group-major arrays were tried: their
// sectors fill up in four visits and DRAM write traffic doubled).
template <int W>
struct WinCtx {
  uint32_t K, L;
  int      t;         // 0..W/2-1: owns windows 2t (low half) and 2t+1 (high half)
  int      lane, grp;
  uint32_t sp_off;    // byte offset of this thread's 16 bytes inside a row group of the INPUT streams (the lanes of a
                      // missing block of a partial item shadow block 0)
  uint32_t s_bytes;   // bytes per input stream of the item (ngroups * 512)
  uint32_t ngroups;   // row groups per stream (L rounded up to 4, / 4)
  bool     idle;      // tdec_win_dyn_kernel: the group has no block in this half iteration: its loads go to the zero page
  bool     noap;      // first half iteration of a block: there is no a-priori information yet (A reads as zero and is
                      // neither fetched nor, unless a hard decision can come before DEC2 has written it, cleared)
  const char*     in_item;  // the item's sys | par0 | par1 streams
  const int16_t*  tail;     // this block's 12 tail samples
  uint32_t*       A32;  // [row][32 lanes] words: extrinsic of DEC2 minus E, natural order (= a-priori of DEC1)
  uint32_t*       E32;  // [row][32 lanes] words: a-posteriori of DEC1 minus A, in DEC2's interleaved order
                        // (each array is written scattered by its producer and read linearly by its consumer)
  const uint32_t* R;    // CRC modes: per-bit CRC contributions of this half iteration's trellis positions,
                        // [row][window] (lte_tables.h:crc_pos_tables), offset to this thread's window pair;
                        // nullptr when no block of the warp checks a CRC
  uint4*          ck;   // main fast path: the warp's checkpoints in shared memory [ck][half][32 lanes], offset by lane
  uint4*          chk;  // general / exact path: beta checkpoints in HBM [chunk][half][32 lanes], already offset by lane
  uint4*          sm;   // general / exact path: chunk of beta in the warp's shared memory [(s*2 + h)*kSmLanes], offset by lane
  char*           stages;  // general path: the warp's kStages staging buffers: [sys 1 KB | par 1 KB | A or E rows 1 KB]
  uint64_t*       mbar;    // the warp's kStages mbarriers
  const uint32_t* stab;    // [2][kMaxGroups][W/2][4] scatter table, offset by t*4: where the differenced outputs of row k's
                           // window pair go, as byte offsets inside the block's part of the array (low / high 16 bits =
                           // window 2t / 2t+1): dir 0 = pi (DEC2 writes A), 1 = pi^-1 (DEC1 writes E)
};
constexpr uint32_t kStabDir = kMaxGroups * 8 * 4;  // 32-bit entries per direction (sized for W = 16)

// 32-bit word of (row k, lane) in an extrinsic array
__device__ __forceinline__ uint32_t ae_word(uint32_t k, uint32_t lane) { return k * 32u + lane; }
// scatter-table entry of row k for a thread whose stab pointer is already offset by t*4
template <int W>
__device__ __forceinline__ uint32_t stab_at(const uint32_t* stab, uint32_t dir, uint32_t k)
{
  return stab[dir * kStabDir + (k >> 2) * (W / 2 * 4) + (k & 3u)];
}

// the inputs of 4 consecutive trellis rows (one row group) for this thread's window pair
struct Group {
  uint32_t x[4], y[4], aux[4];
};

// ---- TMA bulk copies + mbarriers: the inputs of the next kStages-1 chunks are always in flight ---------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t b, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t b, uint32_t parity)
{
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(b), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
// generic-proxy writes (st.global / st.shared) before, async-proxy (TMA) accesses after
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void pf_l2_dyn(const char* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// The chunks a half iteration consumes, in order (depends only on L):
//   phase 0: 4..0        beta over rows 39..0 (boundary metrics for the previous window)
//   phase 1: ctop..0     beta over the window, checkpoint every kChunk rows
//   phase 2: a0/8..ctop  alpha over rows L-40..L-1 (boundary metrics for the next window)
//   phase 3: 0..ctop     beta rebuild + alpha + output, chunk by chunk
struct Pipe {
  uint32_t phase;
  int      ch;    // next chunk to fetch
  uint32_t slot;  // next stage to consume; it is refilled right after (the ring is always full)
  uint32_t par;   // phase parity of every stage's mbarrier (survives across half iterations)
};

template <int W>
__device__ __forceinline__ void pipe_issue(const WinCtx<W>& c, bool dec2, Pipe& p, uint32_t slot)
{
  if (p.phase > 3) return;
  const int ctop = (int)(c.L - 1) >> 3;
  {
    // lane i prepares and issues copy i (DEC1: sys, par0, A; DEC2: par1, E): the operands are computed once, SIMD,
    // and the bulk-copy instruction -- which takes uniform registers -- runs once per active lane
    const uint32_t ch     = (uint32_t)p.ch;
    const uint32_t gbytes = min(2u, c.ngroups - 2 * ch) * kGroupBytes;
    const uint32_t xbytes = min(8u, c.L - 8 * ch) * 128u;
    const uint32_t mb     = smem_u32(c.mbar + slot);
    const uint32_t nc     = (dec2 || c.noap) ? 2u : 3u;
    const uint32_t i      = (uint32_t)c.lane;
    if (i == 0) mbar_expect_tx(mb, (dec2 ? gbytes : 2 * gbytes) + (c.noap ? 0u : xbytes));
    if (i < nc) {
      const bool     isx    = i == (dec2 ? 1u : 2u);
      const uint32_t stream = dec2 ? 2u : i;
      const char*    src    = isx ? reinterpret_cast<const char*>(dec2 ? c.E32 : c.A32) + (size_t)ch * 1024
                                  : c.in_item + (size_t)ch * 2 * kGroupBytes + (size_t)stream * c.s_bytes;
      const uint32_t dst    = smem_u32(c.stages + slot * kStageBytes) + (isx ? 2048u : dec2 ? 1024u : i * 1024u);
      bulk_g2s(dst, src, isx ? xbytes : gbytes, mb);
    }
  }
  if (p.phase == 0 || p.phase == 1) {
    if (p.ch == 0) {
      p.ch = p.phase == 0 ? ctop : (int)(c.L - kWarm) >> 3;
      p.phase++;
    } else {
      p.ch--;
    }
  } else {
    if (p.ch == ctop) {
      p.ch = 0;
      p.phase++;
    } else {
      p.ch++;
    }
  }
}

template <int W>
__device__ __forceinline__ void pipe_start(const WinCtx<W>& c, bool dec2, Pipe& p)
{
  p.phase = 0;
  p.ch    = kWarm / kChunk - 1;
  p.slot  = 0;
#pragma unroll
  for (uint32_t s = 0; s < (uint32_t)kStages; s++) pipe_issue<W>(c, dec2, p, s);
}

// wait for the oldest chunk in flight; returns its staging buffer
template <int W>
__device__ __forceinline__ const char* pipe_wait(const WinCtx<W>& c, Pipe& p)
{
  const uint32_t mb = smem_u32(c.mbar + p.slot), parity = (p.par >> p.slot) & 1u;
  uint32_t       spins = 0;
  while (!mbar_try_wait(mb, parity)) {
    if (++spins > (1u << 24)) __trap();  // a lost copy must not hang the GPU
  }
  p.par ^= 1u << p.slot;
  return c.stages + p.slot * kStageBytes;
}

// every lane is done reading the stage: refill it with the next chunk of the sequence
template <int W>
__device__ __forceinline__ void pipe_release(const WinCtx<W>& c, bool dec2, Pipe& p)
{
  __syncwarp();
  pipe_issue<W>(c, dec2, p, p.slot);
  p.slot = p.slot + 1 == (uint32_t)kStages ? 0u : p.slot + 1;
}

// FAST rows only: row group g (0 / 1) of a staged chunk; x = sys + a-priori with a wrapping add.  DEC2 has no
// separate systematic input (x = E): its copies leave the sys part of the stages alone and the kernel clears it
// before every DEC2 half iteration, so both decoders run the same branch-free code.
template <int W>
__device__ __forceinline__ void load_group(const WinCtx<W>& c, bool dec2, const char* st, int g, Group& q)
{
  const uint4     ys = *reinterpret_cast<const uint4*>(st + 1024 + g * kGroupBytes + c.sp_off);
  const uint32_t* xr = reinterpret_cast<const uint32_t*>(st + 2048 + g * 512) + c.lane;
  const uint4     xs = *reinterpret_cast<const uint4*>(st + g * kGroupBytes + c.sp_off);
  q.y[0] = ys.x; q.y[1] = ys.y; q.y[2] = ys.z; q.y[3] = ys.w;
#pragma unroll
  for (int r = 0; r < 4; r++) q.aux[r] = xr[r * 32];
  q.x[0] = wadd2(q.aux[0], xs.x); q.x[1] = wadd2(q.aux[1], xs.y);
  q.x[2] = wadd2(q.aux[2], xs.z); q.x[3] = wadd2(q.aux[3], xs.w);
}

template <int W>
__device__ __forceinline__ void chk_store(const WinCtx<W>& c, int ch, const uint32_t s[8])
{
  c.chk[(ch * 2 + 0) * 32] = make_uint4(s[0], s[1], s[2], s[3]);
  c.chk[(ch * 2 + 1) * 32] = make_uint4(s[4], s[5], s[6], s[7]);
}
// checkpoints come back from HBM in the forward pass: pull them into L2 a few chunks before they are loaded
template <int W>
__device__ __forceinline__ void chk_prefetch_l2(const WinCtx<W>& c, int ch)
{
  asm volatile("prefetch.global.L2 [%0];" ::"l"(c.chk + (ch * 2 + 0) * 32));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(c.chk + (ch * 2 + 1) * 32));
}
template <int W>
__device__ __forceinline__ void chk_load(const WinCtx<W>& c, int ch, uint32_t s[8])
{
  const uint4 v0 = c.chk[(ch * 2 + 0) * 32], v1 = c.chk[(ch * 2 + 1) * 32];
  s[0] = v0.x; s[1] = v0.y; s[2] = v0.z; s[3] = v0.w;
  s[4] = v1.x; s[5] = v1.y; s[6] = v1.z; s[7] = v1.w;
}

// one row for the exact helpers: plain loads issued one row ahead, the reference's saturating a-priori add
struct RawRow {
  uint32_t a, b, c;
};

template <int W>
__device__ __forceinline__ void issue_row(const WinCtx<W>& c, bool dec2, uint32_t k, RawRow& q)
{
  const char* gp = c.in_item + (size_t)(k >> 2) * kGroupBytes + c.sp_off + (k & 3) * 4;
  if (!dec2) {
    q.a = __ldg(reinterpret_cast<const uint32_t*>(gp));
    q.b = __ldg(reinterpret_cast<const uint32_t*>(gp + c.s_bytes));
    q.c = c.noap ? 0u : c.A32[ae_word(k, c.lane)];
  } else {
    q.a = 0;
    q.b = __ldg(reinterpret_cast<const uint32_t*>(gp + 2 * (size_t)c.s_bytes));
    q.c = c.E32[ae_word(k, c.lane)];
  }
}

__device__ __forceinline__ void finish_row_exact(bool dec2, const RawRow& q, uint32_t& x, uint32_t& y, uint32_t& aux)
{
  y   = q.b;
  aux = q.c;
  x   = dec2 ? aux : sadd2(aux, q.a);  // the a-priori array is all zero in the first half iteration
}

// store the differenced output of row k (see file header) where its consumer will read it linearly, and
// remember its extremes.  DEC1 (natural position) -> E in DEC2's order: pi^-1;  DEC2 -> A in natural order: pi.
// The QPP is contention free: all windows of row k go to ONE destination row, permuted among the windows.
// HARD (CRC modes): the hard decisions (out = d + aux > 0) of the two windows are folded into the running CRC of
// the block: the CRC is linear, every set bit xors in its precomputed contribution, in whatever order they come.
// this code block's 32 bytes of row 0 of the array the half iteration writes
template <int W>
__device__ __forceinline__ char* out_base(const WinCtx<W>& c, bool dec2)
{
  return reinterpret_cast<char*>(dec2 ? c.A32 : c.E32) + c.grp * (2 * W);
}

template <int W, bool HARD>
__device__ __forceinline__ void store_diff(const WinCtx<W>& c, bool dec2, uint32_t k, uint32_t d, uint32_t aux, Range& rd,
                                           uint32_t& crc, char* Y = nullptr)
{
  rd.add1(d);
  if (HARD) {
    // sign bit of -(max(v, -1)) is set exactly when v > 0 (no overflow: max(v,-1) >= -1)
    const uint32_t m  = wneg2(max2(wadd2(d, aux), 0xFFFFFFFFu));
    const uint2    rr = *reinterpret_cast<const uint2*>(c.R + k * W);
    crc ^= (rr.x & (uint32_t)((int32_t)(m << 16) >> 31)) ^ (rr.y & (uint32_t)((int32_t)m >> 31));
  }
  const uint32_t e = stab_at<W>(c.stab, dec2 ? 0u : 1u, k);
  if (!Y) Y = out_base<W>(c, dec2);
  *reinterpret_cast<uint16_t*>(Y + (e & 0xFFFFu)) = (uint16_t)(d & 0xFFFFu);
  *reinterpret_cast<uint16_t*>(Y + (e >> 16))     = (uint16_t)(d >> 16);
}
template <int W>
__device__ __forceinline__ void store_out(const WinCtx<W>& c, bool dec2, uint32_t k, uint32_t o, uint32_t aux, Range& rd,
                                          uint32_t& crc)
{
  if (c.R)
    store_diff<W, true>(c, dec2, k, wsub2(o, aux), aux, rd, crc);
  else
    store_diff<W, false>(c, dec2, k, wsub2(o, aux), aux, rd, crc);
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

// state of one half iteration that the out-of-line exact helpers update (kept in local memory only
// around the helper calls; the fast loops work on register copies)
struct RowState {
  uint32_t s[8];       // beta or alpha metrics
  Range    trk;        // post-normalisation extremes
  Range    rd;         // extremes of the stored outputs
  uint32_t crc;        // running CRC of the hard decisions (CRC modes)
};

// ---- exact rows (reference arithmetic), one row at a time, shared by the exact variant and by the
// ---- rows of the fast variant that sit next to a known-state boundary -------------------------------
// beta recursion over rows k_hi .. k_lo (descending).  mode 0: warm-up (nothing stored), 1: main pass
// (checkpoint when k % kChunk == 0), 2: rebuild (B[k] -> shared memory slot k - sm_lo - 1).
template <int W>
__device__ __noinline__ void beta_rows_exact(const WinCtx<W> c, bool dec2, int k_hi, int k_lo, int mode, int sm_lo,
                                             RowState* st)
{
  uint32_t   s[8];
#pragma unroll
  for (int i = 0; i < 8; i++) s[i] = st->s[i];
  Range trk = st->trk;
  RawRow q;
  if (k_hi >= k_lo) issue_row<W>(c, dec2, (uint32_t)k_hi, q);
#pragma unroll 1
  for (int k = k_hi; k >= k_lo; k--) {
    uint32_t x, y, aux;
    finish_row_exact(dec2, q, x, y, aux);
    if (k > k_lo) issue_row<W>(c, dec2, (uint32_t)(k - 1), q);
    beta_step<false>(s, x, y, sadd2(x, y));
    if (mode == 1 && (k % kChunk) == 0 && k != 0) chk_store<W>(c, k / kChunk - 1, s);
    if (mode == 2) {
      c.sm[((k - sm_lo - 1) * 2 + 0) * kSmLanes] = make_uint4(s[0], s[1], s[2], s[3]);
      c.sm[((k - sm_lo - 1) * 2 + 1) * kSmLanes] = make_uint4(s[4], s[5], s[6], s[7]);
    }
    if ((k & 1) == 0 && k != 0) {
      normalize<false>(s);
      trk.add8(s);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) st->s[i] = s[i];
  st->trk = trk;
}

// alpha recursion over rows k_lo .. k_hi (ascending).  mode 0: warm-up (normalisation counter starts at
// k_norm0 for row k_lo, no output), 1: with a-posteriori output from the beta chunk whose base row is sm_lo.
template <int W>
__device__ __noinline__ void alpha_rows_exact(const WinCtx<W> c, bool dec2, int k_lo, int k_hi, int mode, int sm_lo,
                                              int k_norm0, RowState* st)
{
  uint32_t   a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = st->s[i];
  Range trk = st->trk, rd = st->rd, unused;
  uint32_t crc = st->crc;
  unused.reset();
  RawRow q;
  if (k_lo <= k_hi) issue_row<W>(c, dec2, (uint32_t)k_lo, q);
#pragma unroll 1
  for (int k = k_lo; k <= k_hi; k++) {
    uint32_t x, y, aux;
    finish_row_exact(dec2, q, x, y, aux);
    if (k < k_hi) issue_row<W>(c, dec2, (uint32_t)(k + 1), q);
    const int j = mode == 0 ? k_norm0 + (k - k_lo) : k;
    if (mode == 0) {
      alpha_step<false>(a, x, y, sadd2(x, y));
    } else {
      const uint4    b0 = c.sm[((k - sm_lo) * 2 + 0) * kSmLanes];
      const uint4    b1 = c.sm[((k - sm_lo) * 2 + 1) * kSmLanes];
      const uint32_t bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint32_t       o = alpha_out_step<false>(a, bb, x, y, sadd2(x, y), unused);
      if (W == 8) o = sra1_2(o);  // the 8-window (sse16) decoder halves its output
      store_out<W>(c, dec2, (uint32_t)k, o, aux, rd, crc);
    }
    if ((j & 1) == 0 && j != 0) {
      normalize<false>(a);
      trk.add8(a);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) st->s[i] = a[i];
  st->trk = trk;
  st->rd  = rd;
  st->crc = crc;
}

// what one half iteration reports back
struct HalfResult {
  bool     proven;  // fast variant only: no saturating op of the reference can have clamped
  uint32_t dmax;    // max |stored differenced output| over this thread's rows and lanes
  uint32_t crc;     // CRC modes: xor of the CRC contributions of this thread's set bits
};

template <int WH>
__device__ __forceinline__ void exchange_beta_boundary(uint32_t s[8], int t, const int16_t* tail)
{
  // window d starts from what window d+1 estimated; the last window from the terminated tail
  int16_t tb[8];
  tail_beta(tail, tb);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t nxt = __shfl_down_sync(0xFFFFFFFFu, s[i], 1, WH);
    uint32_t       v   = __byte_perm(s[i], nxt, 0x5432);
    if (t == WH - 1) v = (v & 0xFFFFu) | ((uint32_t)(uint16_t)tb[i] << 16);
    s[i] = v;
  }
}

template <int WH>
__device__ __forceinline__ void exchange_alpha_boundary(uint32_t a[8], int t)
{
  // window d starts from what window d-1 estimated; window 0 from the known all-zero state
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t prv = __shfl_up_sync(0xFFFFFFFFu, a[i], 1, WH);
    uint32_t       v   = __byte_perm(a[i], prv, 0x1076);
    if (t == 0) v = (v & 0xFFFF0000u) | (i == 0 ? 0u : (uint32_t)(uint16_t)kNegInf);
    a[i] = v;
  }
}

__device__ __forceinline__ uint32_t range_absmax(const Range& r)
{
  const int d0 = max(abs(lo16(r.hi)), abs(lo16(r.lo))), d1 = max(abs(hi16(r.hi)), abs(hi16(r.lo)));
  return (uint32_t)max(d0, d1);
}

// ---- EXACT variant of one half iteration: the reference's saturating arithmetic on every row -------
template <int W>
__device__ __noinline__ HalfResult half_iteration_exact(const WinCtx<W> c, bool dec2)
{
  constexpr int  WH = W / 2;
  const int      L  = (int)c.L;
  const int      nchunks = (L + kChunk - 1) / kChunk;
  RowState       st;
  st.trk.reset();
  st.rd.reset();
#pragma unroll
  for (int i = 0; i < 8; i++) st.s[i] = kNegInf2;
  beta_rows_exact<W>(c, dec2, kWarm - 1, 0, 0, 0, &st);
  exchange_beta_boundary<WH>(st.s, c.t, c.tail + (dec2 ? 6 : 0));
  chk_store<W>(c, nchunks - 1, st.s);
  beta_rows_exact<W>(c, dec2, L - 1, 0, 1, 0, &st);

  RowState al;
  al.trk.reset();
  al.rd.reset();
  al.crc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) al.s[i] = kNegInf2;
  alpha_rows_exact<W>(c, dec2, L - kWarm, L - 1, 0, 0, 0, &al);
  exchange_alpha_boundary<WH>(al.s, c.t);
  for (int ch = 0; ch < nchunks; ch++) {
    const int lo = ch * kChunk, hi = min(lo + kChunk, L);
    chk_load<W>(c, ch, st.s);
    c.sm[((hi - lo - 1) * 2 + 0) * kSmLanes] = make_uint4(st.s[0], st.s[1], st.s[2], st.s[3]);
    c.sm[((hi - lo - 1) * 2 + 1) * kSmLanes] = make_uint4(st.s[4], st.s[5], st.s[6], st.s[7]);
    if (hi != L) normalize<false>(st.s);  // hi is even and non-zero: the backward pass normalised after storing
    beta_rows_exact<W>(c, dec2, hi - 1, lo + 1, 2, lo, &st);
    alpha_rows_exact<W>(c, dec2, lo, hi - 1, 1, lo, 0, &al);
  }
  __syncwarp();
  HalfResult res;
  res.proven = true;
  res.dmax   = range_absmax(al.rd);
  res.crc    = al.crc;
  return res;
}

// ---- FAST variant: wrapping adds fused with max, plus (TRACK) the bookkeeping that proves it equals the exact one
// G bounds |x|, |y| and |x + y| of every row of this code block in this half iteration.
// DEC1 and DEC2 share one instantiation (they differ only in which streams a chunk holds): in steady state the
// 12 warps of an SM run the same code, and the hot loops (backward 4 rows, rebuild + forward 8 rows) fit the
// instruction cache.
// Rows next to the terminated tail need no special treatment in any tier: the tail samples are part of G
// (to_internal_kernel), so the last window's start metrics -- 3 tail steps from the known end state, every state
// reached -- already lie in [-3G, 3G] like any other beta (the tracked tier adds that boundary vector to its
// bookkeeping).  Only the rows above the last full row group (L % 4 of them) go through the exact helpers, because
// the fast loops work on whole groups.
// EDGE = true: rows 0..3 of the forward pass, next to the known start state (0, -10000 x 7) of the first window,
// use exact arithmetic.  EDGE = false (G <= kPureFastG): not even those.  Proof: for rows 0..2 alpha + branch lies
// in [-10000 - 3G, 3G], plus beta in [-5G, 5G] that is [-10000 - 8G, 8G]: inside int16 for G <= kPureFastG; from
// row 3 on alpha has gone through 3 steps and a normalisation and is in general position.  |out| <= 5G there too:
// the best bit-1 and bit-0 candidates can be chosen from the same predecessor state, so they differ by at most
// spread(beta) + 2G.
// HARD (CRC modes): the forward pass also accumulates the CRC of the hard decisions (c.R).
// FULL: L % 8 == 0, every chunk holds two complete row groups (K = 6144, 4096, 2048 ...): the partial-chunk tests
// and the exact rows above the last full row group drop out of the loops at compile time.
template <int W, bool TRACK, bool EDGE, bool HARD, bool FULL>
__device__ __forceinline__ HalfResult half_iteration_fast(const WinCtx<W>& c, bool dec2, int G, Pipe& p)
{
  constexpr int WH = W / 2;
  const int     L  = (int)c.L;
  const int     ctop = (L - 1) >> 3;
  // rows [0, kf] are covered by full row groups: fast arithmetic (except, EDGE, rows 0..3 of the forward pass);
  // the L % 4 rows above kf go through the exact helpers
  const int kf = (L & ~3) - 1;
  const int a0 = L - kWarm;  // first row of the alpha warm-up
  uint32_t  s[8];
  Range     rb, ra, rm, rd;
  rb.reset(); ra.reset(); rd.reset();
  rm.hi = kMin2; rm.lo = kMax2;
  RowState st;
  uint32_t crc = 0;
  st.crc = 0;
  pipe_start<W>(c, dec2, p);

  // ---------------- backward pass ----------------
  // phase 0: boundary metrics from the next window's first 40 rows; phase 1: the window itself with a
  // checkpoint every kChunk rows
#pragma unroll
  for (int i = 0; i < 8; i++) s[i] = kNegInf2;
#pragma unroll 1
  for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
    for (int ch = phase ? ctop : kWarm / kChunk - 1; ch >= 0; ch--) {
      const char* stg = pipe_wait<W>(c, p);
#pragma unroll
      for (int g = 1; g >= 0; g--) {
        const int k0 = ch * kChunk + g * 4;
        if (!FULL && k0 > kf) continue;
        Group q;
        load_group<W>(c, dec2, stg, g, q);
#pragma unroll
        for (int r = 3; r >= 0; r--) {
          beta_step<true>(s, q.x[r], q.y[r], wadd2(q.x[r], q.y[r]));
          if (r == 0 && g == 0 && phase && ch != 0) chk_store<W>(c, ch - 1, s);
          if ((r & 1) == 0 && (k0 | r) != 0) {
            normalize<true>(s);
            if (TRACK) rb.add8(s);
          }
        }
      }
      pipe_release<W>(c, dec2, p);
    }
    if (phase == 0) {
      exchange_beta_boundary<WH>(s, c.t, c.tail + (dec2 ? 6 : 0));
      chk_store<W>(c, ctop, s);
      if (TRACK) rb.add8full(s);
      if (!FULL && kf < L - 1) {  // the rows above the last full row group
#pragma unroll
        for (int i = 0; i < 8; i++) st.s[i] = s[i];
        st.trk = rb;
        beta_rows_exact<W>(c, dec2, L - 1, kf + 1, 1, 0, &st);
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = st.s[i];
        rb = st.trk;
      }
    }
  }

  // ---------------- forward pass: boundary metrics from the previous window's last 40 rows ----------------
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = kNegInf2;
#pragma unroll 1
  for (int ch = a0 >> 3; ch <= ctop; ch++) {
    const char* stg = pipe_wait<W>(c, p);
#pragma unroll 1
    for (int g = 0; g < 2; g++) {
      const int k0 = ch * kChunk + g * 4;
      if (k0 + 3 < a0 || k0 >= L) continue;
      Group q;
      load_group<W>(c, dec2, stg, g, q);
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int j = k0 + r - a0;  // the reference normalises on the warm-up counter
        if (j < 0 || k0 + r >= L) continue;
        alpha_step<true>(a, q.x[r], q.y[r], wadd2(q.x[r], q.y[r]));
        if ((j & 1) == 0 && j != 0) {
          normalize<true>(a);
          if (TRACK) ra.add8(a);
        }
      }
    }
    pipe_release<W>(c, dec2, p);
  }
  exchange_alpha_boundary<WH>(a, c.t);

  // ---------------- forward pass over the window, chunk by chunk ----------------
  char* const Yout = out_base<W>(c, dec2);
  uint32_t nxt[8];  // checkpoint of the next chunk, loaded one chunk ahead (and prefetched into L2 kChkAhead ahead)
  chk_load<W>(c, 0, nxt);
#pragma unroll
  for (int i = 1; i < kChkAhead; i++)
    if (i <= ctop) chk_prefetch_l2<W>(c, i);
#pragma unroll 1
  for (int ch = 0; ch <= ctop; ch++) {
    const int lo = ch * kChunk;
    const int hi = FULL ? lo + kChunk : min(lo + kChunk, L);
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = nxt[i];
    if (ch < ctop) chk_load<W>(c, ch + 1, nxt);
    if (ch + kChkAhead <= ctop) chk_prefetch_l2<W>(c, ch + kChkAhead);
    const char* stg = pipe_wait<W>(c, p);
    // rebuild B[lo+1 .. hi] into shared memory, slot (k - lo - 1) holds B[k]
    c.sm[((hi - lo - 1) * 2 + 0) * kSmLanes] = make_uint4(s[0], s[1], s[2], s[3]);
    c.sm[((hi - lo - 1) * 2 + 1) * kSmLanes] = make_uint4(s[4], s[5], s[6], s[7]);
    if (hi != L) {  // hi is even and non-zero: the backward pass normalised after storing
      if (hi <= kf)
        normalize<true>(s);
      else
        normalize<false>(s);
    }
    if (!FULL && hi - 1 > kf) {  // rows above the last full row group
#pragma unroll
      for (int i = 0; i < 8; i++) st.s[i] = s[i];
      st.trk.reset();
      beta_rows_exact<W>(c, dec2, hi - 1, max(kf, lo) + 1, 2, lo, &st);
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] = st.s[i];
    }
#pragma unroll
    for (int g = 1; g >= 0; g--) {
      if (!FULL && lo + g * 4 > kf) continue;
      Group q;
      load_group<W>(c, dec2, stg, g, q);
      uint4* bs = c.sm + (g * 4 - 1) * 2 * kSmLanes;
#pragma unroll
      for (int r = 3; r >= 0; r--) {
        if (r == 0 && g == 0) continue;  // B[lo] belongs to the chunk below
        beta_step<true>(s, q.x[r], q.y[r], wadd2(q.x[r], q.y[r]));
        bs[(r * 2 + 0) * kSmLanes] = make_uint4(s[0], s[1], s[2], s[3]);
        bs[(r * 2 + 1) * kSmLanes] = make_uint4(s[4], s[5], s[6], s[7]);
        if ((r & 1) == 0) normalize<true>(s);
      }
    }
    // alpha recursion + a-posteriori output over the chunk
    if (EDGE && ch == 0) {  // rows 0..3 next to the known start state: exact
#pragma unroll
      for (int i = 0; i < 8; i++) st.s[i] = a[i];
      st.trk = ra;
      st.rd  = rd;
      st.crc = crc;
      alpha_rows_exact<W>(c, dec2, 0, kExactRows - 1, 1, lo, 0, &st);
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = st.s[i];
      ra  = st.trk;
      rd  = st.rd;
      crc = st.crc;
    }
#pragma unroll 1
    for (int g = (EDGE && ch == 0) ? 1 : 0; g < 2; g++) {
      if (!FULL && lo + g * 4 > kf) continue;
      Group q;
      load_group<W>(c, dec2, stg, g, q);
      const uint4* bs = c.sm + g * 4 * 2 * kSmLanes;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const uint4    b0 = bs[(r * 2 + 0) * kSmLanes];
        const uint4    b1 = bs[(r * 2 + 1) * kSmLanes];
        const uint32_t bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        if (W == 8) {  // the 8-window (sse16) decoder halves its output before the extrinsic subtraction
          const uint32_t o = sra1_2(alpha_out_step<true, TRACK>(a, bb, q.x[r], q.y[r], wadd2(q.x[r], q.y[r]), rm));
          store_diff<W, HARD>(c, dec2, (uint32_t)(lo + g * 4 + r), wsub2(o, q.aux[r]), q.aux[r], rd, crc, Yout);
        } else {  // out - aux comes straight out of the step
          const uint32_t d = alpha_out_step<true, TRACK>(a, bb, q.x[r], q.y[r], wadd2(q.x[r], q.y[r]), rm, q.aux[r]);
          store_diff<W, HARD>(c, dec2, (uint32_t)(lo + g * 4 + r), d, q.aux[r], rd, crc, Yout);
        }
        if ((r & 1) == 0 && (EDGE || r != 0 || (lo | g) != 0)) {  // never after row 0
          normalize<true>(a);
          if (TRACK) ra.add8(a);
        }
      }
    }
    pipe_release<W>(c, dec2, p);
    if (!FULL && hi - 1 > kf) {  // rows above the last full row group
#pragma unroll
      for (int i = 0; i < 8; i++) st.s[i] = a[i];
      st.trk = ra;
      st.rd  = rd;
      st.crc = crc;
      alpha_rows_exact<W>(c, dec2, max(kf + 1, lo), hi - 1, 1, lo, 0, &st);
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = st.s[i];
      ra  = st.trk;
      rd  = st.rd;
      crc = st.crc;
    }
  }
  __syncwarp();

  HalfResult res;
  res.dmax = range_absmax(rd);
  res.crc  = crc;
  // Proof obligations, per int16 lane (see DESIGN.md "fast path"): every fast row is at most 2 steps
  // away from a normalisation point whose post-normalisation metrics lie in [lo, hi]; one step moves a
  // metric by at most G.
  bool ok = true;
  if (TRACK) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int bh = h ? hi16(rb.hi) : lo16(rb.hi), bl = h ? hi16(rb.lo) : lo16(rb.lo);
      const int ah = h ? hi16(ra.hi) : lo16(ra.hi), al = h ? hi16(ra.lo) : lo16(ra.lo);
      const int mh = h ? hi16(rm.hi) : lo16(rm.hi), ml = h ? hi16(rm.lo) : lo16(rm.lo);
      ok = ok && (bh + 2 * G <= 32767) && (bl - 2 * G >= -32768) && (bh - bl + 4 * G <= 32767);
      ok = ok && (ah + 2 * G <= 32767) && (al - 2 * G >= -32768) && (ah - al + 4 * G <= 32767);
      ok = ok && (bh + ah + 5 * G <= 32767) && (bl + al - 5 * G >= -32768);
      ok = ok && (mh < ml || mh - ml <= 32767);
    }
    // warm-up starts from 8 equal metrics of -10000 and runs at most 3 steps before normalising
    ok = ok && (G <= kMaxFastG);
  } else {
    // Static proof, no bookkeeping: 3 steps after ANY state vector the spread of the 8 metrics is at most
    // the sum over those steps of (|x| + |y|) <= 3G (every state reaches every state in 3 steps), so after a
    // normalisation all metrics lie in [-3G, 3G], before it in [-5G, 5G]; beta + alpha + branch <= 11G and the
    // output magnitude <= 7G.  G <= kStaticFastG = 32767 / 11 makes all of that fit int16.
    ok = G <= kStaticFastG;
  }
  res.proven = ok;
  return res;
}

template <int W, bool TRACK, bool EDGE, bool HARD>
__device__ __forceinline__ HalfResult half_fast(const WinCtx<W>& c, bool dec2, int G, Pipe& p)
{
  return (c.L & 7u) == 0 ? half_iteration_fast<W, TRACK, EDGE, HARD, true>(c, dec2, G, p)
                         : half_iteration_fast<W, TRACK, EDGE, HARD, false>(c, dec2, G, p);
}

// =====================================================================================================================
// MAIN FAST PATH (L % 16 == 0: K = 6144, 5888, ... 1024): no staging ring, no checkpoints in HBM.
//   * inputs come straight from HBM / L2 into registers: one 128-bit load per stream and row group, issued one to two
//     8-row units ahead of their use (plus L2 prefetches further ahead in the backward pass);
//   * the backward pass keeps a checkpoint every 16 rows in the warp's SHARED memory (24 x 32 B per thread);
//   * the forward pass works on 16-row chunks: from the checkpoint it first runs beta down the upper 8 rows (stage A),
//     then rebuilds the lower 8 rows of beta INTO REGISTERS and runs alpha + output over them (stage B), then does the
//     same for the upper 8 rows (stage C).  1.4 beta steps per row instead of 1, but the checkpoints and the rebuilt
//     chunk never leave the SM: DRAM sees the inputs, A / E and nothing else;
//   * alpha + output step with the systematic term factored out: with ay = alpha + y, the bit-1 candidates are x + n',
//     n' in {alpha, ay}, so M1 = x + max(beta + n'), alpha' = max(n' + x, m) is one fused add-max per state and the
//     stored value out - aux = (M1' + sys) - M0 needs no x + y at all: 34 packed instructions per row instead of 39;
//   * NORM = 4 (tiers without bookkeeping): the state metrics are normalised every FOUR rows instead of the reference's
//     two -- beta after the rows k = 0 mod 4, alpha after the rows k = 2 mod 4.  A normalisation subtracts the same
//     value from all eight metrics, which the outputs (differences) never see, so any schedule gives the reference's
//     bits as long as nothing wraps: after a normalisation the metrics lie in [-3G, 3G] (every state reaches every
//     state in 3 steps; the branch metrics of one step span at most |x| + |y| <= G), a step moves them by at most G,
//     and with the two schedules interleaved the alpha entering row k and the beta above it have TOGETHER taken at
//     most 4 steps since their normalisations: |alpha + branch + beta| <= (3 + 3 + 4 + 1) G = 11 G, the bound of the
//     static tier (kStaticFastG); a single recursion stays within 7 G; next to the known start state (0, -10000 x 7)
//     the sums stay above -10000 - 7 G.  The beta boundary vector (after the 40 warm-up rows) is normalised too.
//     NORM = 2 follows the reference's schedule exactly and carries the tracked tier's bookkeeping.
// =====================================================================================================================
struct Unit {  // two consecutive row groups (8 rows) of this thread's window pair
  uint4 s[2], p[2], a[2];  // systematic, parity, a-priori words
};
struct Rows4 {
  uint32_t x[4], y[4];
};

// Loads are volatile asm with a memory clobber: they stay where the source puts them relative to the stores of the
// forward pass (the compiler otherwise sinks them next to their first use, which shortens the prefetch distance).
template <int OFF>
__device__ __forceinline__ uint4 ld128(const char* p)
{
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4+%5];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "n"(OFF)
               : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ uint32_t ld32(const char* p)
{
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1+%2];" : "=r"(v) : "l"(p), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ void pf_l2(const char* p)
{
  asm volatile("prefetch.global.L2 [%0+%1];" ::"l"(p), "n"(OFF));
}

// The compiler hoists a refill above the last use of the buffer it refills; the old and the new unit then live in
// different registers and the copies it inserts at the loop's back edge wait for loads that were issued moments
// before.  An empty asm makes the addresses depend on a value the consumer produces last, so the refill stays behind
// the consumption and lands in the registers that have just become free.
__device__ __forceinline__ void pin_after(const char*& p, uint32_t dep) { asm volatile("" : "+l"(p) : "r"(dep)); }

// this thread's pointers into the three streams of a half iteration: row group g of the input streams is at
// s / p + 512 g, its four a-priori words at a + 512 g + 128 r (the extrinsic arrays are [row][lane])
struct Streams {
  const char *s, *p, *a;
  __device__ __forceinline__ Streams at(int off) const { return Streams{s + off, p + off, a + off}; }
  __device__ __forceinline__ Streams after(uint32_t dep) const
  {
    Streams r = *this;
    pin_after(r.s, dep);
    pin_after(r.p, dep);
    pin_after(r.a, dep);
    return r;
  }
};

// the two row groups at byte offsets OFF and OFF + 512
template <int OFF>
__device__ __forceinline__ void load_unit(const Streams& q, Unit& u)
{
  constexpr int GB = (int)kGroupBytes;
  u.s[0] = ld128<OFF>(q.s);
  u.p[0] = ld128<OFF>(q.p);
  u.a[0] = make_uint4(ld32<OFF>(q.a), ld32<OFF + 128>(q.a), ld32<OFF + 256>(q.a), ld32<OFF + 384>(q.a));
  u.s[1] = ld128<OFF + GB>(q.s);
  u.p[1] = ld128<OFF + GB>(q.p);
  u.a[1] = make_uint4(ld32<OFF + GB>(q.a), ld32<OFF + GB + 128>(q.a), ld32<OFF + GB + 256>(q.a), ld32<OFF + GB + 384>(q.a));
}
// the input-stream part and the a-priori part of load_unit separately
template <int OFF>
__device__ __forceinline__ void load_unit_sp(const Streams& q, Unit& u)
{
  constexpr int GB = (int)kGroupBytes;
  u.s[0] = ld128<OFF>(q.s);
  u.p[0] = ld128<OFF>(q.p);
  u.s[1] = ld128<OFF + GB>(q.s);
  u.p[1] = ld128<OFF + GB>(q.p);
}
template <int OFF>
__device__ __forceinline__ void load_unit_a(const Streams& q, Unit& u)
{
  constexpr int GB = (int)kGroupBytes;
  u.a[0] = make_uint4(ld32<OFF>(q.a), ld32<OFF + 128>(q.a), ld32<OFF + 256>(q.a), ld32<OFF + 384>(q.a));
  u.a[1] = make_uint4(ld32<OFF + GB>(q.a), ld32<OFF + GB + 128>(q.a), ld32<OFF + GB + 256>(q.a), ld32<OFF + GB + 384>(q.a));
}
// A refill whose consumer sits in the NEXT loop iteration: ptxas sees no use inside the loop body and schedules the
// loads at the very end of it, right in front of the use.  Memory instructions do not move across a warp barrier,
// so the barrier keeps them where the source has them (the lanes are converged anyway).
template <int OFF>
__device__ __forceinline__ void refill_unit(const Streams& q, Unit& u)
{
  load_unit<OFF>(q, u);
  __syncwarp();
}
// L2 prefetch of the unit at OFF: one request per 128-byte line
template <int OFF>
__device__ __forceinline__ void prefetch_unit(const Streams& q)
{
  constexpr int GB = (int)kGroupBytes;
  pf_l2<OFF>(q.s);
  pf_l2<OFF>(q.p);
  pf_l2<OFF + GB>(q.s);
  pf_l2<OFF + GB>(q.p);
#pragma unroll
  for (int r = 0; r < 8; r++) pf_l2_dyn(q.a + OFF + r * 128);
}

__device__ __forceinline__ void rows_of(const uint4& s, const uint4& p, const uint4& a, Rows4& q)
{
  q.x[0] = wadd2(a.x, s.x); q.x[1] = wadd2(a.y, s.y); q.x[2] = wadd2(a.z, s.z); q.x[3] = wadd2(a.w, s.w);
  q.y[0] = p.x; q.y[1] = p.y; q.y[2] = p.z; q.y[3] = p.w;
}

// beta over one row group, rows 4g+3 .. 4g.  skip0: no normalisation after the group's last row (row 0, NORM = 2)
template <int NORM, bool TRACK>
__device__ __forceinline__ void beta_group(uint32_t s[8], const Rows4& q, Range& rb, bool skip0)
{
#pragma unroll
  for (int r = 3; r >= 0; r--) {
    beta_step<true>(s, q.x[r], q.y[r], wadd2(q.x[r], q.y[r]));
    if (NORM == 4 ? r == 0 : (r & 1) == 0) {
      if (r != 0 || !skip0) {
        normalize<true>(s);
        if (TRACK) rb.add8(s);
      }
    }
  }
}
// beta over a unit: its upper row group, then its lower one
template <int NORM, bool TRACK>
__device__ __forceinline__ void beta_unit(uint32_t s[8], const Unit& u, Range& rb, bool skip0)
{
  Rows4 q;
  rows_of(u.s[1], u.p[1], u.a[1], q);
  beta_group<NORM, TRACK>(s, q, rb, false);
  rows_of(u.s[0], u.p[0], u.a[0], q);
  beta_group<NORM, TRACK>(s, q, rb, skip0);
}

// alpha over one row group without output (warm-up).  skip0: no normalisation after the group's first row (NORM = 2)
template <int NORM, bool TRACK>
__device__ __forceinline__ void alpha_group(uint32_t a[8], const Rows4& q, Range& ra, bool skip0)
{
#pragma unroll
  for (int r = 0; r < 4; r++) {
    alpha_step<true>(a, q.x[r], q.y[r], wadd2(q.x[r], q.y[r]));
    if (NORM == 4 ? r == 2 : (r & 1) == 0) {
      if (r != 0 || !skip0) {
        normalize<true>(a);
        if (TRACK) ra.add8(a);
      }
    }
  }
}
template <int NORM, bool TRACK>
__device__ __forceinline__ void alpha_unit(uint32_t a[8], const Unit& u, Range& ra, bool skip0)
{
  Rows4 q;
  rows_of(u.s[0], u.p[0], u.a[0], q);
  alpha_group<NORM, TRACK>(a, q, ra, skip0);
  rows_of(u.s[1], u.p[1], u.a[1], q);
  alpha_group<NORM, TRACK>(a, q, ra, false);
}

// alpha recursion + output with the systematic term factored out (see above).  Returns the a-posteriori value minus
// the a-priori value (what is stored): (M1' + sys) - M0 with sys = x - aux.
template <int W, bool TRACK>
__device__ __forceinline__ uint32_t alpha_out2(uint32_t a[8], const uint32_t bb[8], uint32_t x, uint32_t y, uint32_t sys,
                                               uint32_t aux, Range& rm)
{
  uint32_t ay[8];
#pragma unroll
  for (int i = 0; i < 8; i++) ay[i] = wadd2(a[i], y);
  const uint32_t m[8] = {a[0], ay[3], ay[4], a[7], a[1], ay[2], ay[5], a[6]};   // bit-0 branch into state i
  const uint32_t n[8] = {ay[1], a[2], a[5], ay[6], ay[0], a[3], a[4], ay[7]};   // bit-1 branch into state i, minus x
  uint32_t M0 = wadd2(bb[0], m[0]), M1 = wadd2(bb[0], n[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) {
    M0 = addmax2(bb[i], m[i], M0);
    M1 = addmax2(bb[i], n[i], M1);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = addmax2(n[i], x, m[i]);
  if (TRACK || W == 8) {
    const uint32_t M1x = wadd2(M1, x);
    if (TRACK) rm.add2v(M0, M1x);
    uint32_t o = wsub2(M1x, M0);
    if (W == 8) o = sra1_2(o);  // the 8-window (sse16) decoder halves its output before the extrinsic subtraction
    return wsub2(o, aux);
  }
  return wsub2(wadd2(M1, sys), M0);
}

// alpha + output over one row group with the rebuilt beta in registers: B[i] is the beta above row k0 + i
// top: nullptr, or where (shared memory, [half * 32]) the beta above the group's LAST row is to be read from
// EDGE0: this is the group of rows 0..3 next to the known start state of the first window, in the reference's saturating
// arithmetic (alpha / output only; beta comes from the fast rebuild like everywhere else), normalised after row 2 only.
template <int W, int NORM, bool TRACK, bool HARD, bool EDGE0 = false>
__device__ __forceinline__ void fwd_group(const WinCtx<W>& c, const uint32_t* stab_dir, uint32_t k0, uint32_t a[8],
                                          const uint32_t (*B)[8], const uint4& sv, const uint4& pv, const uint4& av, char* Y,
                                          Range& ra, Range& rm, Range& rd, uint32_t& crc, bool skip0,
                                          const uint4* top = nullptr)
{
  // the scatter offsets as 16-bit loads: the halves come zero-extended from the load/store unit instead of costing
  // a mask and a shift on the ALU pipe, which the add-max instructions of this loop keep busy
  const uint32_t e16 = smem_u32(stab_dir + (k0 >> 2) * (W / 2 * 4));
  const uint32_t sa[4] = {sv.x, sv.y, sv.z, sv.w}, pa[4] = {pv.x, pv.y, pv.z, pv.w}, aa[4] = {av.x, av.y, av.z, av.w};
  uint32_t       d[4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint32_t x = wadd2(aa[r], sa[r]);
    if (EDGE0) {
      Range unused;
      unused.reset();
      const uint32_t xe = sadd2(aa[r], sa[r]);  // (DEC2: the systematic stream reads as zero)
      uint32_t       o  = alpha_out_step<false>(a, B[r], xe, pa[r], sadd2(xe, pa[r]), unused);
      if (W == 8) o = sra1_2(o);
      d[r] = wsub2(o, aa[r]);
    } else if (r == 3 && top) {
      const uint4    t0 = top[0], t1 = top[32];
      const uint32_t bt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
      d[r] = alpha_out2<W, TRACK>(a, bt, x, pa[r], sa[r], aa[r], rm);
    } else {
      d[r] = alpha_out2<W, TRACK>(a, B[r], x, pa[r], sa[r], aa[r], rm);
    }
    if (HARD) {
      // max(v, -1) + 32767 wraps to a negative number exactly when v > 0; a byte permute with sign replication turns
      // the sign bit of each half into a 32-bit mask (7 instructions per row instead of 10)
      const uint32_t m  = wadd2(max2(wadd2(d[r], aa[r]), 0xFFFFFFFFu), 0x7FFF7FFFu);
      const uint2    rr = *reinterpret_cast<const uint2*>(c.R + (k0 + r) * W);
      crc ^= (rr.x & sign_mask<0x9999>(m)) ^ (rr.y & sign_mask<0xBBBB>(m));
    }
    uint32_t o_lo, o_hi;  // (inline asm: the compiler would merge the halves into one load and take them apart again)
    asm("ld.shared.u16 %0, [%1];" : "=r"(o_lo) : "r"(e16 + 4 * r));
    asm("ld.shared.u16 %0, [%1];" : "=r"(o_hi) : "r"(e16 + 4 * r + 2));
    *reinterpret_cast<uint16_t*>(Y + o_lo) = (uint16_t)(d[r] & 0xFFFFu);
    *reinterpret_cast<uint16_t*>(Y + o_hi) = (uint16_t)(d[r] >> 16);
    if (r & 1) rd.add2v(d[r - 1], d[r]);
    if (EDGE0) {
      if (r == 2) {
        normalize<false>(a);
        if (TRACK) ra.add8(a);
      }
    } else if (NORM == 4 ? r == 2 : (r & 1) == 0) {
      if (r != 0 || !skip0) {
        normalize<true>(a);
        if (TRACK) ra.add8(a);
      }
    }
  }
}

// rebuild beta over the 8 rows b .. b+7 into registers: in: s = beta above row b+7 (as the backward pass left it BEFORE
// normalising it; top_norm: it was normalised afterwards, i.e. it is not the window's boundary vector).
// out: B[i] = beta above row b + i, i = 0..6 (values before their normalisation, like the reference stores them); the
// beta above row b + 7 is the input vector itself, which the caller reads again from shared memory when row b + 7 comes
// up (eight registers less through the whole unit).
template <int NORM>
__device__ __forceinline__ void rebuild8(uint32_t s[8], bool top_norm, const Unit& u, uint32_t (*B)[8])
{
  if (top_norm) normalize<true>(s);
#pragma unroll
  for (int g = 1; g >= 0; g--) {
    Rows4 q;
    rows_of(u.s[g], u.p[g], u.a[g], q);
#pragma unroll
    for (int r = 3; r >= 0; r--) {
      if (g == 0 && r == 0) continue;  // the beta above row b - 1 belongs to the chunk below
      beta_step<true>(s, q.x[r], q.y[r], wadd2(q.x[r], q.y[r]));
#pragma unroll
      for (int i = 0; i < 8; i++) B[g * 4 + r - 1][i] = s[i];
      if (NORM == 4 ? r == 0 : (r & 1) == 0) normalize<true>(s);
    }
  }
}

template <int W>
__device__ __forceinline__ Streams half_streams(const WinCtx<W>& c, bool dec2)
{
  // row group 0 of this thread in the three streams of this half iteration; missing inputs read the zero page
  const char* const zp = reinterpret_cast<const char*>(g_zero_page);
  Streams q;
  q.s = (dec2 || c.idle) ? zp + c.lane * 16 : c.in_item + c.sp_off;
  q.p = c.idle ? zp + c.lane * 16 : c.in_item + (dec2 ? 2u : 1u) * (size_t)c.s_bytes + c.sp_off;
  q.a = (c.noap || c.idle) ? zp + c.lane * 4
                           : reinterpret_cast<const char*>(dec2 ? c.E32 : c.A32) + c.lane * 4;
  return q;
}

// ---- checkpoints of the main fast path ------------------------------------------------------------------------------
// The window is cut into 16-row chunks from the TOP: chunk j covers rows L-16(j+1) .. L-16j-1, nc = L / 16 of them, and
// L % 16 rows (0, 4, 8 or 12) remain at the bottom.  The beta above every chunk (and above the bottom rows) is a
// checkpoint in the warp's shared memory: ck[j], j = 0 the window's boundary vector.  They hold beta BEFORE its
// normalisation, like the reference's beta array.
// (Tried: the odd 8-row boundaries as additional checkpoints in HBM, which removes stage A of the forward pass and a
// unit buffer -- 10 % fewer instructions, 48 KB more DRAM traffic per block, but 3 to 6 % slower: the loads of those
// checkpoints, and the two unit buffers whose refills both had their consumers in the next loop iteration, could not
// be scheduled without a wait that sits right behind freshly issued loads.)
template <int W>
__device__ __forceinline__ void ck_put(const WinCtx<W>& c, int j, const uint32_t s[8])
{
  c.ck[(j * 2 + 0) * 32] = make_uint4(s[0], s[1], s[2], s[3]);
  c.ck[(j * 2 + 1) * 32] = make_uint4(s[4], s[5], s[6], s[7]);
}
template <int W>
__device__ __forceinline__ void ck_get(const WinCtx<W>& c, int j, uint32_t s[8])
{
  const uint4 v0 = c.ck[(j * 2 + 0) * 32], v1 = c.ck[(j * 2 + 1) * 32];
  s[0] = v0.x; s[1] = v0.y; s[2] = v0.z; s[3] = v0.w; s[4] = v1.x; s[5] = v1.y; s[6] = v1.z; s[7] = v1.w;
}

// what the backward side hands to the forward side
struct BackOut {
  uint32_t a[8];  // alpha boundary vector of this thread's windows
  Range    rb, ra;
};

// ---- backward side of a half iteration: beta warm-up (rows 39..0), alpha warm-up (rows L-40..L-1), beta over the
// window with a checkpoint at every unit boundary.  The alpha warm-up runs BEFORE the beta pass: its five units are
// the first five of the beta pass, which finds them in registers.  Every buffer is refilled right after its unit has
// been consumed with the unit that is due four units later.
template <int W, int NORM, bool TRACK>
__device__ __forceinline__ void backward_side(const WinCtx<W>& c, bool dec2, BackOut* out)
{
  constexpr int WH = W / 2;
  constexpr int GB = (int)kGroupBytes;
  const int     ng = (int)c.L >> 2;       // row groups (L % 4 == 0)
  const int     nu = ng >> 1;             // 8-row units, numbered from the top
  const bool    rem = (ng & 1) != 0;      // rows 0..3 are a group of their own
  const Streams q0  = half_streams<W>(c, dec2);
  const Streams top = q0.at((ng - 10) * GB);  // rows L-40 ..
  Range rb, ra;
  rb.reset(); ra.reset();
  Unit F0, F1, F2, F3, F4;
  load_unit<8 * GB>(q0, F0);
  load_unit<6 * GB>(q0, F1);
  load_unit<4 * GB>(q0, F2);
  load_unit<2 * GB>(q0, F3);
  load_unit<0>(q0, F4);
  // L2: the units of the beta pass below the top five
#pragma unroll 1
  for (int u = 5; u < nu && u < 5 + 12 && (B200_PF & 1); u++) prefetch_unit<0>(q0.at((ng - 2 - 2 * u) * GB));

  uint32_t s[8];
#pragma unroll
  for (int i = 0; i < 8; i++) s[i] = kNegInf2;
  beta_unit<NORM, TRACK>(s, F0, rb, false);
  load_unit<0>(top, F0);
  beta_unit<NORM, TRACK>(s, F1, rb, false);
  load_unit<2 * GB>(top, F1);
  beta_unit<NORM, TRACK>(s, F2, rb, false);
  load_unit<4 * GB>(top, F2);
  beta_unit<NORM, TRACK>(s, F3, rb, false);
  load_unit<6 * GB>(top, F3);
  beta_unit<NORM, TRACK>(s, F4, rb, NORM == 2);  // the reference does not normalise after row 0; NORM = 4 does (see above)
  load_unit<8 * GB>(top, F4);
  exchange_beta_boundary<WH>(s, c.t, c.tail + (dec2 ? 6 : 0));
  if (TRACK) rb.add8full(s);

  {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = kNegInf2;
    alpha_unit<NORM, TRACK>(a, F0, ra, NORM == 2);  // the reference normalises on the warm-up counter: not after its row 0
    alpha_unit<NORM, TRACK>(a, F1, ra, false);
    alpha_unit<NORM, TRACK>(a, F2, ra, false);
    alpha_unit<NORM, TRACK>(a, F3, ra, false);
    alpha_unit<NORM, TRACK>(a, F4, ra, false);
    exchange_alpha_boundary<WH>(a, c.t);
#pragma unroll
    for (int i = 0; i < 8; i++) out->a[i] = a[i];
  }

  // beta pass: unit k covers the row groups ng-2-2k, ng-1-2k; units 0..4 are in F4, F3, F2, F1, F0.
  // Unit 0 first, so that the loop works on a ring of four buffers (F3, F2, F1, F0 = units k .. k+3) with no special
  // case: a conditional "this buffer came with the warm-up" inside the loop makes the compiler carry a fifth buffer
  // around and copy every refill into it at the back edge, waiting for loads it has only just issued.
  // A checkpoint at every second unit boundary (16 rows).
  ck_put<W>(c, 0, s);
  beta_unit<NORM, TRACK>(s, F4, rb, false);
  Streams q = q0.at((ng - 4) * GB);  // unit k; unit k + i is at q - 2 i GB
  int     k = 1;
#pragma unroll 1
  for (; k + 4 <= nu; k += 4) {  // k is odd: the parity of the boundaries is static
    if (k + 16 <= nu && (B200_PF & 1)) {  // L2: three bodies ahead
      prefetch_unit<-24 * GB>(q);
      prefetch_unit<-26 * GB>(q);
      prefetch_unit<-28 * GB>(q);
      prefetch_unit<-30 * GB>(q);
    }
    beta_unit<NORM, TRACK>(s, F3, rb, false);
    if (k + 4 < nu) refill_unit<-8 * GB>(q.after(s[1]), F3);
    ck_put<W>(c, (k + 1) >> 1, s);
    beta_unit<NORM, TRACK>(s, F2, rb, false);
    if (k + 5 < nu) refill_unit<-10 * GB>(q.after(s[1]), F2);
    beta_unit<NORM, TRACK>(s, F1, rb, false);
    if (k + 6 < nu) refill_unit<-12 * GB>(q.after(s[1]), F1);
    ck_put<W>(c, (k + 3) >> 1, s);
    beta_unit<NORM, TRACK>(s, F0, rb, NORM == 2 && !rem && k + 4 == nu);
    if (k + 7 < nu) refill_unit<-14 * GB>(q.after(s[1]), F0);
    q = q.at(-8 * GB);
  }
  // up to three units are left
  if (k < nu) beta_unit<NORM, TRACK>(s, F3, rb, NORM == 2 && !rem && k + 1 == nu);  // k is odd
  if (k + 1 < nu) {
    ck_put<W>(c, (k + 1) >> 1, s);
    beta_unit<NORM, TRACK>(s, F2, rb, NORM == 2 && !rem && k + 2 == nu);
  }
  if (k + 2 < nu) beta_unit<NORM, TRACK>(s, F1, rb, NORM == 2 && !rem);
  if (rem) {  // rows 3..0
    if ((nu & 1) == 0) ck_put<W>(c, nu >> 1, s);  // they are all that is left below the last chunk
    Rows4 r4;
    rows_of(ld128<0>(q0.s), ld128<0>(q0.p), make_uint4(ld32<0>(q0.a), ld32<128>(q0.a), ld32<256>(q0.a), ld32<384>(q0.a)), r4);
    beta_group<NORM, TRACK>(s, r4, rb, NORM == 2);
  }
  out->rb = rb;
  out->ra = ra;
}

// state of the forward pass that the out-of-line pieces update
struct FwdState {
  uint32_t a[8];
  Range    ra, rm, rd;
  uint32_t crc;
};

// alpha + output over the `nrows` (4 .. 16) rows 0 .. nrows-1 at the bottom of the window, from checkpoint j, one row at
// a time with beta in local memory: where the window is not a whole number of chunks, and (edge) rows 0..3 next to the
// known start state in the reference's saturating arithmetic.  Out of line: a few rows per half iteration.
template <int W, int NORM, bool TRACK, bool HARD>
__device__ __noinline__ void fwd_rows_slow(const WinCtx<W> c, bool dec2, int nrows, int j, bool top_norm, bool edge,
                                           FwdState* st)
{
  constexpr int b = 0;
  const Streams q = half_streams<W>(c, dec2);
  uint32_t x[16], y[16], sy[16], ax[16];
  for (int r = 0; r < nrows; r++) {
    const char* g = q.s + (r >> 2) * (int)kGroupBytes + (r & 3) * 4;
    sy[r] = __ldcg(reinterpret_cast<const uint32_t*>(g));
    y[r]  = __ldcg(reinterpret_cast<const uint32_t*>(q.p + (r >> 2) * (int)kGroupBytes + (r & 3) * 4));
    ax[r] = __ldcg(reinterpret_cast<const uint32_t*>(q.a + r * 128));
    x[r]  = wadd2(ax[r], sy[r]);
  }
  uint32_t B[16][8], s[8];
  ck_get<W>(c, j, s);
  for (int i = 0; i < 8; i++) B[nrows - 1][i] = s[i];
  if (top_norm) normalize<true>(s);
  for (int r = nrows - 1; r >= 1; r--) {
    beta_step<true>(s, x[r], y[r], wadd2(x[r], y[r]));
    for (int i = 0; i < 8; i++) B[r - 1][i] = s[i];
    if (NORM == 4 ? (r & 3) == 0 : (r & 1) == 0) normalize<true>(s);
  }
  uint32_t a[8];
  for (int i = 0; i < 8; i++) a[i] = st->a[i];
  Range ra = st->ra, rm = st->rm, rd = st->rd, unused;
  uint32_t crc = st->crc;
  unused.reset();
  char* const Y = out_base<W>(c, dec2);
  for (int r = 0; r < nrows; r++) {
    const uint32_t k = (uint32_t)(b + r);
    if (edge && k < (uint32_t)kExactRows) {  // exact arithmetic; the reference's schedule (after row 2) is both schedules'
      const uint32_t xe = dec2 ? ax[r] : sadd2(ax[r], sy[r]);
      uint32_t       o  = alpha_out_step<false>(a, B[r], xe, y[r], sadd2(xe, y[r]), unused);
      if (W == 8) o = sra1_2(o);
      store_out<W>(c, dec2, k, o, ax[r], rd, crc);
      if (k == 2) {
        normalize<false>(a);
        if (TRACK) ra.add8(a);
      }
    } else {
      const uint32_t d = alpha_out2<W, TRACK>(a, B[r], x[r], y[r], sy[r], ax[r], rm);
      store_diff<W, HARD>(c, dec2, k, d, ax[r], rd, crc, Y);
      if ((NORM == 4 ? (k & 3) == 2 : (k & 1) == 0) && k != 0) {
        normalize<true>(a);
        if (TRACK) ra.add8(a);
      }
    }
  }
  for (int i = 0; i < 8; i++) st->a[i] = a[i];
  st->ra = ra; st->rm = rm; st->rd = rd; st->crc = crc;
}

// The chunk of rows 0..15 of a window whose length is a multiple of 16, with rows 0..3 in exact arithmetic (static and
// tracked tiers): one pass of forward_side's loop body, out of line, with its own loads (fwd_rows_slow did this chunk row
// by row with beta in local memory, 4 x slower per row: +12 % per half iteration).
template <int W, int NORM, bool TRACK, bool HARD>
__device__ __noinline__ void fwd_chunk_edge(const WinCtx<W> c, bool dec2, int j, FwdState* st)
{
  constexpr int GB = (int)kGroupBytes;
  uint32_t a[8], s[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = st->a[i];
  Range    ra = st->ra, rm = st->rm, rd = st->rd, rb;
  uint32_t crc = st->crc;
  rb.reset();
  const Streams q = half_streams<W>(c, dec2);  // the chunk starts at row 0
  Unit X, Yb;
  load_unit<0>(q, Yb);
  load_unit<2 * GB>(q, X);
  char* const           Yout = out_base<W>(c, dec2);
  const uint32_t* const sdir = c.stab + (dec2 ? 0u : kStabDir);
  uint32_t B[7][8];
  {
    ck_get<W>(c, j, s);
    if (j != 0) normalize<true>(s);
    Rows4 r4;
    rows_of(X.s[1], X.p[1], X.a[1], r4);
    beta_group<NORM, false>(s, r4, rb, false);
    rows_of(X.s[0], X.p[0], X.a[0], r4);
#pragma unroll
    for (int r = 3; r >= 1; r--) {
      beta_step<true>(s, r4.x[r], r4.y[r], wadd2(r4.x[r], r4.y[r]));
      if (NORM == 2 && r == 2) normalize<true>(s);
    }
    beta_step<true>(s, r4.x[0], r4.y[0], wadd2(r4.x[0], r4.y[0]));
    ck_put<W>(c, kMaxCk, s);
  }
  rebuild8<NORM>(s, true, Yb, B);
  fwd_group<W, NORM, TRACK, HARD, true>(c, sdir, 0, a, &B[0], Yb.s[0], Yb.p[0], Yb.a[0], Yout, ra, rm, rd, crc, false);
  fwd_group<W, NORM, TRACK, HARD>(c, sdir, 4, a, &B[4], Yb.s[1], Yb.p[1], Yb.a[1], Yout, ra, rm, rd, crc, false,
                                  c.ck + kMaxCk * 64);
  ck_get<W>(c, j, s);
  rebuild8<NORM>(s, j != 0, X, B);
  fwd_group<W, NORM, TRACK, HARD>(c, sdir, 8, a, &B[0], X.s[0], X.p[0], X.a[0], Yout, ra, rm, rd, crc, false);
  fwd_group<W, NORM, TRACK, HARD>(c, sdir, 12, a, &B[4], X.s[1], X.p[1], X.a[1], Yout, ra, rm, rd, crc, false,
                                  c.ck + j * 64);
#pragma unroll
  for (int i = 0; i < 8; i++) st->a[i] = a[i];
  st->ra = ra; st->rm = rm; st->rd = rd; st->crc = crc;
}

// ---- forward side: alpha + output over the window, 16 rows at a time: from the checkpoint beta first runs down the
// upper 8 rows (stage A), then the lower 8 rows of beta are rebuilt into registers and alpha + the outputs run over
// them (stage B), then the same for the upper 8 rows (stage C).  edge: rows 0..3 in exact arithmetic.
template <int W, int NORM, bool TRACK, bool HARD>
__device__ __forceinline__ HalfResult forward_side(const WinCtx<W>& c, bool dec2, int G, bool edge, const BackOut& bo)
{
  constexpr int GB = (int)kGroupBytes;
  const int     L  = (int)c.L;
  const int     nc = L >> 4, R = L & 15;
  Range rb = bo.rb;
  FwdState f;
#pragma unroll
  for (int i = 0; i < 8; i++) f.a[i] = bo.a[i];
  f.ra = bo.ra;
  f.rd.reset();
  f.rm.hi = kMin2; f.rm.lo = kMax2;
  f.crc = 0;
  // the bottom of the window, out of line: the rows below the last chunk, or the last chunk when it has the edge rows
  int jb = nc - 1;  // lowest chunk not done yet
#ifndef B200_NOSLOW
  if (R) {
    fwd_rows_slow<W, NORM, TRACK, HARD>(c, dec2, R, nc, true, edge, &f);
  } else if (edge) {
    fwd_chunk_edge<W, NORM, TRACK, HARD>(c, dec2, jb, &f);
    jb--;
  }
#endif
  if (jb >= 0) {
    uint32_t a[8], s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = f.a[i];
    Range    ra = f.ra, rm = f.rm, rd = f.rd;
    uint32_t crc = f.crc;
    Streams  q = half_streams<W>(c, dec2).at((L - 16 * (jb + 1)) * 128);  // first row group of the chunk
    // upper unit of the chunk (stages A and C), upper unit of the next chunk (in flight from stage A on), lower unit
    Unit X, Xn, Yb;
    load_unit<2 * GB>(q, X);  // the backward pass has just read these rows
    load_unit<0>(q, Yb);
    char* const           Yout = out_base<W>(c, dec2);
    const uint32_t* const sdir = c.stab + (dec2 ? 0u : kStabDir);
#pragma unroll 1
    for (int j = jb; j >= 0; j--) {
      const uint32_t lo = (uint32_t)(L - 16 * (j + 1));
      if (j >= 2 && (B200_PF & 2)) {  // L2: two chunks ahead
        prefetch_unit<8 * GB>(q);
        prefetch_unit<10 * GB>(q);
      }
      uint32_t B[7][8];
      // ---- stage A: beta from the checkpoint down the upper 8 rows ----
      {
        ck_get<W>(c, j, s);
        if (j != 0) normalize<true>(s);  // every checkpoint but the boundary vector was normalised after it was stored
        Rows4 r4;
        rows_of(X.s[1], X.p[1], X.a[1], r4);
        beta_group<NORM, false>(s, r4, rb, false);
        rows_of(X.s[0], X.p[0], X.a[0], r4);
#pragma unroll
        for (int r = 3; r >= 1; r--) {
          beta_step<true>(s, r4.x[r], r4.y[r], wadd2(r4.x[r], r4.y[r]));
          if (NORM == 2 && r == 2) normalize<true>(s);
        }
        beta_step<true>(s, r4.x[0], r4.y[0], wadd2(r4.x[0], r4.y[0]));  // s = beta above row lo + 7, not yet normalised
        ck_put<W>(c, kMaxCk, s);  // the scratch vector: row lo + 7 reads it back
      }
      if (j > 0) load_unit_sp<6 * GB>(q, Xn);  // (its a-priori words follow after stage B: eight registers less here)
      // ---- stage B: the lower 8 rows ----
      rebuild8<NORM>(s, true, Yb, B);
      fwd_group<W, NORM, TRACK, HARD>(c, sdir, lo, a, &B[0], Yb.s[0], Yb.p[0], Yb.a[0], Yout, ra, rm, rd, crc,
                                      NORM == 2 && lo == 0);
      fwd_group<W, NORM, TRACK, HARD>(c, sdir, lo + 4, a, &B[4], Yb.s[1], Yb.p[1], Yb.a[1], Yout, ra, rm, rd, crc, false,
                                      c.ck + kMaxCk * 64);
      if (j > 0) {
        load_unit_a<6 * GB>(q, Xn);
        refill_unit<4 * GB>(q.after(a[1]), Yb);
      }
      // ---- stage C: the upper 8 rows ----
      ck_get<W>(c, j, s);
      rebuild8<NORM>(s, j != 0, X, B);
      fwd_group<W, NORM, TRACK, HARD>(c, sdir, lo + 8, a, &B[0], X.s[0], X.p[0], X.a[0], Yout, ra, rm, rd, crc, false);
      fwd_group<W, NORM, TRACK, HARD>(c, sdir, lo + 12, a, &B[4], X.s[1], X.p[1], X.a[1], Yout, ra, rm, rd, crc, false,
                                      c.ck + j * 64);
      X = Xn;
      q = q.at(4 * GB);
    }
    f.ra = ra; f.rm = rm; f.rd = rd; f.crc = crc;
  }
  __syncwarp();

  HalfResult res;
  res.dmax = range_absmax(f.rd);
  res.crc  = f.crc;
  bool ok  = true;
  if (TRACK) {  // the proof obligations of the tracked tier (see half_iteration_fast)
    const Range ra = f.ra, rm = f.rm;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int bh = h ? hi16(rb.hi) : lo16(rb.hi), bl = h ? hi16(rb.lo) : lo16(rb.lo);
      const int ah = h ? hi16(ra.hi) : lo16(ra.hi), al = h ? hi16(ra.lo) : lo16(ra.lo);
      const int mh = h ? hi16(rm.hi) : lo16(rm.hi), ml = h ? hi16(rm.lo) : lo16(rm.lo);
      ok = ok && (bh + 2 * G <= 32767) && (bl - 2 * G >= -32768) && (bh - bl + 4 * G <= 32767);
      ok = ok && (ah + 2 * G <= 32767) && (al - 2 * G >= -32768) && (ah - al + 4 * G <= 32767);
      ok = ok && (bh + ah + 5 * G <= 32767) && (bl + al - 5 * G >= -32768);
      ok = ok && (mh < ml || mh - ml <= 32767);
    }
    ok = ok && (G <= kMaxFastG);
  } else {
    ok = G <= kStaticFastG;
  }
  res.proven = ok;
  return res;
}

// one copy of the backward side serves both forward variants (hard: CRC modes)
template <int W, int NORM, bool TRACK>
__device__ __forceinline__ HalfResult half_fast2(const WinCtx<W>& c, bool dec2, int G, bool edge, bool hard)
{
  BackOut bo;
  backward_side<W, NORM, TRACK>(c, dec2, &bo);
  return hard ? forward_side<W, NORM, TRACK, true>(c, dec2, G, edge, bo)
              : forward_side<W, NORM, TRACK, false>(c, dec2, G, edge, bo);
}

// ---- general path (L % 4 != 0): staged chunks through the TMA ring, out of line like the tracked tier below.  The
// staging buffers share the warp's shared memory with the checkpoints of the main path, so every half iteration sets
// them up itself: no a-priori input -> the A part reads as zero; DEC2 has no separate systematic input (x = E) -> the
// sys part reads as zero (its copies leave that part alone), so both decoders run the same branch-free code.
template <int W, bool CRCOK = true>
__device__ __noinline__ HalfResult half_general(const WinCtx<W> c, bool dec2, int G, int tier, bool any_crc_, Pipe* pp)
{
  const bool any_crc = CRCOK && any_crc_;  // (kernels for launches without CRC carry no copy of the CRC variants)
  Pipe pipe = *pp;
  if (c.noap)
    for (int i = c.lane; i < kStages * 64; i += 32)
      *reinterpret_cast<uint4*>(c.stages + (i >> 6) * kStageBytes + 2048 + (i & 63) * 16) = make_uint4(0, 0, 0, 0);
  if (dec2)
    for (int i = c.lane; i < kStages * 64; i += 32)
      *reinterpret_cast<uint4*>(c.stages + (i >> 6) * kStageBytes + (i & 63) * 16) = make_uint4(0, 0, 0, 0);
  // generic-proxy writes (the above, the previous half iteration's outputs) before the bulk copies
  fence_proxy_async();
  __syncwarp();
  HalfResult r;
  if (tier == 0)
    r = any_crc ? half_fast<W, false, false, true>(c, dec2, G, pipe) : half_fast<W, false, false, false>(c, dec2, G, pipe);
  else if (tier == 1)
    r = any_crc ? half_fast<W, false, true, true>(c, dec2, G, pipe) : half_fast<W, false, true, false>(c, dec2, G, pipe);
  else
    r = any_crc ? half_fast<W, true, true, true>(c, dec2, G, pipe) : half_fast<W, true, true, false>(c, dec2, G, pipe);
  *pp = pipe;
  return r;
}

// The tracked tier (NORM = 2, bookkeeping) through the main path.  With it in the kernel -- inlined or out of line --
// ptxas spills eight registers of the plain tier's forward loop right behind the loads that fill them (5.05 instead
// of 4.63 ms per 65 536 blocks), so the window kernels exist twice: TRK2 = false without it (the tracked tier then
// runs through the general path; launched when no block checks a CRC, i.e. srslte_tdec_run_all batches, whose blocks
// stay in the plain tiers unless they converge) and TRK2 = true with it (launched for the CRC modes, where blocks
// converge, their extrinsic values grow and the tracked tier is the common case).
template <int W>
__device__ __noinline__ HalfResult half_tracked2(const WinCtx<W> c, bool dec2, int G, bool hard)
{
  return half_fast2<W, 2, true>(c, dec2, G, true, hard);
}

// QPP of this K as a scatter table, computed from (f1, f2) by the whole CTA:
//   pi(d*L + k) = pi(k) + L * d * (f1 + f2*d*L + 2*f2*k)  (mod K = W*L), so with pi(k) = w0*L + r every window
//   d of row k lands in row r, window (w0 + d*(f1 + f2*d*L) + 2*f2*d*k) mod W.
// dir 0 holds pi (where DEC2's row k goes in natural order), dir 1 its inverse (where natural row r goes in
// DEC2's order).  An entry is the pair of byte offsets (16 bits each) of the destinations of windows 2t and 2t+1
// inside the code block's part of the array: row * 128 + window * 2.
template <int W>
__device__ void build_tables(uint32_t K, uint32_t f1, uint32_t f2, uint32_t* stab)
{
  constexpr uint32_t WH = W / 2;
  const uint32_t L  = K / W;
  const uint32_t mK = (uint32_t)(0x100000000ull / K);
  const uint32_t mL = (uint32_t)((0x100000000ull + L - 1) / L);
  auto off = [](uint32_t row, uint32_t w) { return row * 128u + w * 2u; };
  for (uint32_t k = threadIdx.x; k < L; k += blockDim.x) {
    const uint32_t v = (f1 + f2 * k) * k;  // k < L <= 384 keeps it below 2^32
    uint32_t       p = v - __umulhi(v, mK) * K;
    if (p >= K) p -= K;
    if (p >= K) p -= K;
    const uint32_t w0 = __umulhi(p, mL), r = p - w0 * L;
    uint64_t       fw = 0, iv = 0;
#pragma unroll
    for (uint32_t d = 0; d < (uint32_t)W; d++) {
      const uint32_t w = (w0 + d * (f1 + f2 * d * L) + 2 * f2 * d * k) & (uint32_t)(W - 1);
      fw |= (uint64_t)w << (4 * d);
      iv |= (uint64_t)d << (4 * w);
    }
#pragma unroll
    for (uint32_t t = 0; t < WH; t++) {  // nibbles 2t, 2t+1: the destination windows of this thread's pair
      const uint32_t f = (uint32_t)(fw >> (8 * t)) & 0xFFu, u = (uint32_t)(iv >> (8 * t)) & 0xFFu;
      stab[((k >> 2) * WH + t) * 4 + (k & 3u)]            = off(r, f & 15u) | (off(r, f >> 4) << 16);
      stab[kStabDir + ((r >> 2) * WH + t) * 4 + (r & 3u)] = off(k, u & 15u) | (off(k, u >> 4) << 16);
    }
  }
}

// hard decision of this code block: bit n = (A[n] + E[pi^-1(n)] > 0), MSB first.
// Phase 1: every thread turns its two windows into bit strings (32 rows per word) in shared memory;
// phase 2: bytes are cut out of the concatenated window strings.  The bit strings live in the group's own beta
// chunk slots (free between half iterations).
template <int W>
__device__ void decide(const WinCtx<W>& c, uint8_t* out, bool write)
{
  constexpr int  WH = W / 2;
  constexpr int  NW = (kMaxL + 31) / 32;  // words per window
  const uint32_t L  = c.L;
  // bit-string word f (= window * NW + word) of this code block lives in component f % 4 of beta-chunk
  // slot f / (4 WH) of the group's thread (f / 4) % WH: only this group's own shared-memory slots are touched
  static_assert((W * NW + 4 * WH - 1) / (4 * WH) <= 2 * kChunk, "bit strings must fit the beta chunk slots");
  uint32_t* grp_base = reinterpret_cast<uint32_t*>(c.sm - c.t);
  auto      word     = [&](uint32_t f) -> uint32_t& {
    return grp_base[(((f >> 2) / WH) * kSmLanes + ((f >> 2) % WH)) * 4 + (f & 3)];
  };
  const char* E8 = reinterpret_cast<const char*>(c.E32) + c.grp * (2 * W);
  constexpr int NB = 16;  // rows gathered per batch: the E gather is latency bound, keep many loads in flight
  // 32 rows = one word per window.  Rows past L repeat row L - 1: their bits land below the last valid bit of the
  // window's last word, where phase 2 never looks.
#pragma unroll 1
  for (uint32_t k0 = 0; k0 < L; k0 += 32) {
    uint32_t acc_lo = 0, acc_hi = 0;
#pragma unroll
    for (int h = 0; h < 32 / NB; h++) {
      uint32_t av[NB], el[NB], eh[NB];
#pragma unroll
      for (int j = 0; j < NB; j++) {
        const uint32_t k   = min(k0 + h * NB + j, L - 1);
        const uint32_t e = stab_at<W>(c.stab, 1u, k);
        el[j] = *reinterpret_cast<const uint16_t*>(E8 + (e & 0xFFFFu));
        eh[j] = *reinterpret_cast<const uint16_t*>(E8 + (e >> 16));
        av[j] = c.A32[ae_word(k, (uint32_t)c.lane)];
      }
#pragma unroll
      for (int j = 0; j < NB; j++) {
        const uint32_t v = wadd2(av[j], el[j] | (eh[j] << 16));
        // max(v, -1) + 32767 wraps to a negative number exactly when v > 0: the sign bits are the decisions
        const uint32_t m = wadd2(max2(v, 0xFFFFFFFFu), 0x7FFF7FFFu);
        acc_hi = __funnelshift_l(m, acc_hi, 1);
        acc_lo = __funnelshift_l(m << 16, acc_lo, 1);
      }
    }
    word((2 * c.t) * NW + (k0 >> 5))     = acc_lo;
    word((2 * c.t + 1) * NW + (k0 >> 5)) = acc_hi;
  }
  __syncwarp();
  if (write && (L & 31) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
    // whole words: window d's string is words d * L/32 .. of the output, first row in the top bit (big endian)
    const uint32_t wpw = L >> 5;
    uint32_t*      o32 = reinterpret_cast<uint32_t*>(out);
    for (uint32_t j = (uint32_t)c.t, d = 0, i = (uint32_t)c.t; j < c.K / 32; j += WH, i += WH) {
      while (i >= wpw) {
        i -= wpw;
        d++;
      }
      o32[j] = __byte_perm(word(d * NW + i), 0, 0x0123);
    }
  } else if (write) {
    const uint32_t mL = (uint32_t)((0x100000000ull + L - 1) / L);
    for (uint32_t j = (uint32_t)c.t; j < c.K / 8; j += WH) {
      const uint32_t n = 8 * j;
      uint32_t       d = __umulhi(n, mL);
      uint32_t       off = n - d * L;
      uint32_t       byte;
      if (off + 8 <= L) {
        const uint32_t f  = d * NW + (off >> 5);
        const uint32_t bo = off & 31;
        const uint32_t w0 = word(f), w1 = (bo > 24) ? word(f + 1) : 0u;  // only when the byte straddles two words
        byte = (__funnelshift_l(w1, w0, bo) >> 24) & 0xFFu;
      } else {  // the byte straddles two windows (L not a multiple of 8)
        byte = 0;
        for (int b = 0; b < 8; b++) {
          if (off == L) {
            off = 0;
            d++;
          }
          byte = (byte << 1) | ((word(d * NW + (off >> 5)) >> (31 - (off & 31))) & 1u);
          off++;
        }
      }
      out[j] = (uint8_t)byte;
    }
  }
  __syncwarp();
}

// max over the W/2 threads of a code block
template <int WH>
__device__ __forceinline__ uint32_t group_max(uint32_t v)
{
#pragma unroll
  for (int o = WH / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  return v;
}

// One CTA per SM.  The CTA takes kWarps consecutive work items at a time; the host pads the item list so that
// they all have the same K (items with count 0 are fillers): the QPP tables are built once per CTA round, and the
// warps run the same phase of the same code at the same time (one copy of the hot loops in the instruction cache).
// NOCRC: the launch checks no CRC (srslte_tdec_run_all batches): the kernel carries no copy of the CRC variants of its
// loops (register allocation and the instruction cache see a smaller kernel).
template <int W, bool TRK2, bool NOCRC = false>
__global__ void __launch_bounds__(kThreads, kBlocksPerSm) tdec_win_kernel(const TdecLaunch a)
{
  constexpr int WH  = W / 2;
  extern __shared__ uint4 smem[];
  char*     warp_sm_all = reinterpret_cast<char*>(smem);                                   // [warps][kWarpSmem]
  uint32_t* stab        = reinterpret_cast<uint32_t*>(warp_sm_all + kWarps * kWarpSmem);   // [2][kStabDir]
  uint64_t* mbar_all    = reinterpret_cast<uint64_t*>(stab + 2 * kStabDir);
  __shared__ uint32_t s_item;

  const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int      grp = lane / WH, t = lane % WH;
  const uint32_t slot = blockIdx.x * kWarps + warp;
  uint32_t       fallbacks = 0;
  uint32_t       tiers[4] = {0, 0, 0, 0};  // half iterations of this warp in the pure / static / tracked / exact variant
  Pipe           pipe;
  pipe.par = 0;
  if (lane == 0) {
    for (int s = 0; s < kStages; s++) mbar_init(mbar_all + warp * kStages + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  __syncthreads();

  uint32_t tab_K = 0;  // block size the scatter table was built for
  for (;;) {
    if (tid == 0) s_item = atomicAdd(a.counter, 1u);
    __syncthreads();
    const uint32_t rnd = s_item;
    if (rnd >= a.n_rounds) break;
    const uint2 round = a.rounds[rnd];  // (first item, number of items): one item per warp, all of the same K
    {
      const WorkItem w0 = a.items[round.x];
      if (w0.K != tab_K) {  // rounds of one block size follow each other
        build_tables<W>(w0.K, w0.f1, w0.f2, stab);
        tab_K = w0.K;
        __syncthreads();
      }
    }
    WorkItem wi;
    wi.count = 0;
    if ((uint32_t)warp < round.y) wi = a.items[round.x + (uint32_t)warp];
    if (wi.count != 0) {
      const bool     active = grp < (int)wi.count;
      const uint32_t b_eff  = active ? (uint32_t)grp : 0u;  // idle groups of a partial item shadow block 0
      const uint32_t cb     = a.order[wi.first + b_eff];

      WinCtx<W> c;
      c.K = wi.K;
      c.L = c.K / W;
      c.t = t;
      c.lane = lane;
      c.grp  = grp;
      c.ngroups  = (c.L + 3) >> 2;
      c.sp_off   = b_eff * 8u * W + (uint32_t)t * 16u;
      c.s_bytes  = c.ngroups * kGroupBytes;
      c.in_item  = reinterpret_cast<const char*>(a.in + (size_t)wi.in_pos * a.in_stride);
      c.tail     = reinterpret_cast<const int16_t*>(c.in_item + 3 * (size_t)c.s_bytes) + b_eff * 32;
      c.stab   = stab + t * 4;
      char* const warp_sm = warp_sm_all + warp * kWarpSmem;
      c.ck     = reinterpret_cast<uint4*>(warp_sm) + lane;
      c.stages = warp_sm;
      c.sm     = reinterpret_cast<uint4*>(warp_sm + kStages * kStageBytes) + lane;
      c.mbar   = mbar_all + warp * kStages;
      const uint16_t* meta = reinterpret_cast<const uint16_t*>(c.tail + 16);
      const int smax = meta[0], p0max = meta[1], p1max = meta[2];
      constexpr size_t XB = (W == 16) ? kXArrayBytes16 : kXArrayBytes8;
      char* ae = reinterpret_cast<char*>(a.ws_ae) + (size_t)slot * 2 * XB;
      c.A32 = reinterpret_cast<uint32_t*>(ae);
      c.E32 = reinterpret_cast<uint32_t*>(ae + XB);
      c.chk = reinterpret_cast<uint4*>(reinterpret_cast<char*>(a.ws_chk) + (size_t)slot * kChkSlotBytes) + lane;
      c.R   = nullptr;
      c.idle = false;

      uint8_t* out  = a.out + (size_t)cb * a.out_stride;
      uint32_t n    = 0, iters = 0;
      bool     done = false, ok = false;
      int      amax = 0, emax = 0;  // max |A|, max |E| over the code block
      const uint32_t crc_mode = a.crc_mode_cb ? a.crc_mode_cb[cb] : a.crc_mode;
      const int      which    = crc_mode == CRC_24A ? 0 : 1;
      const bool     any_crc  = NOCRC ? false : __any_sync(0xFFFFFFFFu, crc_mode != CRC_NONE);
      const bool     v2       = (c.L & 3u) == 0 && (a.force_exact & 8u) == 0;  // the main fast path takes this block size
      // CRC modes: this thread's window pair in the tables of the block's polynomial ([dir][row][window])
      const uint32_t* Rblk = any_crc ? a.crc_pos + a.crc_pos_off[wi.kidx] + (size_t)which * 2 * c.K + 2 * t : nullptr;
      // The first half iteration has no a-priori information: it does not read A.  A itself is only cleared when a
      // hard decision can be taken before DEC2 has written every element of it (a single half iteration, or a CRC
      // pass after the first): 48 KB of stores per work item that the plain K = 6144 / 4 half iterations case does
      // not need.
      if (a.max_iter < 2 || any_crc)
        for (uint32_t k = 0; k < c.L; k++) c.A32[k * 32 + lane] = 0;
      do {
        const bool dec2 = (n & 1) != 0;
        c.noap = n == 0;
        if (any_crc) c.R = Rblk + (dec2 ? c.K : 0u);
        // bound on |x|, |y|, |x + y| of this half iteration
        // A block whose CRC has passed has frozen its output: what its lanes compute from here on is never read (they
        // only keep the warp's shuffles and votes complete), so it neither picks the tier for the blocks that are
        // still being decoded nor can it send the warp to the exact variant.
        const int Gx = dec2 ? emax : smax + amax;
        const int G  = done ? 0 : Gx + (dec2 ? p1max : p0max);
        HalfResult r;
        bool       fast_ok = false;
        // what the previous half iteration stored must be visible to this one's loads (other lanes wrote it)
        __syncwarp();
        // the decision must be warp-uniform: the passes below use full-warp shuffles and votes
        // force_exact: bit 0 = exact variant only (tests); bits 1, 2 = skip the pure / the static tier, bit 3 = general
        // path only (measurements)
        const bool pure = __all_sync(0xFFFFFFFFu, (a.force_exact & 3u) == 0 && G <= kPureFastG);
        const bool stat = !pure && __all_sync(0xFFFFFFFFu, (a.force_exact & 5u) == 0 && G <= kStaticFastG);
        const bool trk  = !pure && !stat && __all_sync(0xFFFFFFFFu, (a.force_exact & 1u) == 0 && G <= kMaxFastG);
        if (v2 && (pure || stat)) {
          r       = half_fast2<W, 4, false>(c, dec2, G, stat, any_crc);
          fast_ok = true;
        } else if (TRK2 && v2 && trk) {
          if constexpr (TRK2) r = half_tracked2<W>(c, dec2, G, any_crc);
          fast_ok = __all_sync(0xFFFFFFFFu, r.proven || done);
#ifndef B200_V2ONLY
        } else if (pure || stat || trk) {
          r       = half_general<W, !NOCRC>(c, dec2, G, pure ? 0 : stat ? 1 : 2, any_crc, &pipe);
          fast_ok = (pure || stat) ? true : __all_sync(0xFFFFFFFFu, r.proven || done);
#endif
        }
        if (!fast_ok) {
          r = half_iteration_exact<W>(c, dec2);
          fallbacks++;
          tiers[3]++;
        } else {
          tiers[pure ? 0 : stat ? 1 : 2]++;
        }
        const int dm = (int)group_max<WH>(r.dmax);
        if (dec2) amax = dm; else emax = dm;
        n++;
        if (any_crc) {
          // the block's CRC = xor of its threads' partial CRCs; a block that passes freezes its output now (the
          // arrays of its warp slot keep changing while the other blocks of the item go on)
          uint32_t crc = r.crc;
#pragma unroll
          for (int o = WH / 2; o > 0; o >>= 1) crc ^= __shfl_xor_sync(0xFFFFFFFFu, crc, o);
          const bool check = crc_mode != CRC_NONE && !done && active;
          const bool pass  = check && crc == 0;
          if (check) iters = n;
          if (__any_sync(0xFFFFFFFFu, pass)) decide<W>(c, out, pass);
          if (pass) {
            ok   = true;
            done = true;
          }
        }
      } while (n < a.max_iter && !__all_sync(0xFFFFFFFFu, done || !active));
      // blocks that never passed (or do not check a CRC): the decision after the last half iteration
      if (__any_sync(0xFFFFFFFFu, active && !done)) decide<W>(c, out, active && !done);
      if (crc_mode == CRC_NONE) iters = n;
      if (active && t == 0) {
        if (a.n_iter) a.n_iter[cb] = (uint8_t)iters;
        if (a.crc_ok) a.crc_ok[cb] = ok ? 1 : 0;
      }
    }
    __syncthreads();  // every warp is done with the tables (and with s_item)
  }
  if (lane == 0 && fallbacks && a.stats) atomicAdd(a.stats, fallbacks);
  if (lane < 4 && a.stats && tiers[lane]) atomicAdd(a.stats + 1 + lane, tiers[lane]);
}

// Hard decision of the code blocks that finished in this half iteration, by the WHOLE CTA: a thread takes 32 consecutive
// bits of one block (natural order: bit n = window n / L, row n % L), gathers A[n] and E[pi^-1(n)] from the extrinsic
// arrays of the warp slot that decoded the block and stores one big-endian word.  (decide() above costs the warp that
// calls it about a quarter of a half iteration in load latency; with the CTA's warps in step -- see the kernel below --
// every warp would wait for it in nearly every half iteration.)
// fin[i] = (code block index, warp << 3 | group, bit 6: finished after its first half iteration, A reads as zero).  Call between CTA barriers: the arrays must be complete and unchanged.
template <int W>
__device__ __noinline__ void decide_coop(const int16_t* ws_ae, uint8_t* out, uint32_t out_stride, uint32_t K, const uint32_t* stab,
                                         const uint2* fin, uint32_t nfin)
{
  constexpr uint32_t WH = W / 2;
  constexpr size_t   XB = (W == 16) ? kXArrayBytes16 : kXArrayBytes8;
  constexpr int      NB = 16;
  const uint32_t L  = K / W;
  const uint32_t nw = (K + 31) >> 5;
  const uint32_t mL = (uint32_t)((0x100000000ull + L - 1) / L);
  if ((L & 31u) == 0 && (out_stride & 3u) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
    // Whole words per window.  A warp takes (block, 32 rows): its lanes are (row mod RPL, window), so one load
    // instruction reads whole 2W-byte pieces -- the W windows of a row of a block sit next to each other in A, and the
    // QPP sends them to ONE row of E (contention free) -- instead of 32 scattered sectors: the gathers of the general
    // form below are bound by the load/store unit (one wavefront per lane), these are 32 / W wavefronts per load.
    constexpr uint32_t RPL = 32 / W;  // rows per load instruction
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, r = lane / W, d = lane % W;
    const uint32_t wpw = L >> 5;      // words per window
    // the gathers are latency bound: a warp keeps the loads of two units (2 x 32 rows) in flight
    constexpr int NI = 32 / RPL;  // loads of each kind per lane and unit
    const uint32_t total = nfin * wpw;
    for (uint32_t u0 = warp; u0 < total; u0 += 2 * kWarps) {
      uint16_t    av[2][NI], ev[2][NI];
      uint32_t*   dst[2];
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const uint32_t u  = min(u0 + q * kWarps, total - 1);  // (an odd unit out is done twice)
        const uint32_t bi = u / wpw, j = u - bi * wpw;
        const uint2    f  = fin[bi];
        const uint32_t g  = f.y & 7u;
        const bool     noA = (f.y & 64u) != 0;
        const char*    A8 = reinterpret_cast<const char*>(ws_ae) + (size_t)(blockIdx.x * kWarps + ((f.y >> 3) & 7u)) * 2 * XB;
        const char*    E8 = A8 + XB + g * (2 * W);
        dst[q] = reinterpret_cast<uint32_t*>(out + (size_t)f.x * out_stride) + d * wpw + j;
#pragma unroll
        for (int i = 0; i < NI; i++) {
          const uint32_t k   = 32u * j + i * RPL + r;
          const uint32_t e   = stab[kStabDir + (k >> 2) * (WH * 4) + (d >> 1) * 4 + (k & 3u)];
          const uint32_t off = (d & 1u) ? e >> 16 : e & 0xFFFFu;
          av[q][i] = noA ? (uint16_t)0 : __ldcg(reinterpret_cast<const uint16_t*>(A8 + (k * 32u + g * WH + (d >> 1)) * 4u + (d & 1u) * 2u));
          ev[q][i] = __ldcg(reinterpret_cast<const uint16_t*>(E8 + off));
        }
      }
#pragma unroll
      for (int q = 0; q < 2; q++) {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < NI; i++)
          acc |= ((int16_t)(uint16_t)(av[q][i] + ev[q][i]) > 0 ? 1u : 0u) << (31u - (i * RPL + r));
#pragma unroll
        for (uint32_t o = W; o < 32; o <<= 1) acc |= __shfl_xor_sync(0xFFFFFFFFu, acc, o);
        if (r == 0) *dst[q] = __byte_perm(acc, 0, 0x0123);
      }
    }
    return;
  }
  for (uint32_t u = threadIdx.x; u < nfin * nw; u += blockDim.x) {
    const uint32_t bi = u / nw, j = u - bi * nw;
    const uint2    f  = fin[bi];
    const uint32_t g  = f.y & 7u;
    const bool     noA = (f.y & 64u) != 0;
    const char*    A8 = reinterpret_cast<const char*>(ws_ae) + (size_t)(blockIdx.x * kWarps + ((f.y >> 3) & 7u)) * 2 * XB;
    const char*    E8 = A8 + XB + g * (2 * W);
    const uint32_t n0 = 32u * j, nb = min(32u, K - n0);
    uint32_t       d = __umulhi(n0, mL), k = n0 - d * L;
    uint32_t       word = 0;
#pragma unroll
    for (int h = 0; h < 32 / NB; h++) {
      uint16_t av[NB], ev[NB];
#pragma unroll
      for (int i = 0; i < NB; i++) {
        const uint32_t e   = stab[kStabDir + (k >> 2) * (WH * 4) + (d >> 1) * 4 + (k & 3u)];
        const uint32_t off = (d & 1u) ? e >> 16 : e & 0xFFFFu;
        av[i] = noA ? (uint16_t)0 : __ldcg(reinterpret_cast<const uint16_t*>(A8 + (k * 32u + g * WH + (d >> 1)) * 4u + (d & 1u) * 2u));
        ev[i] = __ldcg(reinterpret_cast<const uint16_t*>(E8 + off));
        if ((uint32_t)(h * NB + i + 1) < nb) {  // bits past the block's end repeat its last bit (never stored)
          if (++k == L) {
            k = 0;
            d++;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < NB; i++) word = (word << 1) | ((int16_t)(uint16_t)(av[i] + ev[i]) > 0 ? 1u : 0u);
    }
    uint8_t* o = out + (size_t)f.x * out_stride + 4u * j;
    if (nb == 32 && (reinterpret_cast<uintptr_t>(o) & 3) == 0) {
      *reinterpret_cast<uint32_t*>(o) = __byte_perm(word, 0, 0x0123);
    } else {
      for (uint32_t b = 0; b < nb / 8; b++) o[b] = (uint8_t)(word >> (24 - 8 * b));
    }
  }
}

// ---- CRC modes: block-granular early termination -------------------------------------------------------------------
// reference: lib/src/phy/phch/sch.c:353-383 -- a code block leaves the loop at the half iteration its CRC passes.
// In the kernel above a warp keeps its W/2-thread groups busy until the LAST of its blocks stops and a CTA waits for its
// slowest warp, so blocks that converge early buy no time.  Here every thread group is its own worker: the work of one
// block size (an "epoch": the items of one K, tables built once) is one queue PER BLOCK POSITION of a work item (a group
// always works at its own lane position, so the item-interleaved input layout and the warp's extrinsic arrays stay as
// they are), and a group whose block has finished -- CRC passed or iteration cap reached -- writes its result and takes
// the next block of its queue while the other groups of the warp go on with theirs.  The groups of a warp are then at
// different half iterations: DEC1 / DEC2 (`dec2`) and "no a-priori input yet" (`noap`) are per-thread values, which the
// main fast path takes as they only select pointers; the tracked tier runs through the main path too (half_tracked2, out
// of line).  The general path stages whole-warp pieces by bulk copies and needs one parity per warp: the block sizes it
// serves (L % 4 != 0) refill a warp only as a whole (`aligned`).
// The warps of a CTA start every half iteration together (one CTA barrier): see the loop below.
template <int W>
__global__ void __launch_bounds__(kThreads, kBlocksPerSm) tdec_win_dyn_kernel(const TdecLaunch a)
{
  constexpr int      WH  = W / 2;
  constexpr uint32_t PER = 32 / WH;  // block positions of an item = queues of an epoch
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  extern __shared__ uint4 smem[];
  char*     warp_sm_all = reinterpret_cast<char*>(smem);
  uint32_t* stab        = reinterpret_cast<uint32_t*>(warp_sm_all + kWarps * kWarpSmem);
  uint64_t* mbar_all    = reinterpret_cast<uint64_t*>(stab + 2 * kStabDir);
  __shared__ uint32_t s_epoch;
  __shared__ uint32_t s_nfin[2];              // blocks that finished in this / the previous half iteration
  __shared__ uint2    s_fin[2][kWarps * PER];  // (code block, warp << 3 | group)

  const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int      grp = lane / WH, t = lane % WH;
  const uint32_t slot = blockIdx.x * kWarps + warp;
  uint32_t       fallbacks = 0;
  uint32_t       tiers[4] = {0, 0, 0, 0};
  Pipe           pipe;
  pipe.par = 0;
  if (lane == 0) {
    for (int s = 0; s < kStages; s++) mbar_init(mbar_all + warp * kStages + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid == 0) s_nfin[0] = s_nfin[1] = 0;
  fence_proxy_async();
  __syncthreads();

  uint32_t tab_K = 0, e_cur = 0, par = 0;  // par: which of the two lists this half iteration fills
  for (;;) {
    // the first epoch from e_cur on that still has blocks in one of its queues
    if (tid == 0) s_epoch = FULL;
    __syncthreads();
    for (uint32_t base = e_cur; base < a.n_epochs; base += kThreads) {
      const uint32_t e = base + (uint32_t)tid;
      if (e < a.n_epochs) {
        const uint32_t n = a.epochs[e].y;
        bool left = false;
        for (uint32_t g = 0; g < PER; g++)
          left = left || *reinterpret_cast<volatile const uint32_t*>(a.dyn_counters + e * PER + g) < n;
        if (left) atomicMin(&s_epoch, e);
      }
      __syncthreads();
      const uint32_t found = s_epoch;
      __syncthreads();  // nobody may update s_epoch for the next stretch before everybody has read it
      if (found != FULL) break;
    }
    const uint32_t e = s_epoch;
    if (e == FULL) break;
    const uint4    ep = a.epochs[e];  // (first item, items with blocks in them, first input position, blocks)
    const WorkItem w0 = a.items[ep.x];
    if (w0.K != tab_K) {
      build_tables<W>(w0.K, w0.f1, w0.f2, stab);
      tab_K = w0.K;
    }
    __syncthreads();

    WinCtx<W> c;
    c.K = w0.K;
    c.L = c.K / W;
    c.t = t;
    c.lane = lane;
    c.grp  = grp;
    c.ngroups  = (c.L + 3) >> 2;
    c.sp_off   = (uint32_t)lane * 16u;
    c.s_bytes  = c.ngroups * kGroupBytes;
    c.stab   = stab + t * 4;
    char* const warp_sm = warp_sm_all + warp * kWarpSmem;
    c.ck     = reinterpret_cast<uint4*>(warp_sm) + lane;
    c.stages = warp_sm;
    c.sm     = reinterpret_cast<uint4*>(warp_sm + kStages * kStageBytes) + lane;
    c.mbar   = mbar_all + warp * kStages;
    constexpr size_t XB = (W == 16) ? kXArrayBytes16 : kXArrayBytes8;
    char* ae = reinterpret_cast<char*>(a.ws_ae) + (size_t)slot * 2 * XB;
    c.A32 = reinterpret_cast<uint32_t*>(ae);
    c.E32 = reinterpret_cast<uint32_t*>(ae + XB);
    c.chk = reinterpret_cast<uint4*>(reinterpret_cast<char*>(a.ws_chk) + (size_t)slot * kChkSlotBytes) + lane;
    c.idle = false;
    const bool v2      = (c.L & 3u) == 0 && (a.force_exact & 8u) == 0;
    const bool aligned = !v2;  // the general path needs one parity per warp: refill the warp only as a whole
    uint32_t* const queue = a.dyn_counters + e * PER + (uint32_t)grp;
    const uint32_t* const Rk = a.crc_pos + a.crc_pos_off[w0.kidx] + 2 * t;

    // per-group state (the same in all threads of a group).  A group without a block shadows allocated memory: the
    // first item of the epoch until it has had a block of its own, its last block afterwards.
    bool     have = false, dry = false;  // dry: the group's queue has run out
    uint32_t phase = 0;                  // half iterations the warp has run since it was last empty
    uint32_t cb = 0, n = 0, crc_mode = CRC_NONE;
    int      amax = 0, emax = 0, smax = 0, p0max = 0, p1max = 0;
    const uint32_t* Rblk = Rk;
    c.in_item = reinterpret_cast<const char*>(a.in + (size_t)w0.in_pos * a.in_stride);
    c.tail    = reinterpret_cast<const int16_t*>(c.in_item + 3 * (size_t)c.s_bytes) + grp * 32;
    if ((uint32_t)warp * (uint32_t)gridDim.x >= a.dyn_items) dry = true;  // small launches: one warp per CTA before a second one anywhere
    for (;;) {
      // ONE CTA barrier per half iteration.  The warps of the CTA start every half iteration together: they then run the
      // same loops at the same time and the instruction cache holds one of them (backward 10 KB, forward 24 KB); left to
      // themselves they spread over all of it and starve (measured: 56 % of the stall samples "no instructions", 1.4 x
      // the time of the round-based kernel).  The barrier also completes the list of the blocks that finished in the
      // previous half iteration, whose hard decisions the whole CTA takes before their groups fetch new blocks.
      const bool go = __syncthreads_or(__any_sync(FULL, have || !dry)) != 0;
      const uint32_t nfin = s_nfin[par ^ 1u];
      if (nfin) {
        decide_coop<W>(a.ws_ae, a.out, a.out_stride, c.K, stab, s_fin[par ^ 1u], nfin);
        __syncthreads();  // ... before the next blocks of these groups clear A, and before the list is reused
        if (tid == 0) s_nfin[par ^ 1u] = 0;
      }
      if (!go) break;
      // A group takes a new block only when the blocks of the other groups of its warp are about to run DEC1 as well
      // (or the warp is empty): the groups of a warp then share the parity, read the same array (A or E) and use whole
      // 128-byte lines of it -- at mixed parities every line of both arrays is fetched for half its bytes (78 instead of
      // 51 KB of DRAM reads per block and half iteration).  A block that finishes after an odd number of half iterations
      // leaves its group idle for one half iteration (it then reads the zero page).  That pays when blocks live long
      // (config 3 at -e 4.0, 7.2 half iterations per block: 12.8 -> 12.2 ms) and costs when they are short lived (-e 6.0,
      // 2.8 half iterations: 5.9 -> 6.3 ms), so a group waits only if its previous block took five half iterations or more.
      const bool all_idle = __all_sync(FULL, !have);
      if (all_idle) phase = 0;
      // (n still holds the count of the group's previous block; variant bit 9: never wait -- measurements)
      const bool in_step = (phase & 1u) == 0 || n < 5u || (a.force_exact & 512u) != 0;
      const bool want = !have && !dry && (aligned ? all_idle : in_step);
      if (__any_sync(FULL, want)) {
        uint32_t j = FULL;
        if (want && t == 0) j = atomicAdd(queue, 1u);
        j = __shfl_sync(FULL, j, grp * WH);
        bool got = false;
        if (want) {
          // item j of the epoch: the items of one block size are consecutive in the schedule, PER blocks and PER input
          // positions each, only the last one may be partial (no look-up in the item list: it would be one more
          // dependent load before the half iteration can start)
          const uint32_t b = j * PER + (uint32_t)grp;
          if (j < ep.y && b < ep.w) {
            got = true;
            cb  = a.order[w0.first + b];
            c.in_item = reinterpret_cast<const char*>(a.in + (size_t)(ep.z + j * PER) * a.in_stride);
            c.tail    = reinterpret_cast<const int16_t*>(c.in_item + 3 * (size_t)c.s_bytes) + grp * 32;
            const uint16_t* meta = reinterpret_cast<const uint16_t*>(c.tail + 16);
            smax = meta[0]; p0max = meta[1]; p1max = meta[2];
            crc_mode = a.crc_mode_cb ? a.crc_mode_cb[cb] : a.crc_mode;
            Rblk = Rk + (size_t)(crc_mode == CRC_24A ? 0 : 1) * 2 * c.K;
            n = 0; amax = 0; emax = 0;
            have = true;
          }
          dry = !got;
        }
        // (A is not cleared: the first half iteration does not read it, DEC2 writes all of it before DEC1 reads it again,
        // and a block that finishes after its very first half iteration is flagged so in the list of finished blocks)
      }
      const bool warp_busy = __any_sync(FULL, have);
      if (warp_busy) {
        // groups without a block follow the parity of the first group that has one
        const uint32_t hmask = __ballot_sync(FULL, have);
        const uint32_t nlead = __shfl_sync(FULL, n, __ffs(hmask) - 1);
        const bool     dec2  = ((have ? n : nlead) & 1u) != 0;
        const bool     mixed = !__all_sync(FULL, dec2) && __any_sync(FULL, dec2);
        c.noap = have && n == 0;
        c.idle = !have;
        c.R    = Rblk + (dec2 ? c.K : 0u);
        phase++;
        const int Gx = dec2 ? emax : smax + amax;
        const int G  = have ? Gx + (dec2 ? p1max : p0max) : 0;
        HalfResult r;
        bool       fast_ok = false;
        __syncwarp();
        const bool pure = __all_sync(FULL, (a.force_exact & 3u) == 0 && G <= kPureFastG);
        const bool stat = !pure && __all_sync(FULL, (a.force_exact & 5u) == 0 && G <= kStaticFastG);
        const bool trk  = !pure && !stat && __all_sync(FULL, (a.force_exact & 1u) == 0 && G <= kMaxFastG);
        if (v2 && (pure || stat)) {
          r       = half_fast2<W, 4, false>(c, dec2, G, stat, true);
          fast_ok = true;
        } else if (v2 && trk) {
          // the tracked tier through the main path (out of line): per-thread parities are fine there, and a warp that
          // went through the general path alone kept the other seven waiting at the barrier for five half iterations
          r       = half_tracked2<W>(c, dec2, G, true);
          fast_ok = __all_sync(FULL, r.proven || !have);
        } else if (!mixed && (pure || stat || trk)) {
          // general path (L % 4 != 0): the groups of the warp were fetched together, so they share parity and `noap`
          c.noap  = __any_sync(FULL, c.noap);
          r       = half_general<W>(c, __any_sync(FULL, dec2), G, pure ? 0 : stat ? 1 : 2, true, &pipe);
          fast_ok = (pure || stat) ? true : __all_sync(FULL, r.proven || !have);
        }
        if (!fast_ok) {
          r = half_iteration_exact<W>(c, dec2);
          fallbacks++;
          tiers[3]++;
        } else {
          tiers[pure ? 0 : stat ? 1 : 2]++;
        }
        const int dm = (int)group_max<WH>(r.dmax);
        if (dec2) amax = dm; else emax = dm;
        n++;
        uint32_t crc = r.crc;
#pragma unroll
        for (int o = WH / 2; o > 0; o >>= 1) crc ^= __shfl_xor_sync(FULL, crc, o);
        const bool pass   = have && crc_mode != CRC_NONE && crc == 0;
        const bool finish = have && (pass || n >= a.max_iter);
        if (finish) {
          if (t == 0) {
            s_fin[par][atomicAdd(&s_nfin[par], 1u)] = make_uint2(cb, (uint32_t)(warp << 3 | grp) | (n == 1 ? 64u : 0u));
            if (a.n_iter) a.n_iter[cb] = (uint8_t)n;
            if (a.crc_ok) a.crc_ok[cb] = pass ? 1 : 0;
          }
          have = false;
        }
      }
      par ^= 1u;
    }
    __syncthreads();  // every warp is done with the tables (and with s_epoch)
    e_cur = e + 1;
  }
  if (lane == 0 && fallbacks && a.stats) atomicAdd(a.stats, fallbacks);
  if (lane < 4 && a.stats && tiers[lane]) atomicAdd(a.stats + 1 + lane, tiers[lane]);
}

// ---- generic decoder (K <= 400): one thread per PAIR of code blocks, wrapping arithmetic ------------------------
// reference: lib/src/phy/fec/turbodecoder_gen.c:54-231 (no windows: beta over all K + 3 rows, normalisation every 4,
// plain wrapping int16 arithmetic -- exactly the packed VIADD.16x2 / VIADDMNMX.S16x2 steps of the window kernels'
// fast variant, here with no proof obligations).
// A work item is up to 64 code blocks of equal K; a lane owns two of them as the halves of its 32-bit registers.
// Everything is laid out [row][block] (inputs by to_internal_kernel, A / E / beta per warp slot), and every row
// index -- also pi(k) -- is the same for all lanes, so each access is one coalesced row of words; the rows of the
// next 4 trellis steps are loaded while the current 4 are computed (the recursion is a chain of dependent
// operations: latency, not bandwidth, is what this kernel has to hide when a launch holds few small blocks).
struct GenCtx {
  int             K;
  uint32_t        stride;  // 32-bit words per input row = (blocks of the item rounded up to 2) / 2
  const uint32_t* in;      // item input, [3k + j][block pair] then 12 tail rows, already offset by this lane's pair
  uint32_t*       A;       // [k][32 lanes], offset by lane: extrinsic of DEC2 minus E (a-priori of DEC1), natural order
  uint32_t*       E;       // [k][32 lanes]: a-posteriori of DEC1 minus A (systematic of DEC2), natural order
  uint4*          beta;    // [k][half][32 lanes], offset by lane; k = 0..K+3
  const uint16_t* pi;      // shared: pi(k), k < K
};

struct GenRaw {
  uint32_t x, y, ap;
};

// the three words of trellis row k < K; r3 = 3 * stride is loop invariant, so a group of 4 rows costs one multiply
__device__ __forceinline__ GenRaw gen_load(const GenCtx& c, bool dec2, uint32_t k, uint32_t r3)
{
  GenRaw r;
  if (!dec2) {
    r.x  = c.in[k * r3];
    r.y  = c.in[k * r3 + c.stride];
    r.ap = c.A[k * 32];
  } else {
    r.x  = c.E[(uint32_t)c.pi[k] * 32];
    r.y  = c.in[k * r3 + 2 * c.stride];
    r.ap = 0;
  }
  return r;
}

// K is a multiple of 8 for every LTE block size: the K rows split into whole groups of 4, the 3 tail rows are peeled
__device__ void gen_half_iteration(const GenCtx& c, bool dec2)
{
  const uint32_t K = (uint32_t)c.K, r3 = 3 * c.stride;
  uint32_t       s[8];
  s[0] = 0;
#pragma unroll
  for (int i = 1; i < 8; i++) s[i] = kNegInf2;
  GenRaw cur[4], nxt[4];
#pragma unroll
  for (int j = 0; j < 4; j++) cur[j] = gen_load(c, dec2, K - 1 - j, r3);
  // ---- beta: the 3 tail rows K+2 .. K (no a-priori, no normalisation) ----
  {
    const uint32_t* tl = c.in + (3 * K + (dec2 ? 6u : 0u)) * c.stride;
    uint32_t        tx[3], ty[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      tx[r] = tl[(2 * r) * c.stride];
      ty[r] = tl[(2 * r + 1) * c.stride];
    }
#pragma unroll
    for (int r = 2; r >= 0; r--) {
      beta_step<true>(s, tx[r], ty[r], wadd2(tx[r], ty[r]));
      c.beta[((K + r) * 2 + 0) * 32] = make_uint4(s[0], s[1], s[2], s[3]);
      c.beta[((K + r) * 2 + 1) * 32] = make_uint4(s[4], s[5], s[6], s[7]);
    }
  }
  // ---- beta over rows K-1 .. 0, four at a time; the rows of the next TWO groups are in flight (a beta step is
  // short, two groups of arithmetic do not cover a trip to L2) ----
  GenRaw nx2[4];
#pragma unroll
  for (int j = 0; j < 4; j++) nxt[j] = gen_load(c, dec2, K >= 8 ? K - 5 - j : (uint32_t)j, r3);
  for (uint32_t k0 = K - 1;; k0 -= 4) {
    const bool more = k0 >= 7;
    if (k0 >= 11) {
#pragma unroll
      for (int j = 0; j < 4; j++) nx2[j] = gen_load(c, dec2, k0 - 8 - j, r3);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t k = k0 - j;
      const uint32_t x = dec2 ? cur[j].x : wadd2(cur[j].x, cur[j].ap);  // the a-priori array is all zero at first
      const uint32_t y = cur[j].y;
      beta_step<true>(s, x, y, wadd2(x, y));
      c.beta[(k * 2 + 0) * 32] = make_uint4(s[0], s[1], s[2], s[3]);
      c.beta[(k * 2 + 1) * 32] = make_uint4(s[4], s[5], s[6], s[7]);
      if (j == 3) normalize<true>(s);  // k % 4 == 0 (k0 = 3 mod 4)
    }
    if (!more) break;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      cur[j] = nxt[j];
      nxt[j] = nx2[j];
    }
  }
  // ---- alpha + output over steps 1 .. K (step k uses row k-1 and beta[k]) ----
  s[0] = 0;
#pragma unroll
  for (int i = 1; i < 8; i++) s[i] = kNegInf2;
  uint4 bcur[4][2], bnxt[4][2];
  Range unused;
  unused.reset();
#pragma unroll
  for (int j = 0; j < 4; j++) {
    cur[j]     = gen_load(c, dec2, (uint32_t)j, r3);
    bcur[j][0] = c.beta[((1 + j) * 2 + 0) * 32];
    bcur[j][1] = c.beta[((1 + j) * 2 + 1) * 32];
  }
  for (uint32_t k0 = 1;; k0 += 4) {  // steps k0 .. k0+3
    const bool more = k0 + 4 <= K;
    if (more) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        nxt[j]     = gen_load(c, dec2, k0 + 3 + j, r3);
        bnxt[j][0] = c.beta[((k0 + 4 + j) * 2 + 0) * 32];
        bnxt[j][1] = c.beta[((k0 + 4 + j) * 2 + 1) * 32];
      }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t k   = k0 + j;
      const uint32_t aux = dec2 ? cur[j].x : cur[j].ap;
      const uint32_t x   = dec2 ? cur[j].x : wadd2(cur[j].x, cur[j].ap);
      const uint32_t y   = cur[j].y;
      const uint32_t bb[8] = {bcur[j][0].x, bcur[j][0].y, bcur[j][0].z, bcur[j][0].w,
                              bcur[j][1].x, bcur[j][1].y, bcur[j][1].z, bcur[j][1].w};
      const uint32_t o = alpha_out_step<true, false>(s, bb, x, y, wadd2(x, y), unused);
      if (j == 3) normalize<true>(s);  // k % 4 == 0 (k0 = 1 mod 4)
      const uint32_t d = wsub2(o, aux);
      if (!dec2)
        c.E[(k - 1) * 32] = d;
      else
        c.A[(uint32_t)c.pi[k - 1] * 32] = d;
    }
    if (!more) break;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      cur[j]     = nxt[j];
      bcur[j][0] = bnxt[j][0];
      bcur[j][1] = bnxt[j][1];
    }
  }
}

__global__ void __launch_bounds__(kGenThreads) tdec_gen_kernel(const TdecLaunch a)
{
  constexpr uint32_t KMAX = 400;
  __shared__ uint16_t s_pi[kGenThreads / 32][KMAX];
  const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t slot = blockIdx.x * (kGenThreads / 32) + warp;
  for (;;) {
    uint32_t it = 0;
    if (lane == 0) it = atomicAdd(a.counter, 1u);
    it = __shfl_sync(0xFFFFFFFFu, it, 0);
    if (it >= a.n_items) break;
    const WorkItem wi = a.items[it];
    // this lane's two code blocks (low / high half of its registers); idle halves and lanes shadow valid storage
    const uint32_t npair = ((uint32_t)wi.count + 1) / 2;
    const bool     act[2] = {2 * (uint32_t)lane < wi.count, 2 * (uint32_t)lane + 1 < wi.count};
    const uint32_t pair_eff = (uint32_t)lane < npair ? (uint32_t)lane : 0u;
    uint32_t       cb[2];
    cb[0] = a.order[wi.first + (act[0] ? 2 * lane : 0)];
    cb[1] = a.order[wi.first + (act[1] ? 2 * lane + 1 : 0)];
    GenCtx c;
    c.K      = (int)wi.K;
    c.stride = npair;
    c.in     = reinterpret_cast<const uint32_t*>(a.in + (size_t)wi.in_pos * a.in_stride) + pair_eff;
    c.A      = reinterpret_cast<uint32_t*>(a.ws_ae) + (size_t)slot * 2 * KMAX * 32 + lane;
    c.E      = c.A + KMAX * 32;
    c.beta   = reinterpret_cast<uint4*>(a.ws_chk) + (size_t)slot * (KMAX + 4) * 2 * 32 + lane;
    c.pi     = s_pi[warp];
    __syncwarp();
    for (uint32_t k = (uint32_t)lane; k < wi.K; k += 32)  // pi(k) = (f1 k + f2 k^2) mod K, K <= 400: fits 32 bits
      s_pi[warp][k] = (uint16_t)(((uint32_t)wi.f1 * k + (uint32_t)wi.f2 * k * k) % wi.K);
    for (int k = 0; k < c.K; k++) c.A[k * 32] = 0;
    __syncwarp();

    uint32_t n = 0;
    uint32_t iters[2] = {0, 0};
    bool     done[2] = {false, false}, ok[2] = {false, false};
    uint32_t mode[2];
#pragma unroll
    for (int h = 0; h < 2; h++) mode[h] = a.crc_mode_cb ? a.crc_mode_cb[cb[h]] : a.crc_mode;
    // hard decision (A + E > 0, MSB first) of the blocks selected by wr, and on the fly the CRC of their bytes
    auto decide_gen = [&](const bool wr[2], uint32_t crc[2]) {
      crc[0] = crc[1] = 0;
      uint8_t* out0 = a.out + (size_t)cb[0] * a.out_stride;
      uint8_t* out1 = a.out + (size_t)cb[1] * a.out_stride;
      const int w0 = mode[0] == CRC_24A ? 0 : 1, w1 = mode[1] == CRC_24A ? 0 : 1;
      for (int j = 0; j < c.K / 8; j++) {
        uint32_t b0 = 0, b1 = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) {
          const uint32_t v = wadd2(c.A[(8 * j + b) * 32], c.E[(8 * j + b) * 32]);
          b0 = (b0 << 1) | (lo16(v) > 0 ? 1u : 0u);
          b1 = (b1 << 1) | (hi16(v) > 0 ? 1u : 0u);
        }
        if (wr[0]) out0[j] = (uint8_t)b0;
        if (wr[1]) out1[j] = (uint8_t)b1;
        crc[0] = ((crc[0] << 8) ^ c_crc_tab[w0][((crc[0] >> 16) & 0xFFu) ^ b0]) & 0xFFFFFFu;
        crc[1] = ((crc[1] << 8) ^ c_crc_tab[w1][((crc[1] >> 16) & 0xFFu) ^ b1]) & 0xFFFFFFu;
      }
    };
    bool busy;
    do {
      gen_half_iteration(c, (n & 1) != 0);
      n++;
      bool chk[2];
#pragma unroll
      for (int h = 0; h < 2; h++) chk[h] = mode[h] != CRC_NONE && !done[h] && act[h];
      if (chk[0] || chk[1]) {
        uint32_t crc[2];
        decide_gen(chk, crc);
#pragma unroll
        for (int h = 0; h < 2; h++)
          if (chk[h]) {
            iters[h] = n;
            if (crc[h] == 0) {
              ok[h]   = true;
              done[h] = true;
            }
          }
      }
      busy = (act[0] && !done[0]) || (act[1] && !done[1]);
    } while (n < a.max_iter && __any_sync(0xFFFFFFFFu, busy));
    {
      bool     fin[2];
      uint32_t crc[2];
#pragma unroll
      for (int h = 0; h < 2; h++) fin[h] = act[h] && mode[h] == CRC_NONE;
      if (fin[0] || fin[1]) decide_gen(fin, crc);
    }
#pragma unroll
    for (int h = 0; h < 2; h++)
      if (act[h]) {
        if (mode[h] == CRC_NONE) iters[h] = n;
        if (a.n_iter) a.n_iter[cb[h]] = (uint8_t)iters[h];
        if (a.crc_ok) a.crc_ok[cb[h]] = ok[h] ? 1 : 0;
      }
    __syncwarp();
  }
}

// ---- layout conversion into the decoder's internal layout -----------------------------------------
// Every code block is stored at its position in the decode schedule (dst_stride int16 per block); the blocks
// of one work item (`count` blocks of equal K starting at position `first`) are interleaved:
// window decoders: [sys | par0 | par1] streams; a stream is [row group of 4][lane = block of the item * W/2 + thread]
//                  [4 rows] 32-bit words (two windows each): 512 bytes per row group whatever the number of blocks in
//                  the item, Lp = L rounded up to 4 rows;
//                  then per block 16 int16 holding the 12 tail samples and 16 int16 of meta data:
//                  max |sys|, max |par0|, max |par1| (uint16) -- the inputs of the fast-path proof.
// generic decoder: [natural index 3i+j, then the 12 tail samples][block of the item, count rounded up to 2] -- one
//                  coalesced row per value, two blocks per 32-bit word.
// place[cb] = (first, count << 8 | index in the item); without it the schedule is the identity (uniform K).
__device__ __forceinline__ uint32_t windows_of(uint32_t K)
{
  return (K % 16 == 0 && K > 800) ? 16u : (K % 8 == 0 && K > 400) ? 8u : 0u;
}

// src_format 0: natural 3i+j (tails at 3K).  1: the reference's sub-block soft-buffer layout.
__global__ void __launch_bounds__(256) to_internal_kernel(const int16_t* __restrict__ src_all, uint32_t src_stride,
                                                          int16_t* __restrict__ dst_all, uint32_t dst_stride,
                                                          const uint32_t* __restrict__ cb_K, uint32_t uniform_K,
                                                          uint32_t src_format, const uint64_t* __restrict__ src_off,
                                                          const uint2* __restrict__ place)
{
  extern __shared__ int16_t stage[];  // natural input of one code block (format 0 only)
  __shared__ uint32_t s_max[3];
  const uint32_t cb = blockIdx.x;
  const uint32_t K  = cb_K ? cb_K[cb] : uniform_K;
  const int16_t* src = src_all + (src_off ? (size_t)src_off[cb] : (size_t)cb * src_stride);
  const uint32_t W = windows_of(K);
  uint32_t       first, cnt, b;
  if (place) {
    const uint2 pl = place[cb];
    first = pl.x;
    cnt   = pl.y >> 8;
    b     = pl.y & 0xFFu;
  } else {
    const uint32_t ipw = W ? 64u / W : 64u;  // code blocks per work item
    first = cb / ipw * ipw;
    cnt   = min(ipw, gridDim.x - first);
    b     = cb - first;
  }
  if (W == 0) {  // generic decoder: value i of the block goes to row i of the item, column b (rows of an even length)
    int16_t*       dst = dst_all + (size_t)first * dst_stride + b;
    const uint32_t rs  = (cnt + 1) & ~1u;
    for (uint32_t i = threadIdx.x; i < 3 * K + 12; i += blockDim.x) dst[(size_t)i * rs] = src[i];
    return;
  }
  const uint32_t L = K / W, Lp = (L + 3) & ~3u, WH = W / 2, S = Lp * W;
  const uint32_t per = 64u / W;                                // lanes are laid out for a full item
  int16_t*       item = dst_all + (size_t)first * dst_stride;  // the work item's interleaved storage
  int16_t*       tailp = item + 3 * (size_t)per * S + b * 32;
  (void)cnt;
  // the staged copy gives every window kStagePad extra int16 so that the W/2 threads that later read the same
  // row of different windows fall into different shared-memory banks (3L int16 per window is a multiple of 64
  // words for L = 384: an 8-way conflict without the padding)
  const uint32_t wstride = 3 * L + kStagePad;
  if (threadIdx.x < 3) s_max[threadIdx.x] = 0;
  if (src_format == 0) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    if (((3 * L) & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 7) == 0) {  // 8-byte aligned window starts
      // asynchronous 8-byte copies straight into shared memory: all of a thread's pieces are in flight at once
      const uint2*   s64 = reinterpret_cast<const uint2*>(src);
      const uint32_t t64 = smem_u32(stage);
      for (uint32_t d = warp; d < W; d += nwarps) {
        const uint2*   sp = s64 + d * (3 * L / 4);
        const uint32_t tp = t64 + d * (wstride / 4) * 8;
        for (uint32_t p = lane; p < 3 * L / 4; p += 32)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tp + p * 8), "l"(sp + p) : "memory");
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (((3 * L) & 1) == 0) {  // window starts are word aligned: 32-bit copies
      const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
      uint32_t*       t32 = reinterpret_cast<uint32_t*>(stage);
      for (uint32_t d = warp; d < W; d += nwarps)
        for (uint32_t p = lane; p < 3 * L / 2; p += 32) t32[d * (wstride / 2) + p] = s32[d * (3 * L / 2) + p];
    } else {
      for (uint32_t d = warp; d < W; d += nwarps)
        for (uint32_t p = lane; p < 3 * L; p += 32) stage[d * wstride + p] = src[d * 3 * L + p];
    }
    if (threadIdx.x < 12) stage[W * wstride + threadIdx.x] = src[3 * K + threadIdx.x];
  }
  __syncthreads();
  const uint32_t groups = Lp / 4;
  // one thread = one window pair t of one row group kg: 4 rows x 3 streams in, three 128-bit words out
  const uint32_t t = threadIdx.x % WH, ng = blockDim.x / WH;
  uint32_t       mx2[3] = {0, 0, 0};  // packed unsigned max of |.| per stream
  const bool wide = src_format == 0 && (L & 3) == 0;  // the 12 samples of a row group are three aligned 64-bit words
  for (uint32_t kg = threadIdx.x / WH; kg < groups; kg += ng) {
    uint32_t w[3][4];
    if (wide) {
      // samples 3r + j (row r, stream j) of both windows: word n / 2, half n % 2 -> one PRMT per output word
      const uint2* pl = reinterpret_cast<const uint2*>(stage + (2 * t) * wstride + 12 * kg);
      const uint2* ph = reinterpret_cast<const uint2*>(stage + (2 * t + 1) * wstride + 12 * kg);
      const uint2  a0 = pl[0], a1 = pl[1], a2 = pl[2], b0 = ph[0], b1 = ph[1], b2 = ph[2];
      const uint32_t A[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y}, B[6] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y};
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const int n = 3 * r + j;
          w[j][r]     = __byte_perm(A[n >> 1], B[n >> 1], (n & 1) ? 0x7632u : 0x5410u);
          mx2[j]      = __vmaxu2(mx2[j], __vabs2(w[j][r]));
        }
    } else
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const uint32_t k = kg * 4 + r;
#pragma unroll
      for (int j = 0; j < 3; j++) w[j][r] = 0;
      if (k < L) {
        if (src_format == 0) {
          const int16_t* pl = stage + (2 * t) * wstride + 3 * k;
          const int16_t* ph = pl + wstride;
#pragma unroll
          for (int j = 0; j < 3; j++) w[j][r] = (uint32_t)(uint16_t)pl[j] | ((uint32_t)(uint16_t)ph[j] << 16);
        } else {
#pragma unroll
          for (int j = 0; j < 3; j++) w[j][r] = reinterpret_cast<const uint32_t*>(src + j * (K + 32))[k * WH + t];
        }
      }
#pragma unroll
      for (int j = 0; j < 3; j++) mx2[j] = __vmaxu2(mx2[j], __vabs2(w[j][r]));
    }
#pragma unroll
    for (int j = 0; j < 3; j++)
      reinterpret_cast<uint4*>(item + (size_t)j * per * S)[(kg * per + b) * WH + t] =
          make_uint4(w[j][0], w[j][1], w[j][2], w[j][3]);
  }
  uint32_t mx[3];
#pragma unroll
  for (int j = 0; j < 3; j++) mx[j] = max(mx2[j] & 0xFFFFu, mx2[j] >> 16);
  {  // the 12 tail samples count for every stream: the decoder's bound G must cover the tail steps too
    const int16_t* tq = src_format == 0 ? stage + W * wstride : src + 3 * (K + 32);
    uint32_t       tm = 0;
    if (threadIdx.x < 12) tm = (uint32_t)abs((int)tq[threadIdx.x]);
#pragma unroll
    for (int j = 0; j < 3; j++) mx[j] = max(mx[j], tm);
  }
  const int16_t* tl = src_format == 0 ? stage + W * wstride : src + 3 * (K + 32);
  if (threadIdx.x < 16) tailp[threadIdx.x] = threadIdx.x < 12 ? tl[threadIdx.x] : (int16_t)0;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    uint32_t v = mx[j];
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    if ((threadIdx.x & 31) == 0) atomicMax(&s_max[j], v);
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    uint16_t* meta = reinterpret_cast<uint16_t*>(tailp + 16);
    meta[threadIdx.x] = threadIdx.x < 3 ? (uint16_t)s_max[threadIdx.x] : (uint16_t)0;
  }
}

// ---- rate de-matching ----------------------------------------------------------------------------
__global__ void rm_rx_kernel(const int16_t* __restrict__ e, int16_t* __restrict__ work,
                             const uint16_t* __restrict__ tab_pool, const RmItem* __restrict__ items)
{
  extern __shared__ int16_t rm_img[];
  const RmItem   it  = items[blockIdx.x];
  const int16_t* src = e + it.e_off;
  rm_rx_body([src](uint32_t p) { return (int)src[p]; }, it.E, it.N, it.wl, tab_pool + it.tab_off, work + it.work_off,
             rm_img, it.overwrite != 0);
}

}  // namespace

// ---- host side -----------------------------------------------------------------------------------
cudaError_t upload_crc_tables()
{
  uint32_t tab[2][256];
  const uint32_t polys[2] = {0x1864CFBu, 0x1800063u};
  for (int w = 0; w < 2; w++)
    for (uint32_t b = 0; b < 256; b++) {
      uint32_t r = b << 16;
      for (int i = 0; i < 8; i++) r = (r & 0x800000u) ? ((r << 1) ^ polys[w]) : (r << 1);
      tab[w][b] = r & 0xFFFFFFu;
    }
  return cudaMemcpyToSymbol(c_crc_tab, tab, sizeof(tab));
}

int tdec_blocks_per_warp(int W) { return W == 16 ? 4 : W == 8 ? 8 : 64; }

// the window kernels take this many consecutive work items per CTA round; they must share K (host pads with
// count-0 items).  1 for the generic kernel.
int tdec_items_per_cta(int W) { return W ? kWarps : 1; }
int tdec_ctas_per_sm() { return kBlocksPerSm; }

uint32_t internal_len(uint32_t K)
{
  const uint32_t W = (K % 16 == 0 && K > 800) ? 16u : (K % 8 == 0 && K > 400) ? 8u : 0u;
  if (W == 0) return 2 * (3 * K + 12);  // rows of an even number of blocks: an item of 1 block takes 2 columns
  const uint32_t L = K / W, Lp = (L + 3) & ~3u;
  return 3 * Lp * W + 32;
}

uint32_t internal_positions(uint32_t K, uint32_t n)
{
  const uint32_t W = (K % 16 == 0 && K > 800) ? 16u : (K % 8 == 0 && K > 400) ? 8u : 0u;
  if (W == 0) return n;
  const uint32_t per = 64u / W;
  return (n + per - 1) / per * per;
}

cudaError_t tdec_geometry(int W, int device, TdecGeometry* g)
{
  int sms = 0;
  cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) return e;
  g->threads = W ? kThreads : kGenThreads;
  const size_t warps_per_block = (size_t)g->threads / 32;
  if (W == 0) {
    g->smem   = 0;
    g->blocks = sms * 4;
    const size_t slots = (size_t)g->blocks * warps_per_block;
    g->ws_ae_bytes  = slots * 32 * 2 * 400 * sizeof(uint32_t);
    g->ws_chk_bytes = slots * (400 + 4) * 2 * 32 * sizeof(uint4);
    return cudaSuccess;
  }
  g->smem = warps_per_block * kWarpSmem + 2 * kStabDir * sizeof(uint32_t) + warps_per_block * kStages * sizeof(uint64_t);
  int per_sm = 0;
  if (W == 16) {
    e = cudaFuncSetAttribute(tdec_win_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tdec_win_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tdec_win_dyn_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tdec_win_kernel<16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tdec_win_kernel<16, false>, kThreads, g->smem);
  } else {
    e = cudaFuncSetAttribute(tdec_win_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tdec_win_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tdec_win_dyn_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tdec_win_kernel<8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tdec_win_kernel<8, false>, kThreads, g->smem);
  }
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  if (const char* e = getenv("B200_BLOCKS_PER_SM")) {  // development probe
    const int v = atoi(e);
    if (v >= 1 && v < per_sm) per_sm = v;
  }
  g->blocks = sms * per_sm;
  const size_t slots = (size_t)g->blocks * warps_per_block;
  g->ws_ae_bytes  = slots * 2 * (W == 16 ? kXArrayBytes16 : kXArrayBytes8);
  g->ws_chk_bytes = slots * kChkSlotBytes;
  return cudaSuccess;
}

cudaError_t tdec_launch(int W, const TdecGeometry& g, const TdecLaunch& a, cudaStream_t s)
{
  if (a.n_items == 0 || (W && a.n_rounds == 0)) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(a.counter, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  const int want   = W ? (int)a.n_rounds : (int)((a.n_items + (kGenThreads / 32) - 1) / (kGenThreads / 32));
  const int blocks = want < g.blocks ? want : g.blocks;
  // CRC modes: blocks converge and their growing extrinsic values make the tracked tier the common case
  const bool trk2 = (a.force_exact & 32u) != 0;  // measured: not worth the slower plain tiers, see half_tracked2
  // CRC modes: block-granular early termination (force bit 6: the round-based kernel, for measurements and tests)
  const bool dyn = W && a.n_epochs && (a.crc_mode != CRC_NONE || a.crc_mode_cb) && (a.force_exact & 64u) == 0;
  if (dyn) {
    e = cudaMemsetAsync(a.dyn_counters, 0, sizeof(uint32_t) * a.n_epochs * (W == 16 ? 4 : 8), s);
    if (e != cudaSuccess) return e;
    // every CTA looks for work itself; a small launch spreads over the SMs instead of filling the first CTAs' warps
    const int dblocks = (int)a.dyn_items < g.blocks ? (int)a.dyn_items : g.blocks;
    if (W == 16)
      tdec_win_dyn_kernel<16><<<dblocks, g.threads, g.smem, s>>>(a);
    else
      tdec_win_dyn_kernel<8><<<dblocks, g.threads, g.smem, s>>>(a);
  } else if (W == 16 && !trk2 && a.crc_mode == CRC_NONE && !a.crc_mode_cb && (a.force_exact & 128u) == 0) {
    tdec_win_kernel<16, false, true><<<blocks, g.threads, g.smem, s>>>(a);
  } else if (W == 8 && !trk2 && a.crc_mode == CRC_NONE && !a.crc_mode_cb && (a.force_exact & 128u) == 0) {
    tdec_win_kernel<8, false, true><<<blocks, g.threads, g.smem, s>>>(a);
  } else if (W == 16 && trk2)
    tdec_win_kernel<16, true><<<blocks, g.threads, g.smem, s>>>(a);
  else if (W == 16)
    tdec_win_kernel<16, false><<<blocks, g.threads, g.smem, s>>>(a);
  else if (W == 8 && trk2)
    tdec_win_kernel<8, true><<<blocks, g.threads, g.smem, s>>>(a);
  else if (W == 8)
    tdec_win_kernel<8, false><<<blocks, g.threads, g.smem, s>>>(a);
  else
    tdec_gen_kernel<<<blocks, g.threads, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t to_internal_launch(const int16_t* src, uint32_t src_stride, const uint64_t* src_off, int src_format,
                               int16_t* dst, uint32_t dst_stride, const uint32_t* cb_K, uint32_t uniform_K,
                               const uint2* place, uint32_t n_cb, cudaStream_t s)
{
  if (n_cb == 0) return cudaSuccess;
  static bool attr_set[64] = {};  // per device: function attributes belong to the device's context
  int         dev = 0;
  cudaError_t e   = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const size_t smem = src_format == 0 ? (3 * 6144 + 16 * kStagePad + 16) * sizeof(int16_t) : 0;
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    e = cudaFuncSetAttribute(to_internal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)((3 * 6144 + 16 * kStagePad + 16) * sizeof(int16_t)));
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  to_internal_kernel<<<n_cb, 256, smem, s>>>(src, src_stride, dst, dst_stride, cb_K, uniform_K, (uint32_t)src_format,
                                             src_off, place);
  return cudaGetLastError();
}

cudaError_t rm_rx_launch(const int16_t* e, int16_t* work, const uint16_t* tab_pool, const RmItem* items,
                         uint32_t n_items, cudaStream_t s)
{
  if (n_items == 0) return cudaSuccess;
  rm_rx_kernel<<<n_items, 256, (kRmMaxWorkLen + 8) * sizeof(int16_t), s>>>(e, work, tab_pool, items);
  return cudaGetLastError();
}

}  // namespace b200

