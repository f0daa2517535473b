// tdec_kernels.h -- launch interface of the sm_100a turbo-decode kernels (internal to the library).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

// One warp-load of code blocks that share K (and therefore window count, window length and QPP).
struct WorkItem {
  uint32_t first;   // index of the first code block of this item in `order`
  uint32_t in_pos;  // position of the item in the internal-layout input (in code blocks).  The window decoders' items
                    // always take tdec_blocks_per_warp() positions, also when they hold fewer blocks, so the host pads
                    // every K group of the schedule to a multiple of that
  uint16_t count;   // code blocks in this item (<= blocks-per-warp of the kernel that runs it)
  uint16_t K;
  uint16_t f1, f2;  // QPP coefficients of K
  uint16_t kidx;    // index of K in the table of the 188 code-block sizes
  uint16_t pad;
};

enum CrcMode : uint32_t {
  CRC_NONE = 0,  // run exactly max(1, nof_iterations) half iterations (srslte_tdec_run_all)
  CRC_24B  = 1,  // stop a block when CRC24B over its K decoded bits is 0 (sch.c, C > 1)
  CRC_24A  = 2,  // same with CRC24A (sch.c, single-block transport block)
};

struct TdecLaunch {
  const int16_t*  in;          // internal-layout input (device), indexed by schedule position, see to_internal_launch
  uint32_t        in_stride;   // int16 elements per code block (multiple of 64)
  uint8_t*        out;         // decoded bytes (device)
  uint32_t        out_stride;  // bytes between code blocks
  uint8_t*        n_iter;      // [n_cb] half iterations run (device, nullable)
  uint8_t*        crc_ok;      // [n_cb] 1 when the block's CRC was 0 (device, nullable)
  const uint32_t* order;       // code-block indices grouped by K (device)
  const WorkItem* items;       // device
  uint32_t        n_items;
  const uint2*    rounds;      // window kernels: CTA rounds (first item, items <= tdec_items_per_cta(), same K) (device)
  uint32_t        n_rounds;
  uint32_t*       counter;     // device work counter, zeroed by the launcher
  const uint4*    epochs;      // window kernels, CRC modes (block-granular early termination): per block size of the launch
  uint32_t        n_epochs;    //   (first item, items that hold blocks, first input position, blocks), in launch order (device)
  uint32_t*       dyn_counters;//   [n_epochs][blocks per warp] queue heads, zeroed by the launcher
  uint32_t        dyn_items;   //   items that hold blocks, all epochs (sizes the number of warps a CTA puts to work)
  uint32_t        max_iter;    // half-iteration cap
  uint32_t        crc_mode;
  const uint8_t*  crc_mode_cb; // [n_cb] per-block CrcMode overriding crc_mode (device, nullable)
  int16_t*        ws_ae;       // extrinsic work arrays, sized by tdec_geometry()
  uint32_t*       ws_chk;      // beta checkpoints, sized by tdec_geometry()
  const uint32_t* crc_pos;     // window kernels, CRC modes: per-bit CRC contributions (lte_tables.h:crc_pos_tables)
  const uint32_t* crc_pos_off; // [188] offset of each K's tables in crc_pos (uint32 units)
  uint32_t        force_exact; // 1: always run the exact saturating variant (tests)
  uint32_t*       stats;       // device counters: [0] half iterations (per warp) that fell back to the exact variant,
                               // [1..4] half iterations (per warp) run in the pure / static / tracked / exact variant
};

struct TdecGeometry {
  int      blocks;          // grid size
  int      threads;         // block size
  size_t   smem;            // dynamic shared memory per block
  size_t   ws_ae_bytes;     // workspace for the a-priori / extrinsic arrays
  size_t   ws_chk_bytes;    // workspace for the beta checkpoints
};

// W = 16 or 8 (window decoders) or 0 (generic decoder, K <= 400)
cudaError_t tdec_geometry(int W, int device, TdecGeometry* g);
cudaError_t tdec_launch(int W, const TdecGeometry& g, const TdecLaunch& a, cudaStream_t s);
int         tdec_blocks_per_warp(int W);
int         tdec_ctas_per_sm();         // resident CTAs of the window kernels per SM
int         tdec_items_per_cta(int W);  // consecutive work items a CTA takes per round; they must share K

// int16 elements of one code block in the decoder's internal layout (pair-major streams + tail + meta
// for window decoders, natural order for the generic decoder).
uint32_t internal_len(uint32_t K);
// positions the internal-layout input of n code blocks of one size takes (a partial last item is padded)
uint32_t internal_positions(uint32_t K, uint32_t n);

// src_format 0: natural (3i+j, tails last); 1: the reference's sub-block soft-buffer layout.
// One code block per CTA; also records max |sys|, |par0|, |par1| per block for the fast-path proof.
// The source of block i is src + src_off[i] when src_off (device, int16 elements) is given, else src + i*src_stride.
// Blocks are written at their position in the decode schedule, the blocks of one work item interleaved:
// place[i] = (input position of i's work item, count << 8 | index in the item) (device); nullptr = the identity
// schedule of a uniform-K batch.  dst must hold internal_positions() x dst_stride elements.
cudaError_t to_internal_launch(const int16_t* src, uint32_t src_stride, const uint64_t* src_off, int src_format,
                               int16_t* dst, uint32_t dst_stride, const uint32_t* cb_K /* device, nullable */,
                               uint32_t uniform_K, const uint2* place, uint32_t n_cb, cudaStream_t s);

// rate de-matching: work[tab[i mod N]] += e[i], i < E, wrapping int16.
struct RmItem {
  uint32_t e_off;     // offset of this block's samples in the e buffer (int16 elements)
  uint32_t E;         // samples
  uint32_t work_off;  // offset of the block's working buffer (int16 elements)
  uint32_t tab_off;   // offset of its (K, rv) index table in the table pool (uint16 elements)
  uint32_t N;         // 3K+12
  uint32_t wl;        // int16 elements of the block's working buffer that the table can address (working_len(K))
  uint32_t overwrite; // 1: the working buffer is logically all zero (fresh HARQ buffer): store instead of add
};
constexpr uint32_t kRmMaxWorkLen = 18600;  // SOFTBUFFER_SIZE of the reference (softbuffer.h:50) >= working_len(6144)

// Shared by the two rate-dematching kernels: sum the wrap-around repeats of every rate-matched position, scatter the
// sums into a shared-memory image of the working buffer (table order), then add the image to the working buffer
// with coalesced 128-bit read-modify-writes (wrapping int16, like the reference's `+=`).
// add (or copy) the shared-memory image to the working buffer: coalesced 128-bit read-modify-writes, wrapping int16
__device__ __forceinline__ void rm_rx_add_image(const int16_t* img, int16_t* dst, uint32_t wl, bool overwrite)
{
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const uint32_t nv = wl / 8;
    uint4*         d4 = reinterpret_cast<uint4*>(dst);
    const uint4*   s4 = reinterpret_cast<const uint4*>(img);
    for (uint32_t j = threadIdx.x; j < nv; j += blockDim.x) {
      const uint4 a = s4[j];
      if (overwrite) {
        d4[j] = a;
      } else {
        uint4 v = d4[j];
        v.x = __vadd2(v.x, a.x); v.y = __vadd2(v.y, a.y); v.z = __vadd2(v.z, a.z); v.w = __vadd2(v.w, a.w);
        d4[j] = v;
      }
    }
    for (uint32_t j = nv * 8 + threadIdx.x; j < wl; j += blockDim.x)
      dst[j] = overwrite ? img[j] : (int16_t)(dst[j] + img[j]);
  } else {
    for (uint32_t j = threadIdx.x; j < wl; j += blockDim.x) dst[j] = overwrite ? img[j] : (int16_t)(dst[j] + img[j]);
  }
}

template <class Src>
__device__ __forceinline__ void rm_rx_body(const Src& src, uint32_t E, uint32_t N, uint32_t wl, const uint16_t* tab,
                                           int16_t* dst, int16_t* img /* shared, wl rounded up to 8 */, bool overwrite)
{
  const uint32_t wl8 = (wl + 7) & ~7u;
  for (uint32_t j = threadIdx.x; j < wl8 / 2; j += blockDim.x) reinterpret_cast<uint32_t*>(img)[j] = 0;
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < N && i < E; i += blockDim.x) {
    int acc = 0;
    for (uint32_t p = i; p < E; p += N) acc += src(p);  // wrap-around repeats hit the same cell
    img[tab[i]] = (int16_t)acc;                         // the table is one-to-one
  }
  __syncthreads();
  rm_rx_add_image(img, dst, wl, overwrite);
}
cudaError_t rm_rx_launch(const int16_t* e, int16_t* work, const uint16_t* tab_pool, const RmItem* items,
                         uint32_t n_items, cudaStream_t s);

// ---- front end: soft demodulation + descrambling (frontend_kernels.cu) ------------------------------
struct FeCodeword {
  uint32_t qm;        // bits per symbol: 2, 4, 6, 8
  uint32_t nsym;      // symbols of the codeword
  uint32_t c_init;    // seed of the scrambling sequence
  uint32_t nof_bits;  // LLRs that are descrambled (<= qm * nsym)
  uint64_t sym_off;   // first symbol in the symbol buffer (complex floats)
  uint64_t llr_off;   // first LLR in the output (demod_descramble only)
  uint32_t ul_cols;   // 0: the LLRs are used in channel order (PDSCH); else N_pusch_symbs: the UL-SCH channel
                      // de-interleaver of 36.212 5.2.2.8 (no UCI) sits between descrambling and rate de-matching
  uint32_t ul_rows;   // nof_bits / qm / ul_cols
  // UCI multiplexed into the PUSCH codeword (data path of srslte_ulsch_decode, sch.c:920-1064); all 0 without UCI
  uint32_t q_ack;     // Q'_ack coded HARQ-ACK symbols: their LLRs are erased (sch.c:961-964)
  uint32_t q_ri;      // Q'_ri coded RI symbols: skipped by the de-interleaver (ulsch_interleave_gen, sch.c:580-598)
  uint32_t g0_src;    // kNoG0 or the channel position whose LLR the reference leaves in g[0] (every RI sample is
                      // written to g[0], the last writer wins: srslte_vec_lut_sis over lut = 0, sch.c:589-590, 910)
  uint32_t g0_raw;    // 1: that LLR was flipped back by the 1-bit RI decoder (decode_ri_ack_1bit, uci.c:627-628)
};
constexpr uint32_t kNoG0 = 0xFFFFFFFFu;
// columns of the interleaver matrix that carry ACK / RI (uci.c:501-502, 526-527); normal CP set when N_pusch_symbs > 10
__host__ __device__ inline uint32_t uci_col(bool ri, bool norm, uint32_t c)
{
  // packed nibbles, entry c at bits 4c
  const uint32_t t = ri ? (norm ? 0xA741u : 0x8530u) : (norm ? 0x9832u : 0x7621u);
  return (t >> (4 * c)) & 15u;
}
struct RmSymItem {
  uint32_t E, work_off, tab_off, N, wl, overwrite;  // as RmItem
  uint32_t cw;                       // codeword the block belongs to
  uint32_t e_off;                    // first LLR of the block inside the codeword
};
constexpr uint32_t kGoldMaxLen = 256 * 1024;  // MAX_SEQ_LEN of the reference (sequence.c:34)
// x1: kGoldMaxLen / 32 packed words of x1(n + 1600); x2mask: kGoldMaxLen masks, x2(n + 1600) = parity(mask & seed)
cudaError_t demod_descramble_launch(const FeCodeword* cws, uint32_t n_cw, uint32_t max_llr, const float* symbols,
                                    int16_t* e, const uint32_t* x1, const uint32_t* x2mask, cudaStream_t s);
cudaError_t rm_rx_sym_launch(const FeCodeword* cws, const float* symbols, int16_t* work, const uint16_t* tab_pool,
                             const RmSymItem* items, uint32_t n_items, const uint32_t* x1, const uint32_t* x2mask,
                             cudaStream_t s);

// TX mirror: turbo encoder + rate matching of one code block per item (frontend_kernels.cu)
struct TxItem {
  uint32_t K, f1, f2;
  uint32_t E;         // rate-matched bits to produce
  uint32_t tab_off;   // offset of the natural-order (K, rv) selection table in the table pool
  uint32_t pad;
  uint64_t bits_off;  // first input bit (one per byte)
  uint64_t e_off;     // first output bit (one per byte)
};
cudaError_t tcod_rm_tx_launch(const TxItem* items, uint32_t n_items, const uint8_t* bits, uint8_t* e,
                              const uint16_t* tab_pool, cudaStream_t s);

cudaError_t upload_crc_tables();  // fills the __constant__ CRC tables (once per context)

}  // namespace b200
