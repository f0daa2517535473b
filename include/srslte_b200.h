/*
 * srslte_b200.h -- C ABI of the B200-native LTE turbo-decode path.
 *
 * This is the ONE added, batched entry of the drop-in: the reference's
 * per-code-block calls are replaced by calls over arrays of independent code blocks.  The
 * unchanged legacy entry points (srslte_tdec_*, srslte_rm_turbo_rx_lut*, ...) are declared in
 * srslte_b200_compat.h and are thin batch-of-one wrappers over the functions below.
 *
 * Plain C types only; every pointer is either a host pointer or a CUDA device pointer as stated.
 * All functions return 0 on success, SRSLTE_B200_ERROR (-1) on a runtime/CUDA failure and
 * SRSLTE_B200_ERROR_INVALID_INPUTS (-2) on bad arguments -- the reference's convention
 * (lib/include/srslte/config.h:56-63).  There is no CPU fallback: without a CUDA device the
 * context cannot be created and every entry fails.
 *
 * Reference interfaces replaced (paths under the reference tree):
 *   srslte_tdec_run_all / srslte_tdec_iteration      lib/include/srslte/phy/fec/turbodecoder.h:121-135
 *                                                    lib/src/phy/fec/turbodecoder.c:539-562
 *   srslte_rm_turbo_rx_lut                           lib/include/srslte/phy/fec/rm_turbo.h:76-80
 *                                                    lib/src/phy/fec/rm_turbo.c:374-426
 *   decode_tb_cb / decode_tb (per-CB loop, CRC early stop)   lib/src/phy/phch/sch.c:299-500
 *   srslte_crc_checksum_byte                         lib/src/phy/fec/crc.c:139-153
 */
#ifndef SRSLTE_B200_H
#define SRSLTE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRSLTE_B200_SUCCESS 0
#define SRSLTE_B200_ERROR (-1)
#define SRSLTE_B200_ERROR_INVALID_INPUTS (-2)

#define SRSLTE_B200_API __attribute__((visibility("default")))

typedef struct srslte_b200_ctx srslte_b200_ctx_t;

/* ---- context ------------------------------------------------------------------------------- */
/* One context per host thread / GPU (the reference's handles are single-threaded too:
 * one srslte_sch_t per worker, SURVEY.md 8b "Threading").                                       */
SRSLTE_B200_API int  srslte_b200_ctx_create(srslte_b200_ctx_t** ctx, int cuda_device);
SRSLTE_B200_API void srslte_b200_ctx_destroy(srslte_b200_ctx_t* ctx);
/* Kernels of the *_dev entries are enqueued on this CUDA stream (cudaStream_t passed as void*;
 * NULL = the context's own stream).                                                             */
SRSLTE_B200_API int  srslte_b200_ctx_set_stream(srslte_b200_ctx_t* ctx, void* cuda_stream);
SRSLTE_B200_API int  srslte_b200_ctx_synchronize(srslte_b200_ctx_t* ctx);
SRSLTE_B200_API const char* srslte_b200_last_error(const srslte_b200_ctx_t* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches).                 */
SRSLTE_B200_API uint64_t srslte_b200_launch_count(const srslte_b200_ctx_t* ctx);

/* Optional per-kernel timing for measurement (bench.py): when enabled, every kernel launch of the
 * context is bracketed by CUDA events on its own stream.  kind: 0 = 16-window decoder, 1 = 8-window
 * decoder, 2 = generic decoder, 3 = natural->working layout, 4 = rate de-matching.
 * _kernel_time synchronizes the device and returns the summed duration and the launch count since
 * timing was (re-)enabled.                                                                       */
SRSLTE_B200_API int srslte_b200_ctx_enable_timing(srslte_b200_ctx_t* ctx, int enable);
SRSLTE_B200_API int srslte_b200_ctx_kernel_time(srslte_b200_ctx_t* ctx, int kind, double* total_ms,
                                                uint32_t* launches);

/* The window decoders run every half iteration first with wrapping int16x2 arithmetic plus a proof
 * that none of the reference's saturating operations could have clamped, and re-run it with exact
 * saturating arithmetic when the proof fails (both on the GPU, results identical by construction).
 * _set_exact(1) forces the exact variant (tests); _fallback_count returns how many (warp, half
 * iteration) pairs have taken the exact re-run so far (synchronizes the device).                 */
SRSLTE_B200_API int srslte_b200_ctx_set_exact(srslte_b200_ctx_t* ctx, int force_exact);
/* Tests and measurements: which of the (bit-identical) variants of the window decoders may run.  Bit 1 / 2: skip the
 * pure / static tier; 3: general path only; 5: the kernels with the tracked tier of the main path; 6: CRC modes through
 * the round-based kernel instead of the block-granular early-termination kernel; 7: launches without CRC through the
 * kernel that also carries the CRC variants; 9: the early-termination kernel refills a thread group at once instead of
 * waiting for its warp's next DEC1 half iteration.  0 = default.  Results never differ.                              */
SRSLTE_B200_API int srslte_b200_ctx_set_variant_bits(srslte_b200_ctx_t* ctx, uint32_t bits);
SRSLTE_B200_API int srslte_b200_ctx_fallback_count(srslte_b200_ctx_t* ctx, uint64_t* count);
/* (warp, half iteration) pairs of the window decoders so far in the pure / static / tracked / exact variant
 * (synchronizes the device)                                                        */
SRSLTE_B200_API int srslte_b200_ctx_tier_counts(srslte_b200_ctx_t* ctx, uint64_t counts[4]);

/* Pinned host memory for the *_host entries (plain cudaHostAlloc; any host pointer is accepted,
 * pinned ones are copied without staging).                                                      */
SRSLTE_B200_API void* srslte_b200_host_alloc(size_t bytes);
SRSLTE_B200_API void  srslte_b200_host_free(void* p);

/* ---- tables (host, no GPU needed) ---------------------------------------------------------- */
/* srslte_cbsegm_cbindex / _cbsize (cbsegm.c:109-135)                                            */
SRSLTE_B200_API int srslte_b200_cb_index(uint32_t long_cb);
SRSLTE_B200_API int srslte_b200_cb_size(uint32_t index);
/* srslte_tdec_autoimp_get_subblocks (turbodecoder.c:394-406): 16, 8 or 0                        */
SRSLTE_B200_API uint32_t srslte_b200_nof_windows(uint32_t long_cb);
/* int16 elements of one code block in the decoder's working layout (= what
 * srslte_rm_turbo_rx_lut writes): 3*(K+32)+12 for window decoders, 3*K+12 for K <= 400.         */
SRSLTE_B200_API uint32_t srslte_b200_working_len(uint32_t long_cb);
/* The receive index table of (K, rv): table[i] = working-layout (sb_layout != 0) or natural
 * (sb_layout == 0) index accumulating rate-matched sample i, i < 3K+12 (rm_turbo.c:160-260).    */
SRSLTE_B200_API int srslte_b200_rm_rx_table(uint32_t long_cb, uint32_t rv, int sb_layout, uint16_t* table);

/* ---- batched turbo decode ------------------------------------------------------------------ */
#define SRSLTE_B200_INPUT_NATURAL 0 /* in[3i+j], tails last: srslte_tdec_force_not_sb semantics   */
#define SRSLTE_B200_INPUT_WORKING 1 /* sub-block soft-buffer layout written by rate de-matching   */

#define SRSLTE_B200_CRC_NONE 0 /* exactly max(1, nof_iterations) half iterations: srslte_tdec_run_all */
#define SRSLTE_B200_CRC_24B 1  /* stop a block when CRC24B over its K bits is 0: sch.c:365-378, C > 1 */
#define SRSLTE_B200_CRC_24A 2  /* same with CRC24A: single-code-block transport block              */

typedef struct {
  uint32_t        n_cb;           /* number of independent code blocks                            */
  const uint32_t* long_cb;        /* host array [n_cb] of K, or NULL when all blocks have uniform_long_cb */
  uint32_t        uniform_long_cb;
  uint32_t        input_format;   /* SRSLTE_B200_INPUT_*                                           */
  uint32_t        in_stride;      /* int16 elements between consecutive blocks' inputs (even)      */
  uint32_t        out_stride;     /* bytes between consecutive blocks' outputs (>= K/8)            */
  uint32_t        nof_iterations; /* cap, in srsLTE HALF iterations (enb.conf.example:184)         */
  uint32_t        crc_mode;       /* SRSLTE_B200_CRC_*                                             */
} srslte_b200_tdec_batch_t;

/* Device-resident variant: llr, out, n_iter, crc_ok are CUDA device pointers; work is enqueued on
 * the context's stream and the call returns without waiting.
 *   llr    [n_cb * in_stride] int16
 *   out    [n_cb * out_stride] decoded bytes, MSB first (K/8 per block)
 *   n_iter [n_cb] half iterations actually run per block (nullable)
 *   crc_ok [n_cb] 1 if the block's CRC check passed (nullable; 0 in CRC_NONE mode)               */
SRSLTE_B200_API int srslte_b200_tdec_batch_dev(srslte_b200_ctx_t* ctx, const srslte_b200_tdec_batch_t* batch,
                                               const int16_t* llr, uint8_t* out, uint8_t* n_iter,
                                               uint8_t* crc_ok);

/* Host variant: same arguments as host pointers; copies are pipelined with the kernels; returns
 * when the results are in host memory (the reference's calls are synchronous).                  */
SRSLTE_B200_API int srslte_b200_tdec_batch_host(srslte_b200_ctx_t* ctx, const srslte_b200_tdec_batch_t* batch,
                                                const int16_t* llr, uint8_t* out, uint8_t* n_iter,
                                                uint8_t* crc_ok);

/* ---- several GPUs of one box from ONE process ------------------------------------------------------
 * Code blocks are independent, so a batch is cut into contiguous shards (block i -> device floor(i * n_devices / n_cb),
 * like sharding.py) and every shard runs through srslte_b200_tdec_batch_host on its own device from its own host
 * thread (bound to the CPUs next to that GPU): no inter-device traffic, no collective.  The reference's unit of
 * concurrency is one worker thread per srslte_sch_t (srsenb/src/phy/sf_worker.cc:579-650); a group is what such
 * workers -- or one caller with one large batch -- share.  devices == NULL: devices 0 .. n_devices-1.            */
typedef struct srslte_b200_group srslte_b200_group_t;
SRSLTE_B200_API int  srslte_b200_group_create(srslte_b200_group_t** group, const int* devices, uint32_t n_devices);
SRSLTE_B200_API void srslte_b200_group_destroy(srslte_b200_group_t* group);
SRSLTE_B200_API uint32_t srslte_b200_group_size(const srslte_b200_group_t* group);
SRSLTE_B200_API srslte_b200_ctx_t* srslte_b200_group_ctx(srslte_b200_group_t* group, uint32_t index);
SRSLTE_B200_API int  srslte_b200_group_tdec_batch_host(srslte_b200_group_t* group, const srslte_b200_tdec_batch_t* batch,
                                                       const int16_t* llr, uint8_t* out, uint8_t* n_iter, uint8_t* crc_ok);
/* Host-to-device ceiling of the box: every device of the group copies `bytes_per_device` from host + i * bytes_per_device
 * at the same time, `reps` times, no kernels; gbs_per_device[i] = GB/s seen by device i while all of them copy.   */
SRSLTE_B200_API int  srslte_b200_group_h2d_probe(srslte_b200_group_t* group, const void* host, size_t bytes_per_device,
                                                 uint32_t reps, double* gbs_per_device);
/* Shares of a batch.  The host-facing entry is bound by the host-to-device copies, and the PCIe side of a box need not
 * be symmetric (8 x B200: 24 GB/s per device on one half, 35 on the other when all copy, profiles/r02_multi_gpu.txt).
 * set_weights: device i takes weights[i] / sum of a batch (NULL: equal shares, the default); calibrate: measure the
 * devices' concurrent copy rates (a few ms) and use them as weights (gbs_per_device, nullable, receives them).      */
SRSLTE_B200_API int  srslte_b200_group_set_weights(srslte_b200_group_t* group, const double* weights);
SRSLTE_B200_API int  srslte_b200_group_calibrate(srslte_b200_group_t* group, double* gbs_per_device);
/* the same for one context (bench.py under torchrun: every rank probes after a barrier)                            */
SRSLTE_B200_API int  srslte_b200_h2d_probe(srslte_b200_ctx_t* ctx, const void* host, size_t bytes, uint32_t reps, double* gbs);


/* ---- batched rate de-matching -------------------------------------------------------------- */
typedef struct {
  uint32_t long_cb;   /* K                                                                         */
  uint32_t rv;        /* redundancy version 0..3                                                   */
  uint32_t e_offset;  /* first sample of this block in e (int16 elements)                          */
  uint32_t e_len;     /* number of rate-matched samples E (may exceed 3K+12: wraps)                */
  uint32_t work_offset; /* start of the block's working/soft buffer (int16 elements, even)         */
} srslte_b200_rm_block_t;

/* work[table[i mod (3K+12)]] += e[i] (wrapping int16) for every block: HARQ combining happens in
 * place, the caller zeroes `work` for a new transmission (srslte_softbuffer_rx_reset).
 * e and work are device pointers; blocks is a host array.                                        */
SRSLTE_B200_API int srslte_b200_rm_rx_batch_dev(srslte_b200_ctx_t* ctx, const srslte_b200_rm_block_t* blocks,
                                                uint32_t n_blocks, const int16_t* e, int16_t* work);

/* ---- front end: soft demodulation + descrambling --------------------------- */
/* Replaces, for many codewords per call, the two steps in front of srslte_dlsch_decode2 / srslte_ulsch_decode:
 *   srslte_demod_soft_demodulate_s(mod, symbols, e, nof_re)          lib/src/phy/modem/demod_soft.c:503-525
 *   srslte_scrambling_s_offset(seq, e, 0, nof_bits)                  lib/src/phy/scrambling/scrambling.c:44-47
 * (call sites lib/src/phy/phch/pdsch.c:760-779, pusch.c:482-500), seq = srslte_sequence_LTE_pr(len, c_init)
 * (lib/src/phy/common/sequence.c:123-136).  Bit-exact with the reference's AVX2/SSE build while
 * |scale * x| < 32768 (the reference's out-of-range float -> short conversions are undefined behaviour).      */
/* UCI multiplexed with the UL-SCH data (36.212 5.2.2.6 - 5.2.2.8), as far as it shapes the DATA path of
 * srslte_ulsch_decode (sch.c:920-1064).  Decoding the ACK / RI / CQI values themselves (uci.c) is control-plane work
 * and stays with the caller: it reads the LLRs of the plain call (ul_nof_symb = 0) at the positions of
 * uci_ulsch_interleave_ack_gen / _ri_gen, and the first Q'_cqi * qm entries of the de-multiplexed output.
 *   q_prime_ack  coded HARQ-ACK symbols: they puncture the data, their LLRs are erased (sch.c:961-964)
 *   q_prime_ri   coded RI symbols: the de-interleaver skips them (ulsch_interleave_gen, sch.c:580-598).  The reference
 *                scatters every RI sample to g[0] (lut = 0 through srslte_vec_lut_sis, sch.c:589-590, 910), so g[0]
 *                ends up holding the RI sample with the highest channel position; that is reproduced
 *   q_prime_cqi  coded CQI symbols at the head of the UL-SCH order: the data starts q_prime_cqi * qm LLRs in
 *   ri_len       1 or 2 RI payload bits (cfg->uci_cfg.cqi.ri_len); the 1-bit RI decoder flips one LLR per RI symbol
 *                back in place (decode_ri_ack_1bit, uci.c:627-628), which is visible through g[0] for QPSK
 * The counts are srslte_b200_uci_q_prime_ri_ack / _cqi below (= Q_prime_ri_ack / Q_prime_cqi, uci.c:266-283, 547-571). */
typedef struct {
  uint32_t q_prime_ack, q_prime_ri, q_prime_cqi, ri_len;
} srslte_b200_ul_uci_t;
/* O payload bits, K_segm = C1*K1 + C2*K2 of the transport block (> 0), beta from 36.213 tables 8.6.3-1/2/3 */
SRSLTE_B200_API uint32_t srslte_b200_uci_q_prime_ri_ack(uint32_t O, uint32_t K_segm, uint32_t L_prb, uint32_t nof_symb,
                                                        float beta);
SRSLTE_B200_API uint32_t srslte_b200_uci_q_prime_cqi(uint32_t O, uint32_t K_segm, uint32_t L_prb, uint32_t nof_symb,
                                                     float beta, uint32_t q_prime_ri);

typedef struct {
  uint32_t qm;          /* bits per symbol: 2 QPSK, 4 16QAM, 6 64QAM, 8 256QAM (srslte_mod_t 1..4)                 */
  uint32_t nof_symbols; /* cfg->grant.nof_re                                                                        */
  uint32_t c_init;      /* scrambling seed: (rnti << 14) + (q << 13) + ((nslot / 2) << 9) + cell_id for PDSCH       */
  uint32_t nof_bits;    /* LLRs that are descrambled: grant.tb[].nof_bits, <= qm * nof_symbols, <= 262144          */
  uint64_t sym_offset;  /* first symbol of the codeword in `symbols` (complex floats: re, im)                       */
  uint64_t llr_offset;  /* first LLR of the codeword in `e` (int16); unused by the fused entry                      */
                        /* Any offsets work.  Codewords that start on 128-bit boundaries (sym_offset even, llr_offset a
                         * multiple of 8, buffers from cudaMalloc) take the kernels' vector path: 8 or 24 LLRs per thread,
                         * about three times the rate of the one-LLR-at-a-time path.             */
  uint32_t ul_nof_symb; /* 0: PDSCH.  PUSCH: cfg->grant.nof_symb (N_pusch_symbs): the outputs are additionally put in
                         * UL-SCH order by the channel de-interleaver of 36.212 5.2.2.8 (ulsch_deinterleave,
                         * sch.c:891-918); nof_bits must be qm * nof_symbols and a multiple of qm * ul_nof_symb     */
  uint32_t reserved;    /* 0 */
  srslte_b200_ul_uci_t uci; /* control information multiplexed into the PUSCH codeword; all 0: none                */
} srslte_b200_codeword_t;

/* e[cw.llr_offset + j] = descrambled LLR j of codeword cw, j < qm * nof_symbols.  symbols, e: device memory.
 * With ul_nof_symb and uci: e is the reference's g_bits array after srslte_ulsch_decode's de-multiplexing
 * (sch.c:1028-1035): (H' - Q'_ri) * qm entries, the CQI LLRs first (Q'_cqi * qm of them, as received), then the data
 * with the ACK positions erased; the remaining Q'_ri * qm entries (stale memory in the reference) are 0.        */
SRSLTE_B200_API int srslte_b200_demod_descramble_dev(srslte_b200_ctx_t* ctx, const srslte_b200_codeword_t* cws,
                                                     uint32_t n_cw, const float* symbols, int16_t* e);

/* Fused with rate de-matching (srslte_rm_turbo_rx_lut semantics, HARQ combining in place): the e array of the
 * reference is never written.  work[blk.work_offset + table(K, rv)[i mod (3K+12)]] += LLR(blk.codeword,
 * blk.e_offset + i) for i < e_len.  symbols, work: device memory. */
typedef struct {
  uint32_t long_cb;     /* K */
  uint32_t rv;
  uint32_t codeword;    /* index into cws[] */
  uint32_t e_offset;    /* first rate-matched LLR of this block inside its codeword (rp of sch.c:324-334) */
  uint32_t e_len;       /* E */
  uint32_t work_offset; /* int16 offset of the block's working buffer in `work` */
} srslte_b200_rm_sym_block_t;
SRSLTE_B200_API int srslte_b200_demod_rm_rx_batch_dev(srslte_b200_ctx_t* ctx, const srslte_b200_codeword_t* cws,
                                                      uint32_t n_cw, const srslte_b200_rm_sym_block_t* blocks,
                                                      uint32_t n_blocks, const float* symbols, int16_t* work);

/* ---- TX mirror: turbo encoder + rate matching ---------------------------------- */
/* For generating test / benchmark vectors on the device and as the encode half of a batched DL path:
 *   srslte_tcod_encode(bits, d, K)                 lib/src/phy/fec/turbocoder.c:95-187   (d = 3K + 12 bits, 3i+j order)
 *   srslte_rm_turbo_tx(..., d, 3K+12, e, E, rv)    lib/src/phy/fec/rm_turbo.c:303-372    (circular-buffer selection)
 * bits / e: device memory, one bit per byte (0 / 1). */
typedef struct {
  uint32_t long_cb;      /* K: input bits of this block (its CRC already attached) */
  uint32_t rv;
  uint32_t e_len;        /* E */
  uint32_t reserved;     /* 0 */
  uint64_t bits_offset;  /* first input bit of the block in `bits` */
  uint64_t e_offset;     /* first rate-matched bit of the block in `e` */
} srslte_b200_tx_block_t;
SRSLTE_B200_API int srslte_b200_tcod_rm_tx_batch_dev(srslte_b200_ctx_t* ctx, const srslte_b200_tx_block_t* blocks,
                                                     uint32_t n_blocks, const uint8_t* bits, uint8_t* e);

/* ---- batched transport-block decode (the sch.c decode_tb loop over many TBs) ----------------- */
/* HARQ state lives on the device: one "soft buffer" = max_cb code-block LLR buffers + CRC flags +
 * saved payloads, the counterpart of srslte_softbuffer_rx_t (softbuffer.h:37-43, softbuffer.c:41-150). */
typedef struct srslte_b200_harq_pool srslte_b200_harq_pool_t;
SRSLTE_B200_API int  srslte_b200_harq_pool_create(srslte_b200_ctx_t* ctx, uint32_t n_softbuffers, uint32_t max_cb,
                                                  srslte_b200_harq_pool_t** pool);
SRSLTE_B200_API void srslte_b200_harq_pool_destroy(srslte_b200_ctx_t* ctx, srslte_b200_harq_pool_t* pool);
/* srslte_softbuffer_rx_reset: zero the LLRs, clear cb_crc / tb_crc and the saved payloads of one soft buffer */
SRSLTE_B200_API int  srslte_b200_harq_reset(srslte_b200_ctx_t* ctx, srslte_b200_harq_pool_t* pool, uint32_t softbuffer);
/* the same for n soft buffers in one call (softbuffers == NULL and n == the pool's size: all of them) */
SRSLTE_B200_API int  srslte_b200_harq_reset_many(srslte_b200_ctx_t* ctx, srslte_b200_harq_pool_t* pool,
                                                 const uint32_t* softbuffers, uint32_t n);
/* copy of softbuffer->cb_crc[0..n) */
SRSLTE_B200_API int  srslte_b200_harq_cb_crc(srslte_b200_harq_pool_t* pool, uint32_t softbuffer, uint8_t* cb_crc,
                                             uint32_t n);

typedef struct {
  uint32_t       tbs;            /* transport block size in bits (without its CRC24A)                  */
  uint32_t       qm;             /* bits per symbol x layers, as decode_tb receives it (sch.c:528)     */
  uint32_t       rv;             /* redundancy version                                                 */
  uint32_t       nof_e_bits;     /* G: rate-matched soft bits of this TB                               */
  uint32_t       softbuffer;     /* soft buffer of the pool (HARQ process) to combine into             */
  const int16_t* e_bits;         /* host: nof_e_bits int16 LLRs                                        */
  uint8_t*       data;           /* host: decoded TB, at least tbs/8 + 6 bytes                         */
  int32_t        ret;            /* out: 0 ok, -1 CRC failure, -2 invalid arguments (decode_tb's value) */
  float          avg_iterations; /* out: srslte_sch_last_noi                                           */
} srslte_b200_tb_t;

/* decode_tb (sch.c:429-500) for n_tb independent transport blocks in one call: rate de-matching into the
 * soft buffers, all code blocks of all TBs decoded in one batch with CRC24B/24A early termination,
 * block -> TB assembly, TB CRC24A, HARQ bookkeeping.  Returns 0 when the batch ran; per-TB results in tbs[]. */
SRSLTE_B200_API int srslte_b200_decode_tb_batch(srslte_b200_ctx_t* ctx, srslte_b200_harq_pool_t* pool,
                                                srslte_b200_tb_t* tbs, uint32_t n_tb, uint32_t max_iterations);

/* The same fed with equalised symbols instead of LLRs: soft demodulation, descrambling and rate de-matching run
 * fused on the device (pdsch.c:760-781: srslte_demod_soft_demodulate_s, srslte_scrambling_s_offset,
 * srslte_dlsch_decode2), and 8 bytes per resource element cross PCIe instead of 2 * Qm. */
typedef struct {
  uint32_t     tbs;            /* transport block size in bits                                        */
  uint32_t     qm;             /* modulation bits per symbol: 2, 4, 6, 8 (one layer)                  */
  uint32_t     rv;
  uint32_t     nof_e_bits;     /* grant.tb[].nof_bits (<= qm * nof_symbols): descrambled; with uci the decoded
                                * data bits are nof_e_bits - (q_prime_ri + q_prime_cqi) * qm (sch.c:1061-1062) */
  uint32_t     softbuffer;
  uint32_t     nof_symbols;    /* grant.nof_re                                                        */
  uint32_t     c_init;         /* scrambling seed of (rnti, codeword, subframe, cell)                 */
  uint32_t     ul_nof_symb;    /* 0 for PDSCH; PUSCH: N_pusch_symbs (UL-SCH de-interleaver)           */
  srslte_b200_ul_uci_t uci;    /* PUSCH: multiplexed control information, all 0 for none              */
  const float* symbols;        /* host: nof_symbols complex floats (re, im)                           */
  uint8_t*     data;           /* host: decoded TB, at least tbs/8 + 6 bytes                          */
  int32_t      ret;            /* out: as srslte_b200_tb_t                                            */
  float        avg_iterations; /* out                                                                 */
} srslte_b200_tb_sym_t;
SRSLTE_B200_API int srslte_b200_decode_tb_sym_batch(srslte_b200_ctx_t* ctx, srslte_b200_harq_pool_t* pool,
                                                    srslte_b200_tb_sym_t* tbs, uint32_t n_tb, uint32_t max_iterations);

#ifdef __cplusplus
}
#endif
#endif /* SRSLTE_B200_H */
