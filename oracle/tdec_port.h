/*
 * oracle/tdec_port.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement ("port") of the 16-bit LTE turbo-decode receive tail of
 * adjacentlink/srsLTE-emane, written from the behavioural description of the
 * reference and pinned against the reference's own compiled
 * code (oracle/_ref, built by oracle/Makefile from /root/reference) and against
 * the golden vectors under tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.  The product path (srslte-emane_b200/) never does.
 *
 * Reference anchors (paths relative to /root/reference):
 *   code-block sizes / segmentation   lib/src/phy/fec/cbsegm.c:30-135
 *   QPP interleaver                    lib/src/phy/fec/tc_interl_lte.c:38-113
 *   AUTO decoder selection             lib/src/phy/fec/turbodecoder.c:394-420
 *   half-iteration controller          lib/include/srslte/phy/fec/turbodecoder_iter.h:68-142
 *   window max-log-MAP (W=8/16)        lib/include/srslte/phy/fec/turbodecoder_win.h:332-679
 *   generic max-log-MAP (K<=400)       lib/src/phy/fec/turbodecoder_gen.c:54-269
 *   rate de-matching                   lib/src/phy/fec/rm_turbo.c:160-260,374-426
 *   CRC                                lib/src/phy/fec/crc.c:38-153
 *   transport-block loop               lib/src/phy/phch/sch.c:299-500
 */
#ifndef TDEC_PORT_H
#define TDEC_PORT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PORT_MAX_K        6144
#define PORT_NOF_K        188
#define PORT_SOFTBUF_LEN  18600 /* int16 per code block, softbuffer.h:50 */
#define PORT_CRC24A       0x1864CFBu
#define PORT_CRC24B       0x1800063u

/* ---- tables -------------------------------------------------------------------- */
int      port_cb_size(uint32_t idx);            /* K of table index, -1 if idx >= 188          */
int      port_cb_index(uint32_t long_cb);       /* smallest idx with K >= long_cb, -1 if none  */
int      port_cb_is_valid(uint32_t K);
int      port_nof_subblocks(uint32_t K);        /* 16, 8 or 0 (generic) -- AVX2 AUTO rule      */
int      port_qpp_params(uint32_t K, uint32_t* f1, uint32_t* f2);
uint32_t port_qpp(uint32_t K, uint32_t i);      /* pi(i) = (f1 i + f2 i^2) mod K               */

/* ---- CRC ------------------------------------------------------------------------ */
uint32_t port_crc_bytes(uint32_t poly, const uint8_t* data, uint32_t nbits); /* crc.c:139-153 */
uint32_t port_crc_bits(uint32_t poly, const uint8_t* bits, uint32_t nbits);  /* crc.c:98-136  */

/* ---- segmentation ---------------------------------------------------------------- */
typedef struct {
  uint32_t F, C, K1, K2, K1_idx, K2_idx, C1, C2, tbs;
} port_cbsegm_t;
int port_cbsegm(port_cbsegm_t* s, uint32_t tbs);

/* ---- rate de-matching ------------------------------------------------------------ */
/* table[i] = soft-buffer index that receives rate-matched sample i (i < 3K+12).
 * sb_layout != 0 selects the sub-block layout the decoder of this K expects.        */
int port_rm_rx_table(uint32_t K, uint32_t rv, int sb_layout, uint16_t* table);
/* softbuf[table[i mod N]] += e[i] (wrapping int16), i < E                           */
int port_rm_turbo_rx(const int16_t* e, uint32_t E, int16_t* softbuf, uint32_t K, uint32_t rv,
                     int sb_layout);

/* ---- decoder ---------------------------------------------------------------------- */
typedef struct port_tdec port_tdec_t;
port_tdec_t* port_tdec_new(void);
void         port_tdec_free(port_tdec_t* h);
int          port_tdec_new_cb(port_tdec_t* h, uint32_t K);
/* one HALF iteration + hard decision.  natural != 0: input is in[3i+j] order, otherwise
 * it is the sub-block layout produced by rate de-matching for this K.                */
void port_tdec_iteration(port_tdec_t* h, const int16_t* input, int natural, uint8_t* out);
int  port_tdec_run_all(port_tdec_t* h, const int16_t* input, int natural, uint8_t* out,
                       uint32_t nof_iterations, uint32_t K);
int  port_tdec_get_nof_iterations(const port_tdec_t* h);
/* statistics of the last new_cb..now span: number of saturating ops that clamped     */
uint64_t port_tdec_clamp_count(const port_tdec_t* h);
/* soft output of the last half iteration in natural order (K values)                 */
const int16_t* port_tdec_last_llr(const port_tdec_t* h);

/* ---- transport block -------------------------------------------------------------- */
typedef struct {
  uint32_t max_cb;
  int16_t* buffer_f; /* max_cb * PORT_SOFTBUF_LEN */
  uint8_t* data;     /* max_cb * 768              */
  uint8_t* cb_crc;   /* max_cb                    */
  uint8_t  tb_crc;
} port_softbuffer_t;

int  port_softbuffer_init(port_softbuffer_t* q, uint32_t max_cb);
void port_softbuffer_reset(port_softbuffer_t* q);
void port_softbuffer_free(port_softbuffer_t* q);

/* decode_tb semantics (sch.c:429-500).  Returns 0 ok, -1 CRC failure, -2 bad args.
 * cb_noi (nullable) receives the number of half iterations run per code block
 * (0 for blocks skipped because cb_crc was already set).                            */
int port_decode_tb(port_tdec_t* dec, port_softbuffer_t* sb, uint32_t tbs, uint32_t Qm, uint32_t rv,
                   uint32_t nof_e_bits, const int16_t* e_bits, uint8_t* data,
                   uint32_t max_iterations, float* avg_iterations, uint32_t* cb_noi);

/* batch helper for CPU baselines: N blocks of the same K, natural or sb input, fixed
 * number of half iterations, `threads` pthreads over disjoint ranges.                */
int port_batch_run_all(const int16_t* in, uint32_t in_stride, int natural, uint8_t* out,
                       uint32_t out_stride, uint32_t n, uint32_t K, uint32_t nof_iterations,
                       uint32_t threads);

/* ---- front end: soft demodulation + descrambling ---------------------- */
/* 36.211 7.2 Gold sequence (sequence.c:46-75) */
void port_gold_sequence(uint32_t seed, uint32_t len, uint8_t* c);
/* srslte_demod_soft_demodulate_s of the AVX2/SSE build (demod_soft.c:503-525); qm = 2, 4, 6, 8 */
int  port_demod_s(int qm, const float* sym, int16_t* llr, uint32_t nsym);
/* srslte_scrambling_s_offset (scrambling.c:44-47) */
void port_descramble_s(int16_t* llr, const uint8_t* c, uint32_t len);
/* the two in the order of pdsch.c:760-779 / pusch.c:482-500 */
/* UL-SCH with multiplexed UCI, data path only (sch.c:920-1064, uci.c:266-283, 497-571) */
uint32_t port_uci_q_prime_ri_ack(uint32_t O, uint32_t K_segm, uint32_t L_prb, uint32_t nof_symb, float beta);
uint32_t port_uci_q_prime_cqi(uint32_t O, uint32_t K_segm, uint32_t L_prb, uint32_t nof_symb, float beta, uint32_t q_prime_ri);
int port_ulsch_demux(const int16_t* q, const uint8_t* c_seq, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs,
                     uint32_t q_prime_ack, uint32_t q_prime_ri, uint32_t ri_len, int16_t* g);
/* ulsch_deinterleave without UCI (sch.c:580-598, 891-918) */
void port_ulsch_deinterleave(const int16_t* q, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs, int16_t* g);
int  port_demod_descramble(int qm, const float* sym, uint32_t nsym, uint32_t c_init, uint32_t nof_bits, int16_t* llr);

#ifdef __cplusplus
}
#endif
#endif
