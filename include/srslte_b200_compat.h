/*
 * srslte_b200_compat.h -- the reference's own entry points for this path, exported unchanged by
 * libsrslte_b200.so so that code written against srsLTE's headers relinks without source changes.
 *
 * Every function below keeps the reference's name, argument meaning, return convention and struct
 * layout; each is a thin batch-of-one wrapper over the batched C ABI in srslte_b200.h and runs on the GPU
 * (device chosen by the environment variable SRSLTE_B200_DEVICE, default 0).  There is no CPU path:
 * *_init fails with -1 and prints to stderr when no CUDA device is usable.
 *
 * Struct layouts are restated here (not copied) and are checked against the compiled reference by
 * tests/test_compat_abi.py (sizeof / offsetof through oracle/_ref).
 *
 * Reference declarations mirrored (paths under the reference tree, lib/include/srslte/phy/...):
 *   fec/turbodecoder.h:63-135      srslte_tdec_t and the srslte_tdec_* functions
 *   fec/turbodecoder_impl.h:28-40  srslte_tdec_impl_type_t
 *   fec/tc_interl.h:36-40          srslte_tc_interl_t
 *   fec/rm_turbo.h:55-93           srslte_rm_turbo_gentables / free_tables / rx_lut / rx_lut_
 *   fec/softbuffer.h:37-43         srslte_softbuffer_rx_t
 *   phch/sch.c:429-500             decode_tb (static in the reference; exported here as
 *                                  srslte_b200_sch_decode_tb for the 10-line patch in INTEGRATION.md)
 *
 * Not provided (out of scope): the experimental 8-bit decoders.  srslte_tdec_*_8bit and
 * srslte_rm_turbo_rx_lut_8bit exist so that callers link, print an error and fail.
 */
#ifndef SRSLTE_B200_COMPAT_H
#define SRSLTE_B200_COMPAT_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef SRSLTE_API
#define SRSLTE_API __attribute__((visibility("default")))
#endif

#ifndef SRSLTE_SUCCESS
#define SRSLTE_SUCCESS 0
#define SRSLTE_ERROR -1
#define SRSLTE_ERROR_INVALID_INPUTS -2
#endif

#define SRSLTE_NOF_TC_CB_SIZES 188
#define SRSLTE_TCOD_MAX_LEN_CB 6144
#define SRSLTE_TDEC_NOF_AUTO_MODES_8 2
#define SRSLTE_TDEC_NOF_AUTO_MODES_16 3
#define SOFTBUFFER_SIZE 18600

typedef enum {
  SRSLTE_TDEC_AUTO = 0,
  SRSLTE_TDEC_GENERIC,
  SRSLTE_TDEC_SSE,
  SRSLTE_TDEC_SSE_WINDOW,
  SRSLTE_TDEC_NEON_WINDOW,
  SRSLTE_TDEC_AVX_WINDOW,
  SRSLTE_TDEC_SSE8_WINDOW,
  SRSLTE_TDEC_AVX8_WINDOW,
  SRSLTE_TDEC_NOF_IMP
} srslte_tdec_impl_type_t;

typedef enum { SRSLTE_TDEC_8, SRSLTE_TDEC_16 } srslte_tdec_llr_type_t;

typedef struct {
  uint16_t* forward;
  uint16_t* reverse;
  uint32_t  max_long_cb;
} srslte_tc_interl_t;

/* Same size and field offsets as the reference's handle (callers embed it by value inside srslte_sch_t).
 * This library keeps its GPU context in dec16_hdlr[0] and a pinned staging buffer in input_conv; the work
 * arrays app1..parity1 and the interleaver tables stay NULL (the GPU computes the QPP on the fly).          */
typedef struct {
  uint32_t max_long_cb;
  void*    dec8_hdlr[SRSLTE_TDEC_NOF_AUTO_MODES_8];
  void*    dec16_hdlr[SRSLTE_TDEC_NOF_AUTO_MODES_16];
  void*    dec8[SRSLTE_TDEC_NOF_AUTO_MODES_8];
  void*    dec16[SRSLTE_TDEC_NOF_AUTO_MODES_16];
  int      nof_blocks8[SRSLTE_TDEC_NOF_AUTO_MODES_8];
  int      nof_blocks16[SRSLTE_TDEC_NOF_AUTO_MODES_16];
  void*    app1;
  void*    app2;
  void*    ext1;
  void*    ext2;
  void*    syst0;
  void*    parity0;
  void*    parity1;
  void*    input_conv;
  bool     force_not_sb;
  srslte_tdec_impl_type_t dec_type;
  srslte_tdec_llr_type_t  current_llr_type;
  uint32_t current_dec;
  uint32_t current_long_cb;
  uint32_t current_inter_idx;
  int      current_cbidx;
  srslte_tc_interl_t interleaver[4][SRSLTE_NOF_TC_CB_SIZES];
  int      n_iter;
} srslte_tdec_t;

typedef struct {
  uint32_t  max_cb;
  int16_t** buffer_f;
  uint8_t** data;
  bool*     cb_crc;
  bool      tb_crc;
} srslte_softbuffer_rx_t;

/* ---- turbodecoder.h:97-135 ---- */
SRSLTE_API int      srslte_tdec_init(srslte_tdec_t* h, uint32_t max_long_cb);
SRSLTE_API int      srslte_tdec_init_manual(srslte_tdec_t* h, uint32_t max_long_cb, srslte_tdec_impl_type_t dec_type);
SRSLTE_API void     srslte_tdec_free(srslte_tdec_t* h);
SRSLTE_API void     srslte_tdec_force_not_sb(srslte_tdec_t* h);
SRSLTE_API int      srslte_tdec_new_cb(srslte_tdec_t* h, uint32_t long_cb);
SRSLTE_API int      srslte_tdec_get_nof_iterations(srslte_tdec_t* h);
SRSLTE_API uint32_t srslte_tdec_autoimp_get_subblocks(uint32_t long_cb);
SRSLTE_API uint32_t srslte_tdec_autoimp_get_subblocks_8bit(uint32_t long_cb);
SRSLTE_API void     srslte_tdec_iteration(srslte_tdec_t* h, int16_t* input, uint8_t* output);
SRSLTE_API int      srslte_tdec_run_all(srslte_tdec_t* h, int16_t* input, uint8_t* output, uint32_t nof_iterations,
                                        uint32_t long_cb);
SRSLTE_API void     srslte_tdec_iteration_8bit(srslte_tdec_t* h, int8_t* input, uint8_t* output);
SRSLTE_API int      srslte_tdec_run_all_8bit(srslte_tdec_t* h, int8_t* input, uint8_t* output, uint32_t nof_iterations,
                                             uint32_t long_cb);

/* ---- rm_turbo.h:55-93 (receive side) ---- */
SRSLTE_API void srslte_rm_turbo_gentables(void);
SRSLTE_API void srslte_rm_turbo_free_tables(void);
SRSLTE_API int  srslte_rm_turbo_rx_lut(int16_t* input, int16_t* output, uint32_t in_len, uint32_t cb_idx, uint32_t rv_idx);
SRSLTE_API int  srslte_rm_turbo_rx_lut_(int16_t* input, int16_t* output, uint32_t in_len, uint32_t cb_idx,
                                        uint32_t rv_idx, bool enable_input_tdec);
SRSLTE_API int  srslte_rm_turbo_rx_lut_8bit(int8_t* input, int8_t* output, uint32_t in_len, uint32_t cb_idx,
                                            uint32_t rv_idx);

/* ---- sch.c:429-500 decode_tb, with the host-resident soft buffer the MAC owns ---- */
/* Returns 0 / -1 (CRC failure) / -2 (bad arguments) like decode_tb; *avg_iterations = q->avg_iterations.   */
/* ---- softbuffer.h:52-66 (softbuffer.c:40-150): the receive soft buffer.  Same struct, same calls; a buffer made by
 * srslte_softbuffer_rx_init keeps its LLRs in a device pool, a reset is a flag, and
 * srslte_b200_sch_decode_tb / srslte_dlsch_decode2 no longer move the LLRs over PCIe.  The reference's softbuffer.c is
 * compiled with these five names defined out of the way (its transmit-side functions stay), see oracle/Makefile.   */
SRSLTE_API int  srslte_softbuffer_rx_init(srslte_softbuffer_rx_t* q, uint32_t nof_prb);
SRSLTE_API void srslte_softbuffer_rx_reset(srslte_softbuffer_rx_t* q);
SRSLTE_API void srslte_softbuffer_rx_reset_tbs(srslte_softbuffer_rx_t* q, uint32_t tbs);
SRSLTE_API void srslte_softbuffer_rx_reset_cb(srslte_softbuffer_rx_t* q, uint32_t nof_cb);
SRSLTE_API void srslte_softbuffer_rx_free(srslte_softbuffer_rx_t* q);

SRSLTE_API int srslte_b200_sch_decode_tb(srslte_softbuffer_rx_t* softbuffer, uint32_t tbs, uint32_t Qm, uint32_t rv,
                                         uint32_t nof_e_bits, int16_t* e_bits, uint8_t* data, uint32_t max_iterations,
                                         float* avg_iterations);

/* ---- sch.h:98-107 / sch.c:502-532: the DL-SCH decode entry points themselves ----------------------------------
 * srslte_dlsch_decode2 is what pdsch.c:786 (and through srslte_dlsch_decode, pmch.c:383) calls.  In the reference
 * it shares sch.c with the encoder and with srslte_sch_init / _free / _set_max_noi / _last_noi, which stay where
 * they are: the reference's sch.c is compiled UNCHANGED with two compile definitions
 *     -Dsrslte_dlsch_decode2=srslte_dlsch_decode2_cpu -Dsrslte_dlsch_decode=srslte_dlsch_decode_cpu
 * (its own two functions get out of the way) and linked with this library, whose srslte_tdec_init the unchanged
 * srslte_sch_init then calls for q->decoder.  The structures below restate the reference layouts these two
 * functions read (phch/ra.h:43-53, phch/pdsch_cfg.h:41-79, phch/sch.h:52-77; SRSLTE_MAX_PRB = 110,
 * SRSLTE_MAX_CODEWORDS = 2); tests/test_compat_abi.py checks sizes and offsets against the compiled reference.   */
typedef struct {
  uint32_t mod; /* srslte_mod_t: BPSK = 0, QPSK, 16QAM, 64QAM, 256QAM (phy_common.h:241-247) */
  int      tbs;
  int      rv;
  uint32_t nof_bits;
  uint32_t cw_idx;
  bool     enabled;
  uint32_t mcs_idx;
} srslte_ra_tb_t;

typedef struct {
  uint32_t       tx_scheme; /* srslte_tx_scheme_t */
  uint32_t       pmi;
  bool           prb_idx[2][110];
  uint32_t       nof_prb;
  uint32_t       nof_re;
  uint32_t       nof_symb_slot[2];
  srslte_ra_tb_t tb[2];
  int            last_tbs[2];
  uint32_t       nof_tb;
  uint32_t       nof_layers;
} srslte_pdsch_grant_t;

typedef struct {
  srslte_pdsch_grant_t grant;
  uint16_t             rnti;
  uint32_t             max_nof_iterations;
  uint32_t             decoder_type; /* srslte_mimo_decoder_t */
  float                p_a;
  uint32_t             p_b;
  float                rs_power;
  bool                 power_scale;
  bool                 csi_enable;
  bool                 use_tbs_index_alt;
  union {
    void*                   tx[2];
    srslte_softbuffer_rx_t* rx[2];
  } softbuffers;
  bool     meas_time_en;
  uint32_t meas_time_value;
} srslte_pdsch_cfg_t;

/* the head of srslte_sch_t (sch.h:52-77); buffers, encoder, decoder, CRC objects follow in the reference's struct */
typedef struct {
  uint32_t max_iterations;
  float    avg_iterations;
  bool     llr_is_8bit;
} srslte_sch_head_t;

SRSLTE_API int srslte_dlsch_decode(void* q /* srslte_sch_t* */, srslte_pdsch_cfg_t* cfg, int16_t* e_bits, uint8_t* data);
SRSLTE_API int srslte_dlsch_decode2(void* q /* srslte_sch_t* */, srslte_pdsch_cfg_t* cfg, int16_t* e_bits, uint8_t* data,
                                    int codeword_idx, uint32_t nof_layers);


/* ---- sch.h:109-115 / sch.c:920-1064: the UL-SCH decode entry point ------------------------------------------------
 * srslte_ulsch_decode is what pusch.c:503 calls.  Like the two DL entry points it shares sch.c with code that stays,
 * so the reference's sch.c is compiled UNCHANGED with one more compile definition
 *     -Dsrslte_ulsch_decode=srslte_ulsch_decode_cpu
 * and linked with this library.  The transport block (decode_tb, sch.c:1058-1062: rate de-matching, HARQ combining,
 * turbo decoding, CRCs) runs on the device through the soft buffer's device pool.  The values of the multiplexed
 * control information (HARQ-ACK, RI, CQI) are control-plane work and stay in the reference: this entry calls the
 * reference's own srslte_uci_decode_ack_ri, srslte_uci_decode_cqi_pusch, srslte_cqi_size, srslte_cqi_value_unpack
 * and srslte_uci_cfg_total_ack (uci.c, cqi.c: weak references, resolved by the link with libsrslte_phy; a PUSCH with
 * control information fails loudly when they are absent), in the reference's order and with the reference's side
 * effects on q_bits, g_bits, cfg->K_segm and cfg->uci_cfg.cqi.rank_is_not_one.  g_bits receives the de-interleaved
 * codeword exactly as ulsch_deinterleave leaves it (sch.c:891-918, including g[0], which ends up holding the RI
 * sample with the highest channel position).
 * The structures restate phch/cqi.h:121-142, phch/uci_cfg.h:27-70, phch/pusch_cfg.h:29-88, fec/turbocoder.h:46-49,
 * fec/crc.h:38-46 and phch/sch.h:52-74 (SRSLTE_MAX_CARRIERS = 5, SRSLTE_MAX_CODEWORDS = 2);
 * tests/test_compat_abi.py checks sizes and offsets against the compiled reference.                                */
typedef struct {
  bool     data_enable;
  bool     ri_present;
  bool     pmi_present;
  bool     four_antenna_ports;
  bool     rank_is_not_one;
  bool     subband_label_2_bits;
  uint32_t L;
  uint32_t N;
  uint32_t type; /* srslte_cqi_type_t: WIDEBAND = 0, SUBBAND, SUBBAND_UE, SUBBAND_HL */
  uint32_t ri_len;
} srslte_cqi_cfg_t;

typedef struct { uint8_t wideband_cqi, spatial_diff_cqi, pmi; } srslte_cqi_format2_wideband_t;
typedef struct { uint8_t subband_cqi, subband_label; } srslte_cqi_format2_subband_t;
typedef struct { uint8_t wideband_cqi, subband_diff_cqi; uint32_t position_subband; } srslte_cqi_ue_subband_t;
typedef struct {
  uint8_t  wideband_cqi_cw0;
  uint32_t subband_diff_cqi_cw0;
  uint8_t  wideband_cqi_cw1;
  uint32_t subband_diff_cqi_cw1;
  uint32_t pmi;
} srslte_cqi_hl_subband_t;

typedef struct {
  union {
    srslte_cqi_format2_wideband_t wideband;
    srslte_cqi_format2_subband_t  subband;
    srslte_cqi_ue_subband_t       subband_ue;
    srslte_cqi_hl_subband_t       subband_hl;
  } u;
  bool data_crc;
} srslte_cqi_value_t;

typedef struct {
  bool     pending_tb[2];
  uint32_t nof_acks;
  uint32_t ncce[9];
  uint32_t N_bundle;
  uint32_t tdd_ack_M;
  uint32_t tdd_ack_m;
  bool     tdd_is_multiplex;
  uint32_t tpc_for_pucch;
  uint32_t grant_cc_idx;
} srslte_uci_cfg_ack_t;

typedef struct {
  srslte_uci_cfg_ack_t ack[5];
  srslte_cqi_cfg_t     cqi;
  bool                 is_scheduling_request_tti;
} srslte_uci_cfg_t;

typedef struct {
  uint8_t ack_value[10];
  bool    valid;
} srslte_uci_value_ack_t;

typedef struct {
  bool                   scheduling_request;
  srslte_cqi_value_t     cqi;
  srslte_uci_value_ack_t ack;
  uint8_t                ri;
} srslte_uci_value_t;

typedef struct {
  uint32_t position;
  uint32_t type; /* srslte_uci_bit_type_t */
} srslte_uci_bit_t;

typedef struct {
  uint32_t I_offset_cqi;
  uint32_t I_offset_ri;
  uint32_t I_offset_ack;
} srslte_uci_offset_cfg_t;

typedef struct {
  bool           is_from_rar;
  uint32_t       L_prb;
  uint32_t       n_prb[2];
  uint32_t       n_prb_tilde[2];
  uint32_t       freq_hopping;
  uint32_t       nof_re;
  uint32_t       nof_symb;
  srslte_ra_tb_t tb;
  srslte_ra_tb_t last_tb;
  uint32_t       n_dmrs;
} srslte_pusch_grant_t;

typedef struct {
  uint16_t                rnti;
  srslte_uci_cfg_t        uci_cfg;
  srslte_uci_offset_cfg_t uci_offset;
  srslte_pusch_grant_t    grant;
  uint32_t                max_nof_iterations;
  uint32_t                last_O_cqi;
  uint32_t                K_segm;
  uint32_t                current_tx_nb;
  bool                    csi_enable;
  bool                    enable_64qam;
  union {
    void*                   tx;
    srslte_softbuffer_rx_t* rx;
  } softbuffers;
  bool     meas_time_en;
  uint32_t meas_time_value;
} srslte_pusch_cfg_t;

typedef struct {
  uint32_t max_long_cb;
  uint8_t* temp;
} srslte_tcod_t;

typedef struct {
  uint64_t table[256];
  int      polynom;
  int      order;
  uint64_t crcinit;
  uint64_t crcmask;
  uint64_t crchighbit;
  uint32_t srslte_crc_out;
} srslte_crc_t;

/* srslte_sch_t (sch.h:52-74) up to and including the two CRC objects; srslte_uci_cqi_pusch_t uci_cqi follows (its
 * address is handed to the reference's CQI decoder, its content is the reference's business) */
typedef struct {
  uint32_t         max_iterations;
  float            avg_iterations;
  bool             llr_is_8bit;
  uint8_t*         cb_in;
  uint8_t*         parity_bits;
  void*            e;
  uint8_t*         temp_g_bits;
  uint32_t*        ul_interleaver;
  srslte_uci_bit_t ack_ri_bits[57600];
  srslte_tcod_t    encoder;
  srslte_tdec_t    decoder;
  srslte_crc_t     crc_tb;
  srslte_crc_t     crc_cb;
  uint64_t         uci_cqi[1]; /* first word of srslte_uci_cqi_pusch_t */
} srslte_sch_ul_t;

SRSLTE_API int srslte_ulsch_decode(void* q /* srslte_sch_t* */, srslte_pusch_cfg_t* cfg, int16_t* q_bits, int16_t* g_bits,
                                   uint8_t* c_seq, uint8_t* data, srslte_uci_value_t* uci_data);

#ifdef __cplusplus
}
#endif
#endif /* SRSLTE_B200_COMPAT_H */
