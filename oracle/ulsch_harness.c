/*
 * oracle/ulsch_harness.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The level-2 check of SURVEY.md 8c for the uplink: transport blocks with multiplexed control information go through the
 * reference's UNCHANGED srslte_sch API the way pusch.c:367-393, 482-503 drives it
 *   srslte_sch_init -> srslte_ulsch_encode -> scrambling + placeholder bits -> LLRs + noise -> descrambling
 *   -> srslte_ulsch_decode (HARQ retransmissions rv 0, 2, 3, 1 on one soft buffer)
 * and every result (return code, iterations, TB bytes, ACK / RI / CQI values, the de-interleaved LLRs) is printed.
 * oracle/Makefile links this file twice against the reference's own objects: once as they are (ulsch_harness_ref) and
 * once with sch.c's decode entry points renamed out of the way by compile definitions, the turbo decoder sources left
 * out, and libsrslte_b200.so providing srslte_ulsch_decode, srslte_tdec_*, srslte_softbuffer_rx_* instead
 * (ulsch_harness_b200).  The two outputs must be identical.
 *
 * Reference API used: phch/sch.h:109-122, phch/pusch_cfg.h, phch/uci_cfg.h, fec/softbuffer.h, common/sequence.h,
 * scrambling/scrambling.h.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "srslte/srslte.h"

static uint32_t lcg(uint32_t* s)
{
  *s = *s * 1664525u + 1013904223u;
  return *s >> 8;
}

static float gauss(uint32_t* s)
{
  float a = 0;
  for (int i = 0; i < 12; i++) a += (float)(lcg(s) & 0xFFFF) / 65536.0f;
  return a - 6.0f;
}

static unsigned long fnv(const void* p, size_t n)
{
  unsigned long        h = 1469598103934665603ul;
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ul;
  return h;
}

typedef struct {
  int      tbs_idx;
  uint32_t L_prb, nof_symb, nof_ack, ri_len, cqi_mode; /* cqi_mode: 0 none, 1 wideband (4 bits), 2 sub-band (22 bits) */
} case_t;

int main(int argc, char** argv)
{
  const int      n_tb   = argc > 1 ? atoi(argv[1]) : 14;
  const float    sigma  = argc > 2 ? (float)atof(argv[2]) : 0.45f;
  const uint32_t max_it = argc > 3 ? (uint32_t)atoi(argv[3]) : 10;
  uint32_t       seed   = 4321;
  srslte_sch_t*  q      = (srslte_sch_t*)calloc(1, sizeof(srslte_sch_t));
  if (srslte_sch_init(q)) {
    fprintf(stderr, "srslte_sch_init failed\n");
    return 1;
  }
  srslte_sch_set_max_noi(q, max_it);
  srslte_softbuffer_tx_t stx;
  srslte_softbuffer_rx_t srx;
  srslte_softbuffer_tx_init(&stx, 100);
  srslte_softbuffer_rx_init(&srx, 100);
  static const case_t cases[] = {
      {20, 50, 12, 0, 0, 0}, {26, 100, 12, 0, 0, 0}, {12, 25, 12, 1, 0, 0}, {5, 6, 12, 2, 1, 0},  {16, 50, 11, 1, 1, 1},
      {24, 96, 12, 2, 1, 2}, {9, 15, 10, 0, 1, 0},   {3, 4, 12, 1, 2, 1},   {22, 72, 12, 2, 0, 2}, {14, 25, 11, 0, 2, 0},
      {26, 100, 12, 2, 1, 2}, {0, 1, 12, 1, 1, 0},   {10, 48, 12, 0, 0, 1}, {19, 36, 12, 2, 2, 2}};
  static const int rvs[4] = {0, 2, 3, 1};
  unsigned long    digest = 1469598103934665603ul;
  for (int t = 0; t < n_tb; t++) {
    const case_t*  c       = &cases[t % 14];
    const int      tbs     = srslte_ra_tbs_from_idx((uint32_t)c->tbs_idx, c->L_prb);
    srslte_mod_t   mod     = c->tbs_idx >= 21 ? SRSLTE_MOD_64QAM : c->tbs_idx >= 11 ? SRSLTE_MOD_16QAM : SRSLTE_MOD_QPSK;
    const uint32_t Qm      = srslte_mod_bits_x_symbol(mod);
    const uint32_t nof_re  = c->L_prb * 12 * c->nof_symb;
    const uint32_t nof_bits = nof_re * Qm;
    srslte_pusch_cfg_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.rnti              = 0x46;
    cfg.grant.L_prb       = c->L_prb;
    cfg.grant.nof_symb    = c->nof_symb;
    cfg.grant.nof_re      = nof_re;
    cfg.grant.tb.tbs      = tbs;
    cfg.grant.tb.mod      = mod;
    cfg.grant.tb.nof_bits = nof_bits;
    cfg.grant.tb.enabled  = true;
    cfg.uci_cfg.ack[0].nof_acks = c->nof_ack;
    cfg.uci_cfg.cqi.ri_len      = c->ri_len;
    if (c->cqi_mode) {
      cfg.uci_cfg.cqi.data_enable = true;
      cfg.uci_cfg.cqi.type        = c->cqi_mode == 1 ? SRSLTE_CQI_TYPE_WIDEBAND : SRSLTE_CQI_TYPE_SUBBAND_HL;
      cfg.uci_cfg.cqi.N           = 9;
    }
    cfg.uci_offset.I_offset_ack = 9;
    cfg.uci_offset.I_offset_ri  = 6;
    cfg.uci_offset.I_offset_cqi = 6;
    srslte_uci_value_t tx_uci;
    memset(&tx_uci, 0, sizeof(tx_uci));
    tx_uci.ack.ack_value[0] = (uint8_t)(lcg(&seed) & 1);
    tx_uci.ack.ack_value[1] = (uint8_t)(lcg(&seed) & 1);
    tx_uci.ri               = (uint8_t)(lcg(&seed) & (c->ri_len > 1 ? 3 : 1));
    if (c->cqi_mode == 1) {
      tx_uci.cqi.wideband.wideband_cqi = (uint8_t)(lcg(&seed) & 15);
    } else if (c->cqi_mode == 2) {
      tx_uci.cqi.subband_hl.wideband_cqi_cw0     = (uint8_t)(lcg(&seed) & 15);
      tx_uci.cqi.subband_hl.subband_diff_cqi_cw0 = lcg(&seed) & 0x3ffff;
    }
    uint8_t* data    = (uint8_t*)calloc(1, tbs / 8 + 16);
    uint8_t* data_rx = (uint8_t*)calloc(1, tbs / 8 + 16);
    uint8_t* g_tx    = (uint8_t*)calloc(1, nof_bits + 64);
    uint8_t* q_tx    = (uint8_t*)calloc(1, nof_bits + 64);
    int16_t* llr     = (int16_t*)srslte_vec_malloc(sizeof(int16_t) * (nof_bits + 64));
    int16_t* g_rx    = (int16_t*)srslte_vec_malloc(sizeof(int16_t) * (nof_bits + 64));
    for (int i = 0; i < tbs / 8; i++) data[i] = (uint8_t)lcg(&seed);
    srslte_sequence_t seq;
    memset(&seq, 0, sizeof(seq));
    /* c_init of srslte_sequence_pusch (phch/sequences.c:65-68): (rnti << 14) + ((nslot / 2) << 9) + cell_id */
    if (srslte_sequence_LTE_pr(&seq, nof_bits, ((uint32_t)cfg.rnti << 14) + ((uint32_t)(t % 10) << 9) + 1u + (uint32_t)t)) {
      fprintf(stderr, "sequence failed\n");
      return 1;
    }
    srslte_softbuffer_tx_reset_tbs(&stx, (uint32_t)tbs);
    srslte_softbuffer_rx_reset_tbs(&srx, (uint32_t)tbs);
    int ret = -1, tx;
    for (tx = 0; tx < 4 && ret != 0; tx++) {
      cfg.grant.tb.rv    = rvs[tx];
      cfg.softbuffers.tx = &stx;
      memset(q_tx, 0, nof_bits / 8 + 8);
      const int n_ri_ack = srslte_ulsch_encode(q, &cfg, data, &tx_uci, g_tx, q_tx);
      if (n_ri_ack < 0) {
        fprintf(stderr, "encode failed\n");
        return 1;
      }
      /* scrambling and the placeholder / repetition bits of ACK and RI, as pusch.c:382-398 does it */
      srslte_scrambling_bytes(&seq, q_tx, (int)nof_bits);
      for (int i = 0; i < n_ri_ack; i++) {
        const uint32_t p = q->ack_ri_bits[i].position;
        if (q->ack_ri_bits[i].type == UCI_BIT_PLACEHOLDER) {
          q_tx[p / 8] |= (uint8_t)(1 << (7 - p % 8));
        } else if (q->ack_ri_bits[i].type == UCI_BIT_REPETITION && p > 1) {
          if (q_tx[(p - 1) / 8] & (1 << (7 - (p - 1) % 8)))
            q_tx[p / 8] |= (uint8_t)(1 << (7 - p % 8));
          else
            q_tx[p / 8] &= (uint8_t) ~(1 << (7 - p % 8));
        }
      }
      const float s = sigma * (1.0f + 0.3f * (float)(t % 4));
      for (uint32_t i = 0; i < nof_bits; i++) {
        const int bit = (q_tx[i >> 3] >> (7 - (i & 7))) & 1;
        float     v   = (bit ? 1.0f : -1.0f) + s * gauss(&seed);
        v *= 40.0f;
        llr[i] = (int16_t)(v > 32000.f ? 32000.f : v < -32000.f ? -32000.f : v);
      }
      srslte_scrambling_s_offset(&seq, llr, 0, (int)nof_bits);
      cfg.softbuffers.rx = &srx;
      memset(data_rx, 0, tbs / 8 + 3);
      memset(g_rx, 0, sizeof(int16_t) * (nof_bits + 64));
      srslte_uci_value_t rx_uci;
      memset(&rx_uci, 0, sizeof(rx_uci));
      ret = srslte_ulsch_decode(q, &cfg, llr, g_rx, seq.c, data_rx, &rx_uci);
      const unsigned long hd = fnv(data_rx, (size_t)(tbs / 8 + 3)), hg = fnv(g_rx, sizeof(int16_t) * nof_bits),
                          hq = fnv(llr, sizeof(int16_t) * nof_bits);
      digest = (((digest ^ hd) * 1099511628211ul) ^ hg) * 1099511628211ul;
      printf("tb %2d tbs %6d prb %3u Qm %u rv %d ack %u ri %u cqi %u: ret %2d noi %.3f match %d K_segm %u rank>1 %d | ack %u%u/%u%u "
             "ri %u/%u cqi %u crc %d | bytes %016lx g %016lx q %016lx\n",
             t, tbs, c->L_prb, Qm, rvs[tx], c->nof_ack, c->ri_len, c->cqi_mode, ret, srslte_sch_last_noi(q),
             memcmp(data, data_rx, tbs / 8) == 0, cfg.K_segm, cfg.uci_cfg.cqi.rank_is_not_one, rx_uci.ack.ack_value[0],
             rx_uci.ack.ack_value[1], tx_uci.ack.ack_value[0], tx_uci.ack.ack_value[1], rx_uci.ri, tx_uci.ri,
             c->cqi_mode == 1 ? rx_uci.cqi.wideband.wideband_cqi : rx_uci.cqi.subband_hl.wideband_cqi_cw0,
             rx_uci.cqi.data_crc, hd, hg, hq);
    }
    srslte_sequence_free(&seq);
    free(data); free(data_rx); free(g_tx); free(q_tx); free(llr); free(g_rx);
  }
  printf("digest %016lx\n", digest);
  srslte_softbuffer_tx_free(&stx);
  srslte_softbuffer_rx_free(&srx);
  srslte_sch_free(q);
  free(q);
  return 0;
}
