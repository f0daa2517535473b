/*
 * oracle/dlsch_harness.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The level-2 check of SURVEY.md 8c: transport blocks go through the reference's UNCHANGED srslte_sch API
 *   srslte_sch_init -> srslte_dlsch_encode2 -> BPSK-like LLRs + AWGN -> srslte_dlsch_decode2 (with HARQ retransmissions)
 * and every result (return code, iterations, TB bytes) is printed.  oracle/Makefile links this file twice against the
 * reference's own objects: once as they are (dlsch_harness_ref) and once with sch.c's two DL decode entry points
 * renamed out of the way by a compile definition, the turbo decoder sources left out, and libsrslte_b200.so providing
 * srslte_dlsch_decode2, srslte_tdec_* instead (dlsch_harness_b200).  The two outputs must be identical.
 *
 * Reference API used: phch/sch.h:79-107, phch/pdsch_cfg.h, fec/softbuffer.h:52-76, fec/cbsegm.h, common/phy_common.h.
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

/* deterministic gaussian (sum of 12 uniforms), so that both binaries see the same noise without libm differences */
static float gauss(uint32_t* s)
{
  float a = 0;
  for (int i = 0; i < 12; i++) a += (float)(lcg(s) & 0xFFFF) / 65536.0f;
  return a - 6.0f;
}

int main(int argc, char** argv)
{
  const int      n_tb   = argc > 1 ? atoi(argv[1]) : 12;
  const float    sigma  = argc > 2 ? (float)atof(argv[2]) : 0.45f;
  const uint32_t max_it = argc > 3 ? (uint32_t)atoi(argv[3]) : 10;
  uint32_t       seed   = 12345;
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
  static const int      tbs_idx[] = {26, 20, 12, 5, 26, 16, 9, 24, 3, 26, 14, 22};
  static const uint32_t prb_tab[] = {100, 50, 25, 6, 75, 100, 15, 50, 100, 100, 25, 6};
  static const int      rvs[4]    = {0, 2, 3, 1};
  unsigned long         hash      = 1469598103934665603ul;
  for (int t = 0; t < n_tb; t++) {
    const uint32_t nof_prb = prb_tab[t % 12];
    const int      tbs     = srslte_ra_tbs_from_idx((uint32_t)tbs_idx[t % 12], nof_prb);
    srslte_mod_t   mod     = tbs_idx[t % 12] >= 16 ? SRSLTE_MOD_64QAM : tbs_idx[t % 12] >= 10 ? SRSLTE_MOD_16QAM : SRSLTE_MOD_QPSK;
    const uint32_t Qm      = srslte_mod_bits_x_symbol(mod);
    const uint32_t nof_re  = nof_prb * 12 * 11;
    const uint32_t nof_bits = nof_re * Qm;
    srslte_pdsch_cfg_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.grant.nof_tb      = 1;
    cfg.grant.nof_layers  = 1;
    cfg.grant.nof_prb     = nof_prb;
    cfg.grant.nof_re      = nof_re;
    cfg.grant.tb[0].tbs     = tbs;
    cfg.grant.tb[0].mod     = mod;
    cfg.grant.tb[0].nof_bits = nof_bits;
    cfg.grant.tb[0].enabled = true;
    uint8_t* data    = (uint8_t*)calloc(1, tbs / 8 + 16);
    uint8_t* data_rx = (uint8_t*)calloc(1, tbs / 8 + 16);
    uint8_t* e       = (uint8_t*)calloc(1, nof_bits + 64);
    int16_t* llr     = (int16_t*)srslte_vec_malloc(sizeof(int16_t) * (nof_bits + 64));
    for (int i = 0; i < tbs / 8; i++) data[i] = (uint8_t)lcg(&seed);
    srslte_softbuffer_tx_reset_tbs(&stx, (uint32_t)tbs);
    srslte_softbuffer_rx_reset_tbs(&srx, (uint32_t)tbs);
    int ret = -1, tx;
    for (tx = 0; tx < 4 && ret != 0; tx++) {
      cfg.grant.tb[0].rv   = rvs[tx];
      cfg.softbuffers.tx[0] = &stx;
      if (srslte_dlsch_encode2(q, &cfg, data, e, 0, 1)) {
        fprintf(stderr, "encode failed\n");
        return 1;
      }
      /* the TB-specific noise level makes some blocks need a retransmission */
      const float s = sigma * (1.0f + 0.25f * (float)(t % 5));
      for (uint32_t i = 0; i < nof_bits; i++) {
        const int bit = (e[i >> 3] >> (7 - (i & 7))) & 1; /* the rate matcher packs its output (sch.c: srslte_rm_turbo_tx_lut) */
        float v = (bit ? 1.0f : -1.0f) + s * gauss(&seed); /* bit 1 -> positive LLR (turbodecoder_test.c:218) */
        v *= 40.0f;
        llr[i] = (int16_t)(v > 32000.f ? 32000.f : v < -32000.f ? -32000.f : v);
      }
      cfg.softbuffers.rx[0] = &srx;
      memset(data_rx, 0, tbs / 8 + 3);
      ret = srslte_dlsch_decode2(q, &cfg, llr, data_rx, 0, 1);
      unsigned long h = 1469598103934665603ul;
      for (int i = 0; i < tbs / 8 + 3; i++) h = (h ^ data_rx[i]) * 1099511628211ul;
      hash = (hash ^ h) * 1099511628211ul;
      printf("tb %2d tbs %6d prb %3u Qm %u rv %d: ret %2d noi %.3f match %d bytes %016lx\n", t, tbs, nof_prb, Qm, rvs[tx], ret,
             srslte_sch_last_noi(q), memcmp(data, data_rx, tbs / 8) == 0, h);
    }
    free(data); free(data_rx); free(e); free(llr);
  }
  printf("digest %016lx\n", hash);
  srslte_softbuffer_tx_free(&stx);
  srslte_softbuffer_rx_free(&srx);
  srslte_sch_free(q);
  free(q);
  return 0;
}
