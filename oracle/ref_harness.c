/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin glue that is compiled TOGETHER WITH the reference's own, unmodified source files
 * (taken where they lie under /root/reference; see oracle/Makefile) into
 * oracle/_ref/libsrslte_ref.so.  It adds nothing to the arithmetic: every function here
 * only allocates reference objects and calls the reference's public API, so that the
 * tests and the CPU-baseline leg of bench.py can drive the real srslte_tdec_* /
 * srslte_dlsch_* code through ctypes without knowing struct layouts.
 *
 * Reference API used: turbodecoder.h:97-135, turbocoder.h:54-61, rm_turbo.h:45-93,
 * crc.h:48-83, sch.h:76-115, softbuffer.h:52-76 (all under lib/include/srslte/phy).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "srslte/srslte.h"

/* ---- sizes, so that python can cross-check the ABI mirror in include/ ---- */
size_t refh_sizeof_tdec(void) { return sizeof(srslte_tdec_t); }
size_t refh_sizeof_sch(void) { return sizeof(srslte_sch_t); }
size_t refh_sizeof_crc(void) { return sizeof(srslte_crc_t); }
size_t refh_sizeof_softbuffer_rx(void) { return sizeof(srslte_softbuffer_rx_t); }
size_t refh_sizeof_cbsegm(void) { return sizeof(srslte_cbsegm_t); }
size_t refh_offsetof_tdec_n_iter(void) { return offsetof(srslte_tdec_t, n_iter); }
size_t refh_offsetof_sch_decoder(void) { return offsetof(srslte_sch_t, decoder); }
/* layouts the DL-SCH compat entry points read: [sizeof ra_tb, sizeof pdsch_grant, sizeof pdsch_cfg, offsets of
 * grant.tb, grant.nof_tb, tb.nof_bits, tb.rv, cfg.softbuffers, sch.max_iterations, sch.avg_iterations, sch.llr_is_8bit] */
void refh_dlsch_layout(size_t out[11])
{
  out[0]  = sizeof(srslte_ra_tb_t);
  out[1]  = sizeof(srslte_pdsch_grant_t);
  out[2]  = sizeof(srslte_pdsch_cfg_t);
  out[3]  = offsetof(srslte_pdsch_cfg_t, grant) + offsetof(srslte_pdsch_grant_t, tb);
  out[4]  = offsetof(srslte_pdsch_cfg_t, grant) + offsetof(srslte_pdsch_grant_t, nof_tb);
  out[5]  = offsetof(srslte_ra_tb_t, nof_bits);
  out[6]  = offsetof(srslte_ra_tb_t, rv);
  out[7]  = offsetof(srslte_pdsch_cfg_t, softbuffers);
  out[8]  = offsetof(srslte_sch_t, max_iterations);
  out[9]  = offsetof(srslte_sch_t, avg_iterations);
  out[10] = offsetof(srslte_sch_t, llr_is_8bit);
}

/* layouts srslte_ulsch_decode reads (include/srslte_b200_compat.h restates them; tests/test_compat_abi.py) */
#include "srslte/phy/phch/pusch_cfg.h"
void refh_ulsch_layout(size_t out[24])
{
  out[0]  = sizeof(srslte_sch_t);
  out[1]  = offsetof(srslte_sch_t, ack_ri_bits);
  out[2]  = offsetof(srslte_sch_t, encoder);
  out[3]  = offsetof(srslte_sch_t, decoder);
  out[4]  = offsetof(srslte_sch_t, crc_tb);
  out[5]  = offsetof(srslte_sch_t, uci_cqi);
  out[6]  = sizeof(srslte_pusch_cfg_t);
  out[7]  = offsetof(srslte_pusch_cfg_t, uci_cfg);
  out[8]  = offsetof(srslte_pusch_cfg_t, uci_cfg) + offsetof(srslte_uci_cfg_t, cqi);
  out[9]  = offsetof(srslte_pusch_cfg_t, uci_offset);
  out[10] = offsetof(srslte_pusch_cfg_t, grant);
  out[11] = offsetof(srslte_pusch_cfg_t, grant) + offsetof(srslte_pusch_grant_t, nof_symb);
  out[12] = offsetof(srslte_pusch_cfg_t, grant) + offsetof(srslte_pusch_grant_t, tb);
  out[13] = offsetof(srslte_pusch_cfg_t, K_segm);
  out[14] = offsetof(srslte_pusch_cfg_t, softbuffers);
  out[15] = sizeof(srslte_uci_cfg_ack_t);
  out[16] = offsetof(srslte_uci_cfg_ack_t, nof_acks);
  out[17] = sizeof(srslte_cqi_cfg_t);
  out[18] = offsetof(srslte_cqi_cfg_t, ri_len);
  out[19] = sizeof(srslte_uci_value_t);
  out[20] = offsetof(srslte_uci_value_t, cqi) + offsetof(srslte_cqi_value_t, data_crc);
  out[21] = offsetof(srslte_uci_value_t, ack);
  out[22] = offsetof(srslte_uci_value_t, ri);
  out[23] = sizeof(srslte_uci_bit_t);
}

/* ---- decoder handle ---- */
srslte_tdec_t* refh_tdec_new(uint32_t max_k, int impl /* 0 = AUTO */)
{
  srslte_tdec_t* h = (srslte_tdec_t*)srslte_vec_malloc(sizeof(srslte_tdec_t));
  if (!h) return NULL;
  if (srslte_tdec_init_manual(h, max_k, (srslte_tdec_impl_type_t)impl)) {
    free(h);
    return NULL;
  }
  return h;
}

void refh_tdec_free(srslte_tdec_t* h)
{
  if (h) {
    srslte_tdec_free(h);
    free(h);
  }
}

/* soft values the decision was taken from (SB storage order for window decoders) */
const int16_t* refh_tdec_llr_ptr(srslte_tdec_t* h)
{
  return (h->n_iter % 2) == 0 ? (const int16_t*)h->app1 : (const int16_t*)h->ext1;
}

/* ---- batch of srslte_tdec_run_all, `threads` pthreads, one handle per thread ---- */
typedef struct {
  const int16_t* in;
  uint8_t*       out;
  uint32_t       in_stride, out_stride, first, last, K, nit;
  int            natural, rc;
} job_t;

static void* worker(void* arg)
{
  job_t*         j = (job_t*)arg;
  srslte_tdec_t* h = refh_tdec_new(6144, 0);
  if (!h) {
    j->rc = -1;
    return NULL;
  }
  if (j->natural) srslte_tdec_force_not_sb(h);
  /* run_all re-reads `input` in SB mode and writes the tail copies into its pads: give it
   * a private aligned copy, as the reference requires SIMD-aligned input anyway.        */
  uint32_t len = j->natural ? 3 * j->K + 12 : 3 * (j->K + 32) + 12;
  int16_t* buf = (int16_t*)srslte_vec_malloc(sizeof(int16_t) * (len + 64));
  for (uint32_t i = j->first; i < j->last; i++) {
    memcpy(buf, j->in + (size_t)i * j->in_stride, sizeof(int16_t) * len);
    if (srslte_tdec_run_all(h, buf, j->out + (size_t)i * j->out_stride, j->nit, j->K)) j->rc = -1;
  }
  free(buf);
  refh_tdec_free(h);
  return NULL;
}

int refh_batch_run_all(const int16_t* in, uint32_t in_stride, int natural, uint8_t* out,
                       uint32_t out_stride, uint32_t n, uint32_t K, uint32_t nof_iterations,
                       uint32_t threads)
{
  if (threads == 0) threads = 1;
  if (threads > n) threads = n ? n : 1;
  pthread_t* th   = (pthread_t*)calloc(threads, sizeof(pthread_t));
  job_t*     jobs = (job_t*)calloc(threads, sizeof(job_t));
  int        rc   = 0;
  for (uint32_t t = 0; t < threads; t++) {
    jobs[t] = (job_t){in, out, in_stride, out_stride, (uint32_t)((uint64_t)n * t / threads),
                      (uint32_t)((uint64_t)n * (t + 1) / threads), K, nof_iterations, natural, 0};
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  for (uint32_t t = 0; t < threads; t++) {
    pthread_join(th[t], NULL);
    if (jobs[t].rc) rc = -1;
  }
  free(th);
  free(jobs);
  return rc;
}

/* ---- persistent pool: one srslte_tdec_t per worker thread, created ONCE (like turbodecoder_test.c:190-266, which
 * initialises the decoder and then loops over srslte_tdec_run_all); refh_pool_run times only the decode loops ---- */
#include <time.h>
typedef struct refh_pool {
  uint32_t        threads;
  pthread_t*      th;
  pthread_mutex_t mu;
  pthread_cond_t  cv_go, cv_done;
  uint64_t        epoch;     /* bumped per run */
  uint32_t        pending;   /* workers still busy in this run */
  uint32_t        ready;     /* workers that have created their decoder handles */
  int             quit;
  /* job */
  const int16_t*  in;
  uint8_t*        out;
  uint32_t        in_stride, out_stride, n, K, nit;
  int             natural, rc;
  double*         busy_s;    /* per worker: seconds spent inside the srslte_tdec_run_all loop of the last run */
} refh_pool_t;

typedef struct {
  refh_pool_t* p;
  uint32_t     idx;
} pool_arg_t;

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void* pool_worker(void* arg)
{
  pool_arg_t*    pa = (pool_arg_t*)arg;
  refh_pool_t*   p  = pa->p;
  const uint32_t t  = pa->idx;
  free(pa);
  srslte_tdec_t* h     = refh_tdec_new(6144, 0);
  srslte_tdec_t* h_nat = refh_tdec_new(6144, 0);
  if (h_nat) srslte_tdec_force_not_sb(h_nat);
  int16_t* buf = (int16_t*)srslte_vec_malloc(sizeof(int16_t) * (3 * (6144 + 32) + 12 + 64));
  uint64_t seen = 0;
  pthread_mutex_lock(&p->mu);
  if (++p->ready == p->threads) pthread_cond_signal(&p->cv_done);
  pthread_mutex_unlock(&p->mu);
  for (;;) {
    pthread_mutex_lock(&p->mu);
    while (!p->quit && p->epoch == seen) pthread_cond_wait(&p->cv_go, &p->mu);
    if (p->quit) {
      pthread_mutex_unlock(&p->mu);
      break;
    }
    seen = p->epoch;
    pthread_mutex_unlock(&p->mu);
    int rc = (h && h_nat && buf) ? 0 : -1;
    const uint32_t first = (uint32_t)((uint64_t)p->n * t / p->threads), last = (uint32_t)((uint64_t)p->n * (t + 1) / p->threads);
    const uint32_t len = p->natural ? 3 * p->K + 12 : 3 * (p->K + 32) + 12;
    srslte_tdec_t* hh = p->natural ? h_nat : h;
    const double   t0 = now_s();
    for (uint32_t i = first; i < last && !rc; i++) {
      memcpy(buf, p->in + (size_t)i * p->in_stride, sizeof(int16_t) * len);
      if (srslte_tdec_run_all(hh, buf, p->out + (size_t)i * p->out_stride, p->nit, p->K)) rc = -1;
    }
    p->busy_s[t] = now_s() - t0;
    pthread_mutex_lock(&p->mu);
    if (rc) p->rc = -1;
    if (--p->pending == 0) pthread_cond_signal(&p->cv_done);
    pthread_mutex_unlock(&p->mu);
  }
  free(buf);
  refh_tdec_free(h);
  refh_tdec_free(h_nat);
  return NULL;
}

refh_pool_t* refh_pool_create(uint32_t threads)
{
  if (threads == 0) threads = 1;
  refh_pool_t* p = (refh_pool_t*)calloc(1, sizeof(refh_pool_t));
  if (!p) return NULL;
  p->threads = threads;
  p->th      = (pthread_t*)calloc(threads, sizeof(pthread_t));
  p->busy_s  = (double*)calloc(threads, sizeof(double));
  pthread_mutex_init(&p->mu, NULL);
  pthread_cond_init(&p->cv_go, NULL);
  pthread_cond_init(&p->cv_done, NULL);
  for (uint32_t t = 0; t < threads; t++) {
    pool_arg_t* pa = (pool_arg_t*)malloc(sizeof(pool_arg_t));
    pa->p   = p;
    pa->idx = t;
    pthread_create(&p->th[t], NULL, pool_worker, pa);
  }
  pthread_mutex_lock(&p->mu);  /* handle creation (QPP tables for all 188 sizes, beta arrays) ends here */
  while (p->ready < threads) pthread_cond_wait(&p->cv_done, &p->mu);
  pthread_mutex_unlock(&p->mu);
  return p;
}

/* decodes n blocks; returns 0 and the wall seconds between the release of the workers and the end of the last one
 * (handle creation, QPP table generation and thread start are NOT in it); busy_sum = sum of the workers' loop times */
int refh_pool_run(refh_pool_t* p, const int16_t* in, uint32_t in_stride, int natural, uint8_t* out, uint32_t out_stride,
                  uint32_t n, uint32_t K, uint32_t nof_iterations, double* wall_s, double* busy_sum_s)
{
  if (!p) return -1;
  pthread_mutex_lock(&p->mu);
  p->in = in; p->out = out; p->in_stride = in_stride; p->out_stride = out_stride;
  p->n = n; p->K = K; p->nit = nof_iterations; p->natural = natural; p->rc = 0;
  p->pending = p->threads;
  p->epoch++;
  const double t0 = now_s();
  pthread_cond_broadcast(&p->cv_go);
  while (p->pending) pthread_cond_wait(&p->cv_done, &p->mu);
  const double t1 = now_s();
  const int    rc = p->rc;
  double       bs = 0;
  for (uint32_t t = 0; t < p->threads; t++) bs += p->busy_s[t];
  pthread_mutex_unlock(&p->mu);
  if (wall_s) *wall_s = t1 - t0;
  if (busy_sum_s) *busy_sum_s = bs;
  return rc;
}

void refh_pool_destroy(refh_pool_t* p)
{
  if (!p) return;
  pthread_mutex_lock(&p->mu);
  p->quit = 1;
  pthread_cond_broadcast(&p->cv_go);
  pthread_mutex_unlock(&p->mu);
  for (uint32_t t = 0; t < p->threads; t++) pthread_join(p->th[t], NULL);
  pthread_mutex_destroy(&p->mu);
  pthread_cond_destroy(&p->cv_go);
  pthread_cond_destroy(&p->cv_done);
  free(p->th);
  free(p->busy_s);
  free(p);
}

/* ---- per-iteration trace of one block, the way sch.c drives the decoder ---- */
int refh_tdec_trace(const int16_t* in, int natural, uint32_t K, uint32_t nof_iterations,
                    uint8_t* out_bytes /* nit * K/8 */, int16_t* out_llr /* nit * K, nullable */)
{
  srslte_tdec_t* h = refh_tdec_new(6144, 0);
  if (!h) return -1;
  if (natural) srslte_tdec_force_not_sb(h);
  uint32_t len = natural ? 3 * K + 12 : 3 * (K + 32) + 12;
  int16_t* buf = (int16_t*)srslte_vec_malloc(sizeof(int16_t) * (len + 64));
  memcpy(buf, in, sizeof(int16_t) * len);
  int rc = srslte_tdec_new_cb(h, K);
  for (uint32_t it = 0; it < nof_iterations && !rc; it++) {
    srslte_tdec_iteration(h, buf, out_bytes + (size_t)it * (K / 8));
    if (out_llr) memcpy(out_llr + (size_t)it * K, refh_tdec_llr_ptr(h), sizeof(int16_t) * K);
  }
  free(buf);
  refh_tdec_free(h);
  return rc;
}

/* ---- encoder side (vector generation only) ---- */
int refh_tcod_encode(const uint8_t* bits, uint8_t* coded /* 3K+12 */, uint32_t K)
{
  srslte_tcod_t tc;
  if (srslte_tcod_init(&tc, 6144)) return -1;
  int rc = srslte_tcod_encode(&tc, (uint8_t*)bits, coded, K);
  srslte_tcod_free(&tc);
  return rc;
}

int refh_rm_turbo_tx(const uint8_t* coded, uint32_t K, uint8_t* e, uint32_t E, uint32_t rv)
{
  uint8_t* w  = (uint8_t*)calloc(3 * 6176 + 64, 1);
  /* the circular buffer w is only (re)built by an rv 0 call: rm_turbo.c:950 */
  int rc = srslte_rm_turbo_tx(w, 3 * 6176, (uint8_t*)coded, 3 * K + 12, e, E, 0);
  if (!rc && rv) rc = srslte_rm_turbo_tx(w, 3 * 6176, (uint8_t*)coded, 3 * K + 12, e, E, rv);
  free(w);
  return rc;
}

/* ---- transport-block level: srslte_dlsch_encode2 / srslte_dlsch_decode2 ---- */
typedef struct {
  srslte_sch_t           sch;
  srslte_softbuffer_tx_t sb_tx;
  srslte_softbuffer_rx_t sb_rx;
  srslte_pdsch_cfg_t     cfg;
} refh_tb_t;

refh_tb_t* refh_tb_new(void)
{
  refh_tb_t* t = (refh_tb_t*)srslte_vec_malloc(sizeof(refh_tb_t));
  if (!t) return NULL;
  memset(t, 0, sizeof(*t));
  if (srslte_sch_init(&t->sch) || srslte_softbuffer_tx_init(&t->sb_tx, 110) ||
      srslte_softbuffer_rx_init(&t->sb_rx, 110)) {
    free(t);
    return NULL;
  }
  return t;
}

void refh_tb_free(refh_tb_t* t)
{
  if (!t) return;
  srslte_sch_free(&t->sch);
  srslte_softbuffer_tx_free(&t->sb_tx);
  srslte_softbuffer_rx_free(&t->sb_rx);
  free(t);
}

static void fill_cfg(refh_tb_t* t, uint32_t tbs, uint32_t qm, uint32_t rv, uint32_t G, int tx)
{
  memset(&t->cfg, 0, sizeof(t->cfg));
  t->cfg.grant.nof_tb     = 1;
  t->cfg.grant.nof_layers = 1;
  t->cfg.grant.tb[0].mod  = qm == 2 ? SRSLTE_MOD_QPSK : qm == 4 ? SRSLTE_MOD_16QAM
                          : qm == 6 ? SRSLTE_MOD_64QAM : SRSLTE_MOD_256QAM;
  t->cfg.grant.tb[0].tbs      = (int)tbs;
  t->cfg.grant.tb[0].rv       = (int)rv;
  t->cfg.grant.tb[0].nof_bits = G;
  t->cfg.grant.tb[0].enabled  = true;
  if (tx)
    t->cfg.softbuffers.tx[0] = &t->sb_tx;
  else
    t->cfg.softbuffers.rx[0] = &t->sb_rx;
}

/* data: tbs/8 bytes in; e_bits: G unpacked bits (one per byte) out */
int refh_tb_encode(refh_tb_t* t, uint32_t tbs, uint32_t qm, uint32_t rv, uint32_t G,
                   const uint8_t* data, uint8_t* e_bits)
{
  uint8_t* packed = (uint8_t*)calloc(G / 8 + 64, 1);
  uint8_t* d      = (uint8_t*)calloc(tbs / 8 + 64, 1);
  memcpy(d, data, tbs / 8);
  fill_cfg(t, tbs, qm, rv, G, 1);
  if (rv == 0) srslte_softbuffer_tx_reset_tbs(&t->sb_tx, tbs);
  int rc = srslte_dlsch_encode2(&t->sch, &t->cfg, d, packed, 0, 1);
  if (!rc) srslte_bit_unpack_vector(packed, e_bits, (int)G);
  free(packed);
  free(d);
  return rc;
}

void refh_tb_rx_reset(refh_tb_t* t, uint32_t tbs) { srslte_softbuffer_rx_reset_tbs(&t->sb_rx, tbs); }

/* out: tbs/8 + 8 bytes.  Returns srslte_dlsch_decode2's value; *avg_it = srslte_sch_last_noi */
int refh_tb_decode(refh_tb_t* t, uint32_t tbs, uint32_t qm, uint32_t rv, uint32_t G,
                   const int16_t* llr, uint8_t* out, uint32_t max_it, float* avg_it,
                   uint8_t* cb_crc /* nullable, C entries */)
{
  int16_t* e = (int16_t*)srslte_vec_malloc(sizeof(int16_t) * (G + 64));
  memcpy(e, llr, sizeof(int16_t) * G);
  fill_cfg(t, tbs, qm, rv, G, 0);
  srslte_sch_set_max_noi(&t->sch, max_it);
  int rc = srslte_dlsch_decode2(&t->sch, &t->cfg, e, out, 0, 1);
  if (avg_it) *avg_it = srslte_sch_last_noi(&t->sch);
  if (cb_crc) {
    srslte_cbsegm_t s;
    srslte_cbsegm(&s, tbs);
    for (uint32_t i = 0; i < s.C; i++) cb_crc[i] = t->sb_rx.cb_crc[i];
  }
  free(e);
  return rc;
}

const int16_t* refh_tb_softbuffer(refh_tb_t* t, uint32_t cb) { return t->sb_rx.buffer_f[cb]; }

/* ---- CRC through the reference's table implementation ---- */
uint32_t refh_crc_bytes(uint32_t poly, const uint8_t* data, uint32_t nbits)
{
  srslte_crc_t c;
  srslte_crc_init(&c, poly, 24);
  return srslte_crc_checksum_byte(&c, (uint8_t*)data, (int)nbits);
}

uint32_t refh_crc_bits(uint32_t poly, const uint8_t* bits, uint32_t nbits)
{
  srslte_crc_t c;
  srslte_crc_init(&c, poly, 24);
  return srslte_crc_checksum(&c, (uint8_t*)bits, (int)nbits);
}

/* ---- front end: the reference's own soft demodulator + descrambler, pdsch.c:760-779 order ---- */
#include "srslte/phy/common/sequence.h"
#include "srslte/phy/modem/demod_soft.h"
#include "srslte/phy/scrambling/scrambling.h"
/* mod: srslte_mod_t (1 QPSK, 2 16QAM, 3 64QAM, 4 256QAM); buffers are copied into SIMD-aligned memory because the
 * reference's SSE demodulators use aligned loads/stores (demod_soft.c:104-119) */
int refh_demod_descramble(int mod, const float* sym, uint32_t nsym, uint32_t c_init, uint32_t nof_bits, int16_t* llr,
                          uint32_t nof_llr)
{
  cf_t*  s = srslte_vec_malloc(sizeof(cf_t) * (nsym + 16));
  short* l = srslte_vec_malloc(sizeof(short) * (nof_llr + 64));
  if (!s || !l) return -1;
  memcpy(s, sym, sizeof(cf_t) * nsym);
  int rc = srslte_demod_soft_demodulate_s((srslte_mod_t)mod, s, l, (int)nsym);
  if (rc == 0 && nof_bits) {
    srslte_sequence_t seq;
    memset(&seq, 0, sizeof(seq));
    if (srslte_sequence_LTE_pr(&seq, nof_bits, c_init)) rc = -1;
    if (rc == 0) srslte_scrambling_s_offset(&seq, l, 0, (int)nof_bits);
    srslte_sequence_free(&seq);
  }
  memcpy(llr, l, sizeof(short) * nof_llr);
  free(s);
  free(l);
  return rc;
}

/* the reference's UL-SCH channel de-interleaver (sch.c:891-918; not static) with no RI bits */
void ulsch_deinterleave(int16_t* q_bits, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs, int16_t* g_bits,
                        srslte_uci_bit_t* ri_bits, uint32_t nof_ri_bits, uint8_t* ri_present, uint32_t* inteleaver_lut);
int refh_ulsch_deinterleave(const int16_t* q, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs, int16_t* g)
{
  const uint32_t n = H_prime_total * Qm;
  int16_t*  qq  = malloc(sizeof(int16_t) * (n + 8));
  uint8_t*  pre = calloc(n + 8, 1);
  uint32_t* lut = calloc(n + 8, sizeof(uint32_t));
  if (!qq || !pre || !lut) return -1;
  memcpy(qq, q, sizeof(int16_t) * n);
  ulsch_deinterleave(qq, Qm, H_prime_total, N_pusch_symbs, g, NULL, 0, pre, lut);
  free(qq);
  free(pre);
  free(lut);
  return 0;
}

/* ---- UL-SCH with multiplexed UCI: the reference's own srslte_ulsch_encode / srslte_ulsch_decode (sch.c:1013-1232) ----
 * u[12] = { tbs, qm, rv, nb_q (= Qm * nof_re), L_prb, nof_symb, nof_ack, ri_len, cqi_mode (0 none, 1 wideband = 4 bits,
 *           2 higher-layer sub-band with N = 9 = 22 bits -> the "long" CQI branch), I_offset_ack, I_offset_ri, I_offset_cqi } */
#include "srslte/phy/phch/pusch_cfg.h"
static void fill_ul_cfg(refh_tb_t* t, srslte_pusch_cfg_t* c, const uint32_t* u, int tx)
{
  memset(c, 0, sizeof(*c));
  c->grant.tb.mod      = u[1] == 2 ? SRSLTE_MOD_QPSK : u[1] == 4 ? SRSLTE_MOD_16QAM : SRSLTE_MOD_64QAM;
  c->grant.tb.tbs      = (int)u[0];
  c->grant.tb.rv       = (int)u[2];
  c->grant.tb.nof_bits = u[3];
  c->grant.tb.enabled  = true;
  c->grant.L_prb       = u[4];
  c->grant.nof_symb    = u[5];
  c->grant.nof_re      = u[3] / u[1];
  c->uci_cfg.ack[0].nof_acks = u[6];
  c->uci_cfg.cqi.ri_len      = u[7];
  if (u[8]) {
    c->uci_cfg.cqi.data_enable = true;
    c->uci_cfg.cqi.type        = u[8] == 1 ? SRSLTE_CQI_TYPE_WIDEBAND : SRSLTE_CQI_TYPE_SUBBAND_HL;
    c->uci_cfg.cqi.N           = 9;
  }
  c->uci_offset.I_offset_ack = u[9];
  c->uci_offset.I_offset_ri  = u[10];
  c->uci_offset.I_offset_cqi = u[11];
  if (tx)
    c->softbuffers.tx = &t->sb_tx;
  else
    c->softbuffers.rx = &t->sb_rx;
}

/* data: tbs/8 bytes; q_bits: nb_q unpacked bits out (placeholder / repetition positions of ACK and RI hold whatever
 * the encoder left there: pusch.c resolves them while scrambling, the data path does not depend on them) */
int refh_ulsch_encode(refh_tb_t* t, const uint32_t* u, const uint8_t* data, const uint8_t* ack, uint32_t ri, uint8_t* q_bits)
{
  srslte_pusch_cfg_t c;
  srslte_uci_value_t v;
  fill_ul_cfg(t, &c, u, 1);
  memset(&v, 0, sizeof(v));
  for (uint32_t i = 0; i < u[6] && i < SRSLTE_UCI_MAX_ACK_BITS; i++) v.ack.ack_value[i] = ack[i];
  v.ri = (uint8_t)ri;
  v.cqi.subband_hl.wideband_cqi_cw0     = 11;
  v.cqi.subband_hl.subband_diff_cqi_cw0 = 0x2a5a5;
  uint8_t* g = (uint8_t*)calloc(u[3] / 8 + 64, 1);
  uint8_t* q = (uint8_t*)calloc(u[3] / 8 + 64, 1);
  uint8_t* d = (uint8_t*)calloc(u[0] / 8 + 64, 1);
  memcpy(d, data, u[0] / 8);
  if (u[2] == 0) srslte_softbuffer_tx_reset_tbs(&t->sb_tx, u[0]);
  int rc = srslte_ulsch_encode(&t->sch, &c, d, &v, g, q);
  if (rc >= 0) srslte_bit_unpack_vector(q, q_bits, (int)u[3]); /* returns the number of ACK + RI bits */
  free(g);
  free(q);
  free(d);
  return rc;
}

/* q_llr: nb_q descrambled LLRs (copied; the reference modifies them); c_seq: the scrambling bits, one per byte;
 * g_out: nb_q de-interleaved LLRs; out: tbs/8 + 8 bytes; uci_out[4] = { ack0, ack1, ri, cqi crc ok } */
int refh_ulsch_decode(refh_tb_t* t, const uint32_t* u, const int16_t* q_llr, const uint8_t* c_seq, int16_t* g_out,
                      uint8_t* out, uint32_t max_it, float* avg_it, uint8_t* uci_out)
{
  srslte_pusch_cfg_t c;
  srslte_uci_value_t v;
  fill_ul_cfg(t, &c, u, 0);
  memset(&v, 0, sizeof(v));
  int16_t* q = (int16_t*)srslte_vec_malloc(sizeof(int16_t) * (u[3] + 64));
  int16_t* g = (int16_t*)srslte_vec_malloc(sizeof(int16_t) * (u[3] + 64));
  uint8_t* s = (uint8_t*)malloc(u[3] + 64);
  memcpy(q, q_llr, sizeof(int16_t) * u[3]);
  memset(g, 0, sizeof(int16_t) * (u[3] + 64));
  memcpy(s, c_seq, u[3]);
  srslte_sch_set_max_noi(&t->sch, max_it);
  int rc = srslte_ulsch_decode(&t->sch, &c, q, g, s, out, &v);
  if (avg_it) *avg_it = srslte_sch_last_noi(&t->sch);
  memcpy(g_out, g, sizeof(int16_t) * u[3]);
  if (uci_out) {
    uci_out[0] = v.ack.ack_value[0];
    uci_out[1] = v.ack.ack_value[1];
    uci_out[2] = v.ri;
    uci_out[3] = v.cqi.data_crc;
  }
  free(q);
  free(g);
  free(s);
  return rc;
}
