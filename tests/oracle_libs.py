"""ctypes loaders for the CPU checkers (TEST INFRASTRUCTURE).

  port()  -> oracle/libtdec_port.so          this repo's scalar restatement (oracle/tdec_port.c)
  ref()   -> oracle/_ref/libsrslte_ref.so    the reference's own sources compiled by oracle/Makefile
             (None when it has not been built; it is built in the dev container where
             /root/reference exists and travels to the GPU box as a prebuilt file)

Nothing under srslte-emane_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE, "libtdec_port.so")
REF_SO = os.path.join(ORACLE, "_ref", "libsrslte_ref.so")

CRC24A = 0x1864CFB
CRC24B = 0x1800063

_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")

_port = None
_ref = None


def build_port():
    if (not os.path.exists(PORT_SO)
            or os.path.getmtime(PORT_SO) < os.path.getmtime(os.path.join(ORACLE, "tdec_port.c"))):
        subprocess.check_call(["make", "-s", "-C", ORACLE, "port"])


def build_ref():
    """(Re)build oracle/_ref when the reference tree is present; no-op otherwise."""
    if os.path.isdir("/root/reference/lib/src/phy/fec"):
        subprocess.check_call(["make", "-s", "-C", ORACLE, "ref"])
    return os.path.exists(REF_SO)


class PortSoftbuffer(C.Structure):
    _fields_ = [("max_cb", C.c_uint32), ("buffer_f", C.POINTER(C.c_int16)),
                ("data", C.POINTER(C.c_uint8)), ("cb_crc", C.POINTER(C.c_uint8)),
                ("tb_crc", C.c_uint8)]


class PortCbsegm(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("F", "C", "K1", "K2", "K1_idx", "K2_idx", "C1", "C2", "tbs")]


def port():
    global _port
    if _port is not None:
        return _port
    build_port()
    L = C.CDLL(PORT_SO)
    L.port_cb_size.argtypes = [C.c_uint32]
    L.port_cb_index.argtypes = [C.c_uint32]
    L.port_nof_subblocks.argtypes = [C.c_uint32]
    L.port_qpp.argtypes = [C.c_uint32, C.c_uint32]
    L.port_qpp.restype = C.c_uint32
    L.port_crc_bytes.argtypes = [C.c_uint32, _u8p, C.c_uint32]
    L.port_crc_bytes.restype = C.c_uint32
    L.port_crc_bits.argtypes = [C.c_uint32, _u8p, C.c_uint32]
    L.port_crc_bits.restype = C.c_uint32
    L.port_cbsegm.argtypes = [C.POINTER(PortCbsegm), C.c_uint32]
    L.port_rm_rx_table.argtypes = [C.c_uint32, C.c_uint32, C.c_int, _u16p]
    L.port_rm_turbo_rx.argtypes = [_i16p, C.c_uint32, _i16p, C.c_uint32, C.c_uint32, C.c_int]
    L.port_tdec_new.restype = C.c_void_p
    L.port_tdec_free.argtypes = [C.c_void_p]
    L.port_tdec_new_cb.argtypes = [C.c_void_p, C.c_uint32]
    L.port_tdec_iteration.argtypes = [C.c_void_p, _i16p, C.c_int, _u8p]
    L.port_tdec_run_all.argtypes = [C.c_void_p, _i16p, C.c_int, _u8p, C.c_uint32, C.c_uint32]
    L.port_tdec_get_nof_iterations.argtypes = [C.c_void_p]
    L.port_tdec_clamp_count.argtypes = [C.c_void_p]
    L.port_tdec_clamp_count.restype = C.c_uint64
    L.port_tdec_last_llr.argtypes = [C.c_void_p]
    L.port_tdec_last_llr.restype = C.POINTER(C.c_int16)
    L.port_softbuffer_init.argtypes = [C.POINTER(PortSoftbuffer), C.c_uint32]
    L.port_softbuffer_reset.argtypes = [C.POINTER(PortSoftbuffer)]
    L.port_softbuffer_free.argtypes = [C.POINTER(PortSoftbuffer)]
    L.port_decode_tb.argtypes = [C.c_void_p, C.POINTER(PortSoftbuffer), C.c_uint32, C.c_uint32,
                                 C.c_uint32, C.c_uint32, _i16p, _u8p, C.c_uint32,
                                 C.POINTER(C.c_float), _u32p]
    L.port_batch_run_all.argtypes = [_i16p, C.c_uint32, C.c_int, _u8p, C.c_uint32, C.c_uint32,
                                     C.c_uint32, C.c_uint32, C.c_uint32]
    L.port_gold_sequence.argtypes = [C.c_uint32, C.c_uint32, _u8p]
    L.port_demod_s.argtypes = [C.c_int, _f32p, _i16p, C.c_uint32]
    L.port_descramble_s.argtypes = [_i16p, _u8p, C.c_uint32]
    L.port_demod_descramble.argtypes = [C.c_int, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, _i16p]
    L.port_ulsch_deinterleave.argtypes = [_i16p, C.c_uint32, C.c_uint32, C.c_uint32, _i16p]
    L.port_ulsch_deinterleave.restype = None
    L.port_uci_q_prime_ri_ack.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float]
    L.port_uci_q_prime_ri_ack.restype = C.c_uint32
    L.port_uci_q_prime_cqi.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32]
    L.port_uci_q_prime_cqi.restype = C.c_uint32
    L.port_ulsch_demux.argtypes = [_i16p, _u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _i16p]
    _port = L
    return L


def ref():
    global _ref
    if _ref is not None:
        return _ref
    if not os.path.exists(REF_SO):
        if not build_ref():
            return None
    L = C.CDLL(REF_SO)
    for n in ("refh_sizeof_tdec", "refh_sizeof_sch", "refh_sizeof_crc", "refh_sizeof_softbuffer_rx",
              "refh_sizeof_cbsegm", "refh_offsetof_tdec_n_iter", "refh_offsetof_sch_decoder"):
        getattr(L, n).restype = C.c_size_t
    L.refh_batch_run_all.argtypes = [_i16p, C.c_uint32, C.c_int, _u8p, C.c_uint32, C.c_uint32,
                                     C.c_uint32, C.c_uint32, C.c_uint32]
    if hasattr(L, "refh_pool_create"):
        L.refh_pool_create.argtypes = [C.c_uint32]
        L.refh_pool_create.restype = C.c_void_p
        L.refh_pool_run.argtypes = [C.c_void_p, _i16p, C.c_uint32, C.c_int, _u8p, C.c_uint32, C.c_uint32, C.c_uint32,
                                    C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.refh_pool_destroy.argtypes = [C.c_void_p]
    L.refh_tdec_trace.argtypes = [_i16p, C.c_int, C.c_uint32, C.c_uint32, _u8p, C.c_void_p]
    L.refh_tcod_encode.argtypes = [_u8p, _u8p, C.c_uint32]
    L.refh_rm_turbo_tx.argtypes = [_u8p, C.c_uint32, _u8p, C.c_uint32, C.c_uint32]
    L.refh_tb_new.restype = C.c_void_p
    L.refh_tb_free.argtypes = [C.c_void_p]
    L.refh_tb_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, _u8p]
    L.refh_tb_rx_reset.argtypes = [C.c_void_p, C.c_uint32]
    L.refh_tb_decode.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _i16p,
                                 _u8p, C.c_uint32, C.POINTER(C.c_float), C.c_void_p]
    L.refh_tb_softbuffer.argtypes = [C.c_void_p, C.c_uint32]
    L.refh_tb_softbuffer.restype = C.POINTER(C.c_int16)
    L.refh_crc_bytes.argtypes = [C.c_uint32, _u8p, C.c_uint32]
    L.refh_crc_bytes.restype = C.c_uint32
    L.refh_crc_bits.argtypes = [C.c_uint32, _u8p, C.c_uint32]
    L.refh_crc_bits.restype = C.c_uint32
    L.srslte_rm_turbo_gentables.argtypes = []
    L.srslte_rm_turbo_rx_lut.argtypes = [_i16p, _i16p, C.c_uint32, C.c_uint32, C.c_uint32]
    L.srslte_rm_turbo_rx_lut_.argtypes = [_i16p, _i16p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_bool]
    L.srslte_cbsegm.argtypes = [C.POINTER(PortCbsegm), C.c_uint32]  # identical 9 x uint32 layout
    L.srslte_cbsegm_cbindex.argtypes = [C.c_uint32]
    L.srslte_tdec_autoimp_get_subblocks.argtypes = [C.c_uint32]
    L.srslte_tdec_autoimp_get_subblocks.restype = C.c_uint32
    L.refh_demod_descramble.argtypes = [C.c_int, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, _i16p, C.c_uint32]
    L.refh_ulsch_deinterleave.argtypes = [_i16p, C.c_uint32, C.c_uint32, C.c_uint32, _i16p]
    L.refh_ulsch_encode.argtypes = [C.c_void_p, _u32p, _u8p, _u8p, C.c_uint32, _u8p]
    L.refh_ulsch_decode.argtypes = [C.c_void_p, _u32p, _i16p, _u8p, _i16p, _u8p, C.c_uint32, C.POINTER(C.c_float), _u8p]
    L.srslte_rm_turbo_gentables()
    _ref = L
    return L


ALL_K = (list(range(40, 513, 8)) + list(range(528, 1025, 16)) + list(range(1056, 2049, 32))
         + list(range(2112, 6145, 64)))


# ------------------------------------------------------------------------------------------
# convenience wrappers
# ------------------------------------------------------------------------------------------
def port_run_all(llr, K, nit, natural=True):
    """llr: [n, len] int16 -> bytes [n, K/8] through the scalar port."""
    P = port()
    llr = np.ascontiguousarray(llr, dtype=np.int16)
    n = llr.shape[0]
    out = np.zeros((n, K // 8), np.uint8)
    rc = P.port_batch_run_all(llr.reshape(-1), llr.shape[1], int(natural), out.reshape(-1), K // 8, n, K, nit,
                              min(n, os.cpu_count() or 1))
    assert rc == 0
    return out


def ref_run_all(llr, K, nit, natural=True, threads=None):
    R = ref()
    llr = np.ascontiguousarray(llr, dtype=np.int16)
    n = llr.shape[0]
    out = np.zeros((n, K // 8), np.uint8)
    rc = R.refh_batch_run_all(llr.reshape(-1), llr.shape[1], int(natural), out.reshape(-1), K // 8, n, K, nit,
                              threads or min(n, os.cpu_count() or 1))
    assert rc == 0
    return out


class RefPool:
    """Persistent pool of reference decoders: one srslte_tdec_t per pthread, created once (outside any timing), the
    way the reference's own turbodecoder_test initialises once and loops.  run() returns (bytes, wall seconds of the
    decode loops, summed per-thread loop seconds)."""

    def __init__(self, threads):
        self.R = ref()
        self.threads = int(threads)
        self.h = self.R.refh_pool_create(self.threads)
        assert self.h

    def run(self, llr, K, nit, natural=True):
        llr = np.ascontiguousarray(llr, dtype=np.int16)
        n = llr.shape[0]
        out = np.zeros((n, K // 8), np.uint8)
        wall, busy = C.c_double(0), C.c_double(0)
        rc = self.R.refh_pool_run(self.h, llr.reshape(-1), llr.shape[1], int(natural), out.reshape(-1), K // 8, n, K, nit,
                                  C.byref(wall), C.byref(busy))
        assert rc == 0
        return out, wall.value, busy.value

    def close(self):
        if self.h:
            self.R.refh_pool_destroy(self.h)
            self.h = None


def port_trace(llr1, K, nit, natural=True):
    """per-half-iteration decisions [nit, K/8] and soft outputs [nit, K] (natural order)."""
    P = port()
    h = P.port_tdec_new()
    try:
        assert P.port_tdec_new_cb(h, K) == 0
        llr1 = np.ascontiguousarray(llr1, np.int16)
        by = np.zeros((nit, K // 8), np.uint8)
        so = np.zeros((nit, K), np.int16)
        for it in range(nit):
            P.port_tdec_iteration(h, llr1, int(natural), by[it])
            so[it] = np.ctypeslib.as_array(P.port_tdec_last_llr(h), (K,))
        return by, so, int(P.port_tdec_clamp_count(h))
    finally:
        P.port_tdec_free(h)


def ref_trace(llr1, K, nit, natural=True):
    """same through the compiled reference; soft outputs are de-permuted to natural order."""
    R = ref()
    llr1 = np.ascontiguousarray(llr1, np.int16)
    by = np.zeros((nit, K // 8), np.uint8)
    so = np.zeros((nit, K), np.int16)
    assert R.refh_tdec_trace(llr1, int(natural), K, nit, by.reshape(-1), so.ctypes.data) == 0
    W = int(R.srslte_tdec_autoimp_get_subblocks(K))
    if W:
        L = K // W
        n = np.arange(K)
        so = so[:, (n % L) * W + n // L]
    return by, so


# ---- front end (soft demodulation + descrambling) ---------------------------------
REF_MOD = {2: 1, 4: 2, 6: 3, 8: 4}   # bits per symbol -> srslte_mod_t


def port_demod_descramble(qm, sym, c_init, nof_bits=None):
    """sym: complex64 [n] -> int16 [qm * n] through the port (nof_bits defaults to all of them)."""
    sym = np.ascontiguousarray(sym, dtype=np.complex64)
    n = sym.shape[0]
    nb = qm * n if nof_bits is None else nof_bits
    llr = np.zeros(qm * n, np.int16)
    rc = port().port_demod_descramble(qm, sym.view(np.float32), n, c_init, nb, llr)
    assert rc == 0
    return llr


def ref_demod_descramble(qm, sym, c_init, nof_bits=None):
    """the same through the reference's own srslte_demod_soft_demodulate_s + srslte_scrambling_s_offset."""
    sym = np.ascontiguousarray(sym, dtype=np.complex64)
    n = sym.shape[0]
    nb = qm * n if nof_bits is None else nof_bits
    llr = np.zeros(qm * n, np.int16)
    rc = ref().refh_demod_descramble(REF_MOD[qm], sym.view(np.float32), n, c_init, nb, llr, qm * n)
    assert rc == 0
    return llr


def port_ulsch_deinterleave(q, qm, n_symb):
    """q: int16 [H' * qm] in channel order -> g in UL-SCH order (no UCI), through the port."""
    q = np.ascontiguousarray(q, dtype=np.int16)
    g = np.zeros_like(q)
    port().port_ulsch_deinterleave(q, qm, q.size // qm, n_symb, g)
    return g


def ref_ulsch_deinterleave(q, qm, n_symb):
    q = np.ascontiguousarray(q, dtype=np.int16)
    g = np.zeros_like(q)
    assert ref().refh_ulsch_deinterleave(q, qm, q.size // qm, n_symb, g) == 0
    return g


# ---- UL-SCH with multiplexed UCI (data path; sch.c:920-1064) ------------------------------------------
# 36.213 tables 8.6.3-1/2/3 as the reference holds them (sch.c:43-52)
BETA_HARQ = [2.0, 2.5, 3.125, 4.0, 5.0, 6.250, 8.0, 10.0, 12.625, 15.875, 20.0, 31.0, 50.0, 80.0, 126.0, -1.0]
BETA_RI = [1.25, 1.625, 2.0, 2.5, 3.125, 4.0, 5.0, 6.25, 8.0, 10.0, 12.625, 15.875, 20.0, -1.0, -1.0, -1.0]
BETA_CQI = [-1.0, -1.0, 1.125, 1.25, 1.375, 1.625, 1.750, 2.0, 2.25, 2.5, 2.875, 3.125, 3.5, 4.0, 5.0, 6.25]
CQI_LEN = {0: 0, 1: 4, 2: 22}   # cqi_mode of the harness -> payload bits (srslte_cqi_size)


def ul_cfg(tbs, qm, rv, nb_q, l_prb, nof_symb, nof_ack=0, ri_len=0, cqi_mode=0, i_ack=5, i_ri=5, i_cqi=7):
    return np.array([tbs, qm, rv, nb_q, l_prb, nof_symb, nof_ack, ri_len, cqi_mode, i_ack, i_ri, i_cqi], np.uint32)


def port_uci_q_primes(u):
    """(Q'_ack, Q'_ri, Q'_cqi) of an ul_cfg through the port's restatement of uci.c:266-283, 547-571."""
    P = port()
    tbs, qm, rv, nb_q, l_prb, nof_symb, nof_ack, ri_len, cqi_mode, i_ack, i_ri, i_cqi = (int(x) for x in u)
    seg = PortCbsegm()
    assert P.port_cbsegm(C.byref(seg), tbs) == 0
    k_segm = seg.C1 * seg.K1 + seg.C2 * seg.K2
    o_cqi = CQI_LEN[cqi_mode]
    q_ack = P.port_uci_q_prime_ri_ack(nof_ack, k_segm, l_prb, nof_symb, BETA_HARQ[i_ack]) if nof_ack else 0
    q_ri = P.port_uci_q_prime_ri_ack(ri_len, k_segm, l_prb, nof_symb, BETA_RI[i_ri]) if ri_len else 0
    q_cqi = P.port_uci_q_prime_cqi(o_cqi, k_segm, l_prb, nof_symb, BETA_CQI[i_cqi], q_ri) if o_cqi else 0
    return q_ack, q_ri, q_cqi


def port_ulsch_demux(q, c_seq, qm, n_symb, q_ack, q_ri, ri_len):
    """descrambled LLRs in channel order -> UL-SCH order with ACK erasure and RI skipping (g[0] quirk included);
    returns the (H' - Q'_ri) * qm defined entries."""
    q = np.ascontiguousarray(q, np.int16)
    g = np.zeros_like(q)
    rc = port().port_ulsch_demux(q, np.ascontiguousarray(c_seq, np.uint8), qm, q.size // qm, n_symb, q_ack, q_ri, ri_len, g)
    assert rc == 0
    return g[: q.size - q_ri * qm]


def port_ulsch_decode(u, q, c_seq, max_it):
    """The data path of srslte_ulsch_decode through the port: Q' counts, de-multiplexing, decode_tb on the data part.
    Returns (ret, bytes[tbs/8 + 3], avg half iterations, g, (Q'_ack, Q'_ri, Q'_cqi))."""
    P = port()
    tbs, qm, rv, nb_q, l_prb, nof_symb, nof_ack, ri_len = (int(x) for x in u[:8])
    q_ack, q_ri, q_cqi = port_uci_q_primes(u)
    g = port_ulsch_demux(q, c_seq, qm, nof_symb, q_ack, q_ri, ri_len)
    seg = PortCbsegm()
    assert P.port_cbsegm(C.byref(seg), tbs) == 0
    sb = PortSoftbuffer()
    P.port_softbuffer_init(C.byref(sb), seg.C)
    dec = P.port_tdec_new()
    out = np.zeros(tbs // 8 + 8, np.uint8)
    avg = C.c_float()
    noi = np.zeros(seg.C, np.uint32)
    G = nb_q - (q_ri + q_cqi) * qm
    rc = P.port_decode_tb(dec, C.byref(sb), tbs, qm, rv, G, np.ascontiguousarray(g[q_cqi * qm:]), out, max_it,
                          C.byref(avg), noi)
    P.port_softbuffer_free(C.byref(sb))
    P.port_tdec_free(dec)
    return rc, out[: tbs // 8 + 3].copy(), avg.value, g, (q_ack, q_ri, q_cqi)
