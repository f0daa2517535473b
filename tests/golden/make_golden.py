"""Generate the committed golden fixtures from the reference's OWN compiled code (oracle/_ref).

Run in the dev container (where /root/reference exists and `make -C oracle ref` has been run):
    python tests/golden/make_golden.py
Outputs small .npz files next to this script.  The GPU box only reads them.

Fixtures:
  tdec_vectors.npz   noisy int16 LLR blocks (natural order) for several K in all three decoder regimes,
                     with the reference's decoded bytes after every half iteration 1..10
                     (srslte_tdec_new_cb + srslte_tdec_iteration), incl. a saturating (scale 700) case
  rm_tables.npz      srslte_rm_turbo_rx_lut scatter results for a ramp input: pins the receive index tables
  kat.npz            the reference tests' own known-answer data: crc_test (srand(1), 5001 bits ->
                     0x1C5C97 / 0x36D1F0) and turbodecoder_test.h known_data / known_data_encoded (K=504)
  tb_vectors.npz     transport-block cases through srslte_dlsch_encode2 / srslte_dlsch_decode2
"""
import ctypes as C
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import oracle_libs as ol  # noqa: E402

vec = ge.load_package().vectors
R = ol.ref()
assert R is not None, "build oracle/_ref first (make -C oracle ref)"


def tdec_vectors():
    out = {}
    cases = [(40, 1.092, 100), (104, 1.457, 100), (400, 1.092, 100), (408, 1.092, 100), (512, 0.9, 700),
             (800, 1.457, 100), (816, 1.092, 100), (1024, 0.9, 700), (2048, 1.092, 100), (5824, 1.092, 100),
             (6144, 1.457, 100), (6144, 1.092, 100), (6144, 0.9, 700)]
    for ci, (K, sigma, scale) in enumerate(cases):
        bits, llr = vec.make_blocks(2, K, sigma, scale, seed=1000 + ci)
        dec = np.zeros((2, 10, K // 8), np.uint8)
        for i in range(2):
            dec[i], _ = ol.ref_trace(llr[i], K, 10)
        out[f"c{ci}_K"] = np.array([K, int(scale)], np.int32)
        out[f"c{ci}_bits"] = np.packbits(bits, axis=1)
        out[f"c{ci}_llr"] = llr
        out[f"c{ci}_dec"] = dec
    np.savez_compressed(os.path.join(HERE, "tdec_vectors.npz"), **out)


def rm_tables():
    out = {}
    for K in (40, 400, 408, 512, 816, 1024, 6144):
        idx = ol.ALL_K.index(K)
        N = 3 * K + 12
        for rv in range(4):
            ramp = (np.arange(N) + 1).astype(np.int16)  # all distinct and non-zero (N <= 18444 < 32767)
            for sb in (0, 1):
                buf = np.zeros(18600, np.int16)
                assert R.srslte_rm_turbo_rx_lut_(ramp.copy(), buf, N, idx, rv, bool(sb)) == 0
                # table[i] = position that received value i+1
                pos = np.zeros(N, np.uint16)
                nz = np.nonzero(buf)[0]
                pos[buf[nz].astype(np.int64) - 1] = nz
                out[f"K{K}_rv{rv}_sb{sb}"] = pos
    np.savez_compressed(os.path.join(HERE, "rm_tables.npz"), **out)


def kat():
    libc = C.CDLL("libc.so.6")
    libc.srand(1)
    bits = np.array([libc.rand() % 2 for _ in range(5001)], np.uint8)  # crc_test.c:98-100 with -s 1
    a = R.refh_crc_bits(ol.CRC24A, bits.copy(), 5001)
    b = R.refh_crc_bits(ol.CRC24B, bits.copy(), 5001)
    assert a == 0x1C5C97 and b == 0x36D1F0, (hex(a), hex(b))  # crc_test.h:37-38
    src = open("/root/reference/lib/src/phy/fec/test/turbodecoder_test.h").read()

    def arr(name):
        m = re.search(name + r"\[[^\]]*\]\s*=\s*\{(.*?)\};", src, re.S)
        return np.array([int(x) for x in re.findall(r"\d+", m.group(1))], np.uint8)

    kd, ke = arr("known_data"), arr("known_data_encoded")
    assert kd.size == 504 and ke.size == 3 * 504 + 12, (kd.size, ke.size)
    np.savez_compressed(os.path.join(HERE, "kat.npz"), crc_bits=np.packbits(bits), crc_nbits=np.array([5001]),
                        crc24a=np.array([a], np.uint32), crc24b=np.array([b], np.uint32),
                        known_data=kd, known_data_encoded=ke)


def tb_vectors():
    out = {}
    t = R.refh_tb_new()
    cases = [(75376, 6, 90000, 0.0, 700), (75376, 6, 90000, 0.45, 100), (6200, 4, 9600, 0.45, 100),
             (2216, 2, 4800, 0.7, 100), (14112, 4, 28800, 0.7, 100)]
    for ci, (tbs, qm, G, sigma, scale) in enumerate(cases):
        rng = np.random.default_rng(2000 + ci)
        data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        R.refh_tb_rx_reset(t, tbs)
        seg = ol.PortCbsegm()
        R.srslte_cbsegm(C.byref(seg), tbs)
        for rv in (0, 2):
            e = np.zeros(G, np.uint8)
            assert R.refh_tb_encode(t, tbs, qm, rv, G, data, e) == 0
            llr = vec.awgn_llr(e, sigma, scale, rng)
            o = np.zeros(tbs // 8 + 8, np.uint8)
            avg = C.c_float()
            crc = np.zeros(seg.C, np.uint8)
            rc = R.refh_tb_decode(t, tbs, qm, rv, G, llr, o, 8, C.byref(avg), crc.ctypes.data)
            out[f"t{ci}_rv{rv}_llr"] = llr
            out[f"t{ci}_rv{rv}_out"] = o[: tbs // 8 + 3]
            out[f"t{ci}_rv{rv}_res"] = np.array([rc, round(avg.value * seg.C)], np.int32)
            out[f"t{ci}_rv{rv}_cbcrc"] = crc
        out[f"t{ci}_par"] = np.array([tbs, qm, G, 8], np.int32)
        out[f"t{ci}_data"] = data
    R.refh_tb_free(t)
    np.savez_compressed(os.path.join(HERE, "tb_vectors.npz"), **out)


if __name__ == "__main__":
    tdec_vectors()
    rm_tables()
    kat()
    tb_vectors()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
