"""The reference's own entry points (srslte_tdec_*, srslte_rm_turbo_rx_lut, decode_tb) exported by
libsrslte_b200.so, driven the way the reference's tests drive them (turbodecoder_test.c:190-266,
rm_turbo_test.c:171-189, sch.c:299-500).  B200 only."""
import ctypes as C

import numpy as np
import pytest

import oracle_libs as ol

pytestmark = pytest.mark.gpu

_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def L(pkg):
    lib = pkg.lib()
    lib.srslte_tdec_init.argtypes = [C.c_void_p, C.c_uint32]
    lib.srslte_tdec_free.argtypes = [C.c_void_p]
    lib.srslte_tdec_free.restype = None
    lib.srslte_tdec_force_not_sb.argtypes = [C.c_void_p]
    lib.srslte_tdec_force_not_sb.restype = None
    lib.srslte_tdec_new_cb.argtypes = [C.c_void_p, C.c_uint32]
    lib.srslte_tdec_get_nof_iterations.argtypes = [C.c_void_p]
    lib.srslte_tdec_iteration.argtypes = [C.c_void_p, _i16p, _u8p]
    lib.srslte_tdec_iteration.restype = None
    lib.srslte_tdec_run_all.argtypes = [C.c_void_p, _i16p, _u8p, C.c_uint32, C.c_uint32]
    lib.srslte_rm_turbo_rx_lut.argtypes = [_i16p, _i16p, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.srslte_rm_turbo_rx_lut_.argtypes = [_i16p, _i16p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_bool]
    return lib


def test_turbodecoder_test_flow(L, vec):
    """srslte_tdec_init + force_not_sb + run_all on natural-order LLRs, like turbodecoder_test."""
    h = C.create_string_buffer(18264)
    assert L.srslte_tdec_init(h, 6144) == 0
    L.srslte_tdec_force_not_sb(h)
    for K in (40, 504, 1024, 6144):
        bits, llr = vec.make_blocks(3, K, vec.harness_sigma(4.0), 100, seed=K)
        for nit in (1, 4, 10):
            for i in range(3):
                out = np.zeros(K // 8, np.uint8)
                assert L.srslte_tdec_run_all(h, llr[i].copy(), out, nit, K) == 0
                assert np.array_equal(out, ol.port_run_all(llr[i:i + 1], K, nit)[0]), (K, nit, i)
                assert L.srslte_tdec_get_nof_iterations(h) == nit
    assert L.srslte_tdec_run_all(h, np.zeros(3 * 6144 + 12, np.int16), np.zeros(768, np.uint8), 1, 6145) == -1
    L.srslte_tdec_free(h)
    assert h.raw == b"\x00" * 18264          # srslte_tdec_free zeroes the handle like the reference


def test_manual_decoder_selection(L, vec):
    """srslte_tdec_init_manual (turbodecoder.c:146-260, turbodecoder_test -d): the three 16-bit decoders AUTO itself uses
    can be pinned; a pinned handle decodes the block sizes for which AUTO picks the same decoder (same result) and
    refuses the others; the non-windowed SSE decoder and the 8-bit decoders are refused at init."""
    L.srslte_tdec_init_manual.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
    GENERIC, SSE, SSE_WINDOW, AVX_WINDOW, SSE8_WINDOW = 1, 2, 3, 5, 6
    for typ, good, bad in ((GENERIC, (40, 400), (408, 6144)), (SSE_WINDOW, (408, 512, 800), (400, 816)),
                           (AVX_WINDOW, (816, 5824, 6144), (800, 40))):
        h = C.create_string_buffer(18264)
        assert L.srslte_tdec_init_manual(h, 6144, typ) == 0
        L.srslte_tdec_force_not_sb(h)
        for K in good:
            _bits, llr = vec.make_blocks(2, K, vec.harness_sigma(3.0), 100, seed=K + typ)
            for i in range(2):
                out = np.zeros(K // 8, np.uint8)
                assert L.srslte_tdec_run_all(h, llr[i].copy(), out, 4, K) == 0
                assert np.array_equal(out, ol.port_run_all(llr[i:i + 1], K, 4)[0]), (typ, K, i)
        for K in bad:
            assert L.srslte_tdec_run_all(h, np.zeros(3 * K + 12, np.int16), np.zeros(K // 8, np.uint8), 2, K) == -1
            assert L.srslte_tdec_new_cb(h, K) == -1
        L.srslte_tdec_free(h)
    for typ in (SSE, SSE8_WINDOW, 99):
        h = C.create_string_buffer(18264)
        assert L.srslte_tdec_init_manual(h, 6144, typ) != 0


def test_iteration_by_iteration_sub_block_input(L, vec):
    """sch.c style: srslte_tdec_new_cb, then one srslte_tdec_iteration per call on the soft-buffer layout."""
    h = C.create_string_buffer(18264)
    assert L.srslte_tdec_init(h, 6144) == 0
    for K in (408, 5824, 104):
        bits, llr = vec.make_blocks(1, K, vec.harness_sigma(4.0), 100, seed=K)
        sb = vec.sb_layout_from_natural(llr, K)[0]
        buf = np.zeros(18600, np.int16)
        buf[: sb.size] = sb
        want, _, _ = ol.port_trace(sb, K, 6, natural=False)
        assert L.srslte_tdec_new_cb(h, K) == 0
        for it in range(6):
            out = np.zeros(K // 8, np.uint8)
            L.srslte_tdec_iteration(h, buf, out)
            assert np.array_equal(out, want[it]), (K, it)
            assert L.srslte_tdec_get_nof_iterations(h) == it + 1
        if K > 400:   # the tail copies the reference leaves in the caller's pads (turbodecoder_iter.h:56-65)
            assert buf[K] == sb[3 * (K + 32)] and buf[K + 32 + K] == sb[3 * (K + 32) + 1]
    L.srslte_tdec_free(h)


def test_rm_turbo_rx_lut_like_rm_turbo_test(L):
    P = ol.port()
    rng = np.random.default_rng(3)
    for K in (40, 512, 1024, 6144):
        idx = ol.ALL_K.index(K)
        for rv in range(4):
            for E in (1920, 8192, 3 * K + 12, 5 * K):
                e = rng.integers(-5, 5, E).astype(np.int16)        # rm_turbo_test uses small random ints
                for sb in (True, False):
                    got = rng.integers(-100, 100, 18600).astype(np.int16)
                    want = got.copy()
                    assert L.srslte_rm_turbo_rx_lut_(e, got, E, idx, rv, sb) == 0
                    assert P.port_rm_turbo_rx(e, E, want, K, rv, int(sb)) == 0
                    assert np.array_equal(got, want), (K, rv, E, sb)
    assert L.srslte_rm_turbo_rx_lut(np.zeros(8, np.int16), np.zeros(18600, np.int16), 8, 188, 0) == -2
    assert L.srslte_rm_turbo_rx_lut(np.zeros(8, np.int16), np.zeros(18600, np.int16), 8, 0, 4) == -2


class SoftbufferRx(C.Structure):
    _fields_ = [("max_cb", C.c_uint32), ("buffer_f", C.POINTER(C.POINTER(C.c_int16))),
                ("data", C.POINTER(C.POINTER(C.c_uint8))), ("cb_crc", C.POINTER(C.c_bool)), ("tb_crc", C.c_bool)]


def _make_softbuffer(max_cb):
    bufs = [np.zeros(18600, np.int16) for _ in range(max_cb)]
    datas = [np.zeros(768, np.uint8) for _ in range(max_cb)]
    crc = (C.c_bool * max_cb)()
    pf = (C.POINTER(C.c_int16) * max_cb)(*[b.ctypes.data_as(C.POINTER(C.c_int16)) for b in bufs])
    pd = (C.POINTER(C.c_uint8) * max_cb)(*[d.ctypes.data_as(C.POINTER(C.c_uint8)) for d in datas])
    sb = SoftbufferRx(max_cb, pf, pd, crc, False)
    return sb, (bufs, datas, crc, pf, pd)


def test_decode_tb_with_host_softbuffer(L, golden):
    """srslte_b200_sch_decode_tb on a MAC-style host soft buffer: rv 0 then rv 2, vs the reference's
    srslte_dlsch_decode2 results (incl. LLR accumulation visible in the host buffer)."""
    L.srslte_b200_sch_decode_tb.argtypes = [C.POINTER(SoftbufferRx), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                            _i16p, _u8p, C.c_uint32, C.POINTER(C.c_float)]
    g = golden["tb_vectors"]
    P = ol.port()
    for c in sorted({k.split("_")[0] for k in g if k.startswith("t")}, key=lambda s: int(s[1:])):
        tbs, qm, G, max_it = (int(x) for x in g[f"{c}_par"])
        seg = ol.PortCbsegm()
        P.port_cbsegm(C.byref(seg), tbs)
        sb, keep = _make_softbuffer(13)
        psb = ol.PortSoftbuffer()
        P.port_softbuffer_init(C.byref(psb), 13)
        dec = P.port_tdec_new()
        for rv in (0, 2):
            llr = g[f"{c}_rv{rv}_llr"]
            out = np.zeros(tbs // 8 + 8, np.uint8)
            avg = C.c_float()
            rc = L.srslte_b200_sch_decode_tb(C.byref(sb), tbs, qm, rv, G, llr.copy(), out, max_it, C.byref(avg))
            want_rc, want_its = (int(x) for x in g[f"{c}_rv{rv}_res"])
            assert rc == want_rc and round(avg.value * seg.C) == want_its, (c, rv)
            assert np.array_equal(out[: tbs // 8 + 3], g[f"{c}_rv{rv}_out"]), (c, rv)
            assert [bool(x) for x in keep[2][: seg.C]] == [bool(x) for x in g[f"{c}_rv{rv}_cbcrc"]]
            # host-visible LLR state equals the oracle's soft buffer (outside the pads)
            o2 = np.zeros(tbs // 8 + 8, np.uint8)
            P.port_decode_tb(dec, C.byref(psb), tbs, qm, rv, G, llr, o2, max_it, None, np.zeros(13, np.uint32))
            for cb in range(seg.C):
                K = seg.K1 if cb < seg.C1 else seg.K2
                want_llr = np.ctypeslib.as_array(psb.buffer_f, (13 * 18600,))[cb * 18600:(cb + 1) * 18600]
                if K > 400:
                    for j in range(3):
                        a = j * (K + 32)
                        assert np.array_equal(keep[0][cb][a:a + K], want_llr[a:a + K]), (c, rv, cb, j)
                    a = 3 * (K + 32)
                    assert np.array_equal(keep[0][cb][a:a + 12], want_llr[a:a + 12])
                else:
                    assert np.array_equal(keep[0][cb][: 3 * K + 12], want_llr[: 3 * K + 12])
        P.port_tdec_free(dec)
        P.port_softbuffer_free(C.byref(psb))


def test_device_resident_softbuffer_api(L, golden):
    """softbuffer.h:52-66 through this library: srslte_softbuffer_rx_init / reset_tbs / reset_cb / free keep the HARQ state
    on the device.  rv 0 then rv 2 on the same buffer against the reference's srslte_dlsch_decode2 goldens,
    then a reset and rv 0 again (must repeat the first result: nothing left of the combined LLRs, flags cleared)."""
    L.srslte_b200_sch_decode_tb.argtypes = [C.POINTER(SoftbufferRx), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                            _i16p, _u8p, C.c_uint32, C.POINTER(C.c_float)]
    for f in ("srslte_softbuffer_rx_reset", "srslte_softbuffer_rx_free"):
        getattr(L, f).argtypes = [C.POINTER(SoftbufferRx)]
        getattr(L, f).restype = None
    L.srslte_softbuffer_rx_init.argtypes = [C.POINTER(SoftbufferRx), C.c_uint32]
    L.srslte_softbuffer_rx_reset_tbs.argtypes = [C.POINTER(SoftbufferRx), C.c_uint32]
    L.srslte_softbuffer_rx_reset_tbs.restype = None
    L.srslte_softbuffer_rx_reset_cb.argtypes = [C.POINTER(SoftbufferRx), C.c_uint32]
    L.srslte_softbuffer_rx_reset_cb.restype = None
    g = golden["tb_vectors"]
    P = ol.port()
    sbs = []
    for nprb in (100, 6, 110, 50):
        sb = SoftbufferRx()
        assert L.srslte_softbuffer_rx_init(C.byref(sb), nprb) == 0
        assert sb.max_cb == {100: 16, 6: 1, 110: 16, 50: 8}[nprb]
        sbs.append(sb)
    assert L.srslte_softbuffer_rx_init(C.byref(SoftbufferRx()), 0) != 0
    sb = sbs[0]
    for c in sorted({k.split("_")[0] for k in g if k.startswith("t")}, key=lambda s: int(s[1:])):
        tbs, qm, G, max_it = (int(x) for x in g[f"{c}_par"])
        seg = ol.PortCbsegm()
        P.port_cbsegm(C.byref(seg), tbs)
        L.srslte_softbuffer_rx_reset_tbs(C.byref(sb), tbs)
        first = None
        for rv in (0, 2, -1):
            if rv < 0:                        # new transmission of the same TB: reset, rv 0 again
                L.srslte_softbuffer_rx_reset_cb(C.byref(sb), seg.C)
                assert not any(sb.cb_crc[i] for i in range(sb.max_cb)) and not sb.tb_crc
                rv = 0
            llr = g[f"{c}_rv{rv}_llr"]
            out = np.zeros(tbs // 8 + 8, np.uint8)
            avg = C.c_float()
            rc = L.srslte_b200_sch_decode_tb(C.byref(sb), tbs, qm, rv, G, llr.copy(), out, max_it, C.byref(avg))
            if first is not None and rv == 0:
                assert (rc, avg.value) == first[:2] and np.array_equal(out, first[2]), c
                continue
            want_rc, want_its = (int(x) for x in g[f"{c}_rv{rv}_res"])
            assert rc == want_rc and round(avg.value * seg.C) == want_its, (c, rv)
            assert np.array_equal(out[: tbs // 8 + 3], g[f"{c}_rv{rv}_out"]), (c, rv)
            assert [bool(sb.cb_crc[i]) for i in range(seg.C)] == [bool(x) for x in g[f"{c}_rv{rv}_cbcrc"]]
            assert bool(sb.tb_crc) == all(bool(sb.cb_crc[i]) for i in range(seg.C))      # sch.c:396-399
            if rv == 0:
                first = (rc, avg.value, out.copy())
    for sb in sbs:
        L.srslte_softbuffer_rx_free(C.byref(sb))
        assert sb.max_cb == 0


def test_ulsch_decode_entry_without_control_information(L, golden):
    """srslte_ulsch_decode (sch.h:109-115) through the library alone, with the reference's srslte_pusch_cfg_t / srslte_sch_t
    layouts built by hand: the golden transport blocks of the reference, put into channel order by the inverse of the
    UL-SCH de-interleaver, must come back as the reference's srslte_dlsch_decode2 results (decode_tb is shared by both
    directions, sch.c:1058-1062), g_bits must hold the de-interleaved LLRs, K_segm must be set.  (With multiplexed
    ACK / RI / CQI the entry needs the reference's UCI decoders: tests/test_gpu_relink.py.)"""
    L.srslte_ulsch_decode.argtypes = [C.c_void_p, C.c_void_p, _i16p, _i16p, C.c_void_p, _u8p, C.c_void_p]
    L.srslte_ulsch_decode.restype = C.c_int
    L.srslte_softbuffer_rx_init.argtypes = [C.POINTER(SoftbufferRx), C.c_uint32]
    L.srslte_softbuffer_rx_reset_tbs.argtypes = [C.POINTER(SoftbufferRx), C.c_uint32]
    L.srslte_softbuffer_rx_reset_tbs.restype = None
    L.srslte_softbuffer_rx_free.argtypes = [C.POINTER(SoftbufferRx)]
    L.srslte_softbuffer_rx_free.restype = None
    g = golden["tb_vectors"]
    P = ol.port()
    sb = SoftbufferRx()
    assert L.srslte_softbuffer_rx_init(C.byref(sb), 100) == 0
    q = np.zeros(490792 + 7480, np.uint8)             # srslte_sch_t
    ran = 0
    for c in sorted({k.split("_")[0] for k in g if k.startswith("t")}, key=lambda s: int(s[1:])):
        tbs, qm, G, max_it = (int(x) for x in g[f"{c}_par"])
        cols = next((n for n in (12, 11, 10, 9) if (G // qm) % n == 0), 0)
        if G % qm or not cols:
            continue
        rows = G // qm // cols
        seg = ol.PortCbsegm()
        P.port_cbsegm(C.byref(seg), tbs)
        q[:4].view(np.uint32)[0] = max_it
        L.srslte_softbuffer_rx_reset_tbs(C.byref(sb), tbs)
        for rv in (0, 2):
            llr = g[f"{c}_rv{rv}_llr"][:G]
            # channel order: the matrix is sent column by column
            chan = np.ascontiguousarray(llr.reshape(rows, cols, qm).transpose(1, 0, 2)).reshape(-1).copy()
            cfg = np.zeros(520, np.uint8)            # srslte_pusch_cfg_t, offsets: tests/test_compat_abi.py
            w = cfg.view(np.uint32)
            w[416 // 4] = cols                       # grant.nof_symb
            w[420 // 4 + 0] = {2: 1, 4: 2, 6: 3}[qm]  # grant.tb.mod
            w[420 // 4 + 1] = tbs                    # grant.tb.tbs
            w[420 // 4 + 2] = rv                     # grant.tb.rv
            w[420 // 4 + 3] = G                      # grant.tb.nof_bits
            cfg[504:512].view(np.uint64)[0] = C.addressof(sb)   # softbuffers.rx
            out = np.zeros(tbs // 8 + 8, np.uint8)
            gb = np.zeros(G + 8, np.int16)
            rc = L.srslte_ulsch_decode(q.ctypes.data, cfg.ctypes.data, chan, gb, None, out, None)
            want_rc, want_its = (int(x) for x in g[f"{c}_rv{rv}_res"])
            assert rc == want_rc, (c, rv)
            assert round(float(q[4:8].view(np.float32)[0]) * seg.C) == want_its, (c, rv)   # q->avg_iterations
            assert np.array_equal(out[: tbs // 8 + 3], g[f"{c}_rv{rv}_out"]), (c, rv)
            assert np.array_equal(gb[:G], llr), (c, rv)
            assert int(w[488 // 4]) == seg.C1 * seg.K1 + seg.C2 * seg.K2                  # cfg->K_segm
            ran += 1
    assert ran >= 4
    L.srslte_softbuffer_rx_free(C.byref(sb))
