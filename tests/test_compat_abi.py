"""The struct layouts restated in include/srslte_b200_compat.h equal the reference's (CPU only).

A tiny C program is compiled against OUR header and prints sizeof/offsetof; the numbers are compared with
the ones the compiled reference reports through oracle/_ref (refh_sizeof_*), or -- when oracle/_ref is not
available -- with the values recorded from the reference's AVX2 build (SURVEY.md section 7 "ABI")."""
import os
import subprocess

import pytest

import oracle_libs as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROBE = r"""
#include <stdio.h>
#include <stddef.h>
#include "srslte_b200_compat.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(srslte_tdec_t), offsetof(srslte_tdec_t, n_iter),
         sizeof(srslte_softbuffer_rx_t), offsetof(srslte_tdec_t, interleaver), sizeof(srslte_tc_interl_t));
  return 0;
}
"""


def test_struct_layouts_match_reference(tmp_path):
    src = tmp_path / "probe.c"
    src.write_text(PROBE)
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    R = ol.ref()
    if R is not None:
        want_tdec, want_niter, want_sb = R.refh_sizeof_tdec(), R.refh_offsetof_tdec_n_iter(), R.refh_sizeof_softbuffer_rx()
    else:
        want_tdec, want_niter, want_sb = 18264, 18256, 40
    assert got[0] == want_tdec == 18264
    assert got[1] == want_niter
    assert got[2] == want_sb
    assert got[4] == 24


PROBE_DLSCH = r"""
#include <stdio.h>
#include <stddef.h>
#include "srslte_b200_compat.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(srslte_ra_tb_t), sizeof(srslte_pdsch_grant_t),
         sizeof(srslte_pdsch_cfg_t), offsetof(srslte_pdsch_cfg_t, grant) + offsetof(srslte_pdsch_grant_t, tb),
         offsetof(srslte_pdsch_cfg_t, grant) + offsetof(srslte_pdsch_grant_t, nof_tb), offsetof(srslte_ra_tb_t, nof_bits),
         offsetof(srslte_ra_tb_t, rv), offsetof(srslte_pdsch_cfg_t, softbuffers), offsetof(srslte_sch_head_t, max_iterations),
         offsetof(srslte_sch_head_t, avg_iterations), offsetof(srslte_sch_head_t, llr_is_8bit));
  return 0;
}
"""


def test_dlsch_layouts_match_reference(tmp_path):
    """srslte_pdsch_cfg_t / srslte_ra_tb_t / the head of srslte_sch_t as srslte_dlsch_decode2 reads them"""
    import ctypes as C
    src = tmp_path / "probe2.c"
    src.write_text(PROBE_DLSCH)
    exe = tmp_path / "probe2"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    R = ol.ref()
    if R is not None and hasattr(R, "refh_dlsch_layout"):
        arr = (C.c_size_t * 11)()
        R.refh_dlsch_layout(arr)
        want = list(arr)
    else:   # recorded from the reference's build (oracle/ref_harness.c:refh_dlsch_layout)
        want = [28, 316, 368, 244, 308, 12, 8, 344, 0, 4, 8]
    assert got == want


PROBE_ULSCH = r"""
#include <stdio.h>
#include <stddef.h>
#include "srslte_b200_compat.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu ", sizeof(srslte_sch_ul_t) - sizeof(uint64_t), offsetof(srslte_sch_ul_t, ack_ri_bits),
         offsetof(srslte_sch_ul_t, encoder), offsetof(srslte_sch_ul_t, decoder), offsetof(srslte_sch_ul_t, crc_tb),
         offsetof(srslte_sch_ul_t, uci_cqi));
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu ", sizeof(srslte_pusch_cfg_t), offsetof(srslte_pusch_cfg_t, uci_cfg),
         offsetof(srslte_pusch_cfg_t, uci_cfg) + offsetof(srslte_uci_cfg_t, cqi), offsetof(srslte_pusch_cfg_t, uci_offset),
         offsetof(srslte_pusch_cfg_t, grant), offsetof(srslte_pusch_cfg_t, grant) + offsetof(srslte_pusch_grant_t, nof_symb),
         offsetof(srslte_pusch_cfg_t, grant) + offsetof(srslte_pusch_grant_t, tb), offsetof(srslte_pusch_cfg_t, K_segm),
         offsetof(srslte_pusch_cfg_t, softbuffers));
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(srslte_uci_cfg_ack_t), offsetof(srslte_uci_cfg_ack_t, nof_acks),
         sizeof(srslte_cqi_cfg_t), offsetof(srslte_cqi_cfg_t, ri_len), sizeof(srslte_uci_value_t),
         offsetof(srslte_uci_value_t, cqi) + offsetof(srslte_cqi_value_t, data_crc), offsetof(srslte_uci_value_t, ack),
         offsetof(srslte_uci_value_t, ri), sizeof(srslte_uci_bit_t));
  return 0;
}
"""


def test_ulsch_layouts_match_reference(tmp_path):
    """srslte_pusch_cfg_t / srslte_uci_value_t / srslte_sch_t as srslte_ulsch_decode reads them.  srslte_sch_ul_t ends with
    the first word of uci_cqi: its size without that word is the reference's offset of uci_cqi."""
    import ctypes as C
    src = tmp_path / "probe3.c"
    src.write_text(PROBE_ULSCH)
    exe = tmp_path / "probe3"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    R = ol.ref()
    if R is not None and hasattr(R, "refh_ulsch_layout"):
        arr = (C.c_size_t * 24)()
        R.refh_ulsch_layout(arr)
        want = list(arr)
    else:   # recorded from the reference's build (oracle/ref_harness.c:refh_ulsch_layout)
        want = [490792, 56, 460856, 460872, 479136, 483312, 520, 4, 344, 372, 384, 416, 420, 488, 504, 68, 4, 24, 20, 40, 24,
                28, 39, 8]
    assert want[5] == 483312 and want[6] == 520
    want[0] = want[5]          # our struct stops where uci_cqi starts
    assert got == want


def test_compat_symbols_are_exported(pkg):
    L = pkg.lib()
    for name in ("srslte_tdec_init", "srslte_tdec_init_manual", "srslte_tdec_free", "srslte_tdec_force_not_sb",
                 "srslte_tdec_new_cb", "srslte_tdec_get_nof_iterations", "srslte_tdec_autoimp_get_subblocks", "srslte_ulsch_decode",
                 "srslte_tdec_autoimp_get_subblocks_8bit", "srslte_tdec_iteration", "srslte_tdec_run_all",
                 "srslte_tdec_iteration_8bit", "srslte_tdec_run_all_8bit", "srslte_rm_turbo_gentables",
                 "srslte_rm_turbo_free_tables", "srslte_rm_turbo_rx_lut", "srslte_rm_turbo_rx_lut_",
                 "srslte_rm_turbo_rx_lut_8bit", "srslte_b200_sch_decode_tb", "srslte_dlsch_decode",
                 "srslte_dlsch_decode2", "srslte_softbuffer_rx_init", "srslte_softbuffer_rx_free", "srslte_softbuffer_rx_reset",
                 "srslte_softbuffer_rx_reset_tbs", "srslte_softbuffer_rx_reset_cb"):
        assert hasattr(L, name), name
    # host-only entry points work without a GPU
    assert [L.srslte_tdec_autoimp_get_subblocks(k) for k in (40, 408, 816, 6144)] == [0, 8, 16, 16]
    assert [L.srslte_tdec_autoimp_get_subblocks_8bit(k) for k in (40, 408, 816, 2112, 6144)] == [0, 8, 16, 32, 32]


def test_ulsch_decode_without_the_reference_uci_decoders_fails_loudly(pkg, capfd):
    """Loaded alone (ctypes, no libsrslte_phy in the process) the weak references to the reference's UCI decoders are
    NULL: a PUSCH with multiplexed control information must be refused with a message before anything is decoded, and
    so must the 8-bit mode and a transport block size without a segmentation (no GPU is touched on these paths)."""
    import ctypes as C
    L = pkg.lib()
    L.srslte_ulsch_decode.argtypes = [C.c_void_p] * 7
    L.srslte_ulsch_decode.restype = C.c_int
    q = (C.c_uint8 * (490792 + 64))()          # srslte_sch_t, zeroed: max_iterations 0, llr_is_8bit false
    cfg = (C.c_uint8 * 520)()                  # srslte_pusch_cfg_t
    def put32(buf, off, val):
        C.cast(C.byref(buf, off), C.POINTER(C.c_uint32))[0] = val
    llr = (C.c_int16 * 4096)()
    g = (C.c_int16 * 4096)()
    seq = (C.c_uint8 * 4096)()
    data = (C.c_uint8 * 512)()
    uci = (C.c_uint8 * 40)()
    put32(cfg, 420 + 0, 1)                     # grant.tb.mod = QPSK
    put32(cfg, 420 + 4, 504)                   # grant.tb.tbs
    put32(cfg, 420 + 12, 1728)                 # grant.tb.nof_bits
    put32(cfg, 416, 12)                        # grant.nof_symb
    put32(cfg, 4 + 4, 1)                       # uci_cfg.ack[0].nof_acks = 1
    rc = L.srslte_ulsch_decode(q, cfg, llr, g, seq, data, uci)
    assert rc == -1
    assert "UCI decoders" in capfd.readouterr().err
    put32(cfg, 4 + 4, 0)
    q[8] = 1                                   # llr_is_8bit
    assert L.srslte_ulsch_decode(q, cfg, llr, g, seq, data, uci) == -1
    assert "8-bit" in capfd.readouterr().err
    q[8] = 0
    assert L.srslte_ulsch_decode(None, cfg, llr, g, seq, data, uci) == -2
    put32(cfg, 416, 0)                         # grant.nof_symb = 0: no interleaver matrix
    assert L.srslte_ulsch_decode(q, cfg, llr, g, seq, data, uci) == -2
    put32(cfg, 416, 12)
    put32(cfg, 420 + 0, 7)                     # not a modulation
    assert L.srslte_ulsch_decode(q, cfg, llr, g, seq, data, uci) == -2


def test_manual_decoder_types_that_are_not_provided_are_refused(pkg, capfd):
    """srslte_tdec_init_manual: the non-windowed SSE decoder, NEON and the 8-bit decoders are refused before any device
    is touched; without a GPU the accepted types fail loudly too (no CPU fallback)."""
    import ctypes as C
    import torch
    L = pkg.lib()
    L.srslte_tdec_init_manual.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
    for typ in (2, 4, 6, 7, 8, 42):            # SSE, NEON_WINDOW, SSE8_WINDOW, AVX8_WINDOW, NOF_IMP, junk
        h = C.create_string_buffer(18264)
        assert L.srslte_tdec_init_manual(h, 6144, typ) == -1
        assert "not supported" in capfd.readouterr().err
    if not torch.cuda.is_available():
        for typ in (0, 1, 3, 5):               # AUTO, GENERIC, SSE_WINDOW, AVX_WINDOW
            h = C.create_string_buffer(18264)
            assert L.srslte_tdec_init_manual(h, 6144, typ) == -1
