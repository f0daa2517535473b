"""'Links unchanged' (SURVEY.md 8c level 2): the reference's own, unmodified sources linked against
libsrslte_b200.so give the reference's results.

oracle/Makefile `relink` builds (in the dev container, where /root/reference exists; the binaries travel to the GPU box
inside oracle/_ref/):
  turbodecoder_test_b200   /root/reference/lib/src/phy/fec/test/turbodecoder_test.c as it is, the reference's turbo decoder
                           sources left out of the link -> srslte_tdec_* come from this library
  dlsch_harness_ref/_b200  srslte_sch_init -> srslte_dlsch_encode2 -> srslte_dlsch_decode2 with HARQ retransmissions through the
                           reference's sch.c (compiled unchanged; for _b200 its two DL decode entry points are renamed out
                           of the way by a compile definition) -> srslte_dlsch_decode2 and srslte_tdec_* from this library
  ulsch_harness_ref/_b200  srslte_ulsch_encode -> scrambling -> srslte_ulsch_decode with multiplexed ACK / RI / CQI and HARQ
                           retransmissions, the way pusch.c drives it -> srslte_ulsch_decode from this library (which calls
                           the reference's UCI decoders for the control values and decodes the transport block on the device)
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _run(exe, *args):
    p = os.path.join(REFDIR, exe)
    if not os.path.exists(p):
        pytest.skip(f"{exe} was not built (oracle/Makefile relink needs the reference tree)")
    r = subprocess.run([p] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (exe, r.returncode, r.stderr[-2000:])
    return r.stdout


def _undefined(exe):
    out = subprocess.run(["nm", "-u", os.path.join(REFDIR, exe)], capture_output=True, text=True).stdout
    return {l.split()[-1].split("@")[0] for l in out.splitlines() if l.strip()}


def test_relink_binaries_take_the_hot_path_from_the_library():
    """not a GPU test: the two b200 binaries leave exactly the hot-path entry points to libsrslte_b200.so"""
    for exe in ("turbodecoder_test_b200", "dlsch_harness_b200"):
        if not os.path.exists(os.path.join(REFDIR, exe)):
            pytest.skip("relink binaries not built")
    u = _undefined("turbodecoder_test_b200")
    assert {"srslte_tdec_init", "srslte_tdec_run_all", "srslte_tdec_free"} <= u
    u = _undefined("dlsch_harness_b200")
    assert {"srslte_dlsch_decode2", "srslte_tdec_init", "srslte_tdec_free"} <= u
    assert "srslte_dlsch_encode2" not in u and "srslte_sch_init" not in u   # those are the reference's own objects
    if os.path.exists(os.path.join(REFDIR, "ulsch_harness_b200")):
        u = _undefined("ulsch_harness_b200")
        assert {"srslte_ulsch_decode", "srslte_tdec_init", "srslte_softbuffer_rx_init"} <= u
        assert "srslte_ulsch_encode" not in u and "srslte_uci_decode_ack_ri" not in u   # the reference's own objects


def _tdec_test_numbers(text):
    """(Eb/No, frame, BER) triples and the error totals of turbodecoder_test's output (timings differ, those do not)"""
    ber = re.findall(r"Eb/No:\s*([-\d.]+)\s+(\d+)/\d+\s+BER:\s*([-\d.e+]+)", text)
    errs = re.findall(r"(\d+) Errors", text)
    return ber, errs


@pytest.mark.gpu
@pytest.mark.parametrize("args", [("-n", 20, "-s", 1, "-e", "1.5", "-l", 6144, "-i", 4),      # BASELINE.json config 1
                                  ("-n", 30, "-s", 7, "-l", 1008),                               # the test's default sweep
                                  ("-n", 20, "-s", 3, "-e", "0.5", "-l", 504),
                                  ("-n", 20, "-s", 5, "-e", "1.0", "-l", 40),
                                  ("-n", 5, "-k"),                                              # the test's known code word
                                  ("-n", 10, "-s", 2, "-e", "1.5", "-l", 6144, "-i", 4, "-d", 5),  # pinned: 16-window AVX2
                                  ("-n", 10, "-s", 4, "-e", "1.0", "-l", 504, "-d", 3),          # pinned: 8-window SSE
                                  ("-n", 10, "-s", 6, "-e", "1.0", "-l", 40, "-d", 1)])         # pinned: generic
def test_reference_turbodecoder_test_relinked(args):
    ref = _tdec_test_numbers(_run("turbodecoder_test", *args))
    got = _tdec_test_numbers(_run("turbodecoder_test_b200", *args))
    assert len(ref[0]) >= 5
    assert got == ref


@pytest.mark.gpu
@pytest.mark.parametrize("args", [(12, 0.45, 10), (12, 0.30, 4), (8, 0.70, 8)])
def test_reference_sch_api_relinked(args):
    ref = _run("dlsch_harness_ref", *args)
    got = _run("dlsch_harness_b200", *args)
    assert "digest" in ref and ref.count("\n") > args[0]
    assert got == ref


@pytest.mark.gpu
@pytest.mark.parametrize("args", [(14, 0.45, 10), (14, 0.30, 4), (28, 0.60, 8)])
def test_reference_ulsch_api_relinked(args):
    """PUSCH transport blocks with multiplexed HARQ-ACK / RI / CQI: return codes, iterations, TB bytes, K_segm, the decoded
    control values, the de-interleaved LLRs (g_bits) and the modified q_bits of every transmission are identical"""
    ref = _run("ulsch_harness_ref", *args)
    got = _run("ulsch_harness_b200", *args)
    assert "digest" in ref and ref.count("\n") > args[0]
    assert got == ref


def test_ulsch_harness_drives_the_reference_correctly():
    """not a GPU test: the all-reference build of oracle/ulsch_harness.c decodes what it encoded -- every transport block
    ends with ret 0 and the transmitted bytes, the configured HARQ-ACK bits and the 1-bit RI come back, the CQI CRC holds --
    so the comparison of the two builds above compares meaningful PUSCH decodes, not two identical failures"""
    out = _run("ulsch_harness_ref", 14, 0.45, 10)
    rows = [l for l in out.splitlines() if l.startswith("tb ")]
    last = {}
    for l in rows:
        last[int(l.split()[1])] = l
    assert len(last) == 14 and len(rows) > 14          # some blocks needed a retransmission
    for t, l in last.items():
        m = re.search(r"ack (\d) ri (\d) cqi (\d): ret\s+(-?\d+) noi [\d.]+ match (\d) .*\| ack (\d)(\d)/(\d)(\d) ri (\d+)/(\d+) cqi \d+ crc (\d)", l)
        assert m, l
        nof_ack, ri_len, cqi_mode, ret, match = (int(m.group(i)) for i in range(1, 6))
        assert ret == 0 and match == 1, l
        rx_ack, tx_ack = (int(m.group(6)), int(m.group(7))), (int(m.group(8)), int(m.group(9)))
        assert rx_ack[:nof_ack] == tx_ack[:nof_ack], l
        if ri_len == 1:
            assert int(m.group(10)) == int(m.group(11)), l
        if cqi_mode == 2:                              # the long CQI carries a CRC
            assert int(m.group(12)) == 1, l
