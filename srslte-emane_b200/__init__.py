"""srslte-emane_b200: B200-native LTE turbo-decode receive tail (host-side Python binding).

The product is the C-ABI shared library ``libsrslte_b200.so`` (include/srslte_b200.h).  This module is
only a ctypes mirror of that ABI for the tests and bench.py; it contains no decoding logic and never
touches oracle/.  It fails loudly when the library is missing or no CUDA device is usable -- there is
no CPU fallback.

Because the directory name contains a hyphen, import it through ``__graft_entry__.load_package()``.
"""
import ctypes as C
import os

import numpy as np

from . import sharding  # noqa: F401
from . import vectors  # noqa: F401  (synthetic vector generation; no decoder inside)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsrslte_b200.so")

INPUT_NATURAL, INPUT_WORKING = 0, 1
CRC_NONE, CRC_24B, CRC_24A = 0, 1, 2
SUCCESS, ERROR, ERROR_INVALID_INPUTS = 0, -1, -2


class TdecBatch(C.Structure):
    _fields_ = [("n_cb", C.c_uint32), ("long_cb", C.POINTER(C.c_uint32)), ("uniform_long_cb", C.c_uint32),
                ("input_format", C.c_uint32), ("in_stride", C.c_uint32), ("out_stride", C.c_uint32),
                ("nof_iterations", C.c_uint32), ("crc_mode", C.c_uint32)]


class RmBlock(C.Structure):
    _fields_ = [("long_cb", C.c_uint32), ("rv", C.c_uint32), ("e_offset", C.c_uint32), ("e_len", C.c_uint32),
                ("work_offset", C.c_uint32)]


class UlUci(C.Structure):
    """srslte_b200_ul_uci_t: UCI multiplexed into a PUSCH codeword (data-path view)."""
    _fields_ = [("q_prime_ack", C.c_uint32), ("q_prime_ri", C.c_uint32), ("q_prime_cqi", C.c_uint32),
                ("ri_len", C.c_uint32)]


def _uci(d):
    u = d.get("uci") or {}
    return UlUci(u.get("q_prime_ack", 0), u.get("q_prime_ri", 0), u.get("q_prime_cqi", 0), u.get("ri_len", 0))


class Codeword(C.Structure):
    _fields_ = [("qm", C.c_uint32), ("nof_symbols", C.c_uint32), ("c_init", C.c_uint32), ("nof_bits", C.c_uint32),
                ("sym_offset", C.c_uint64), ("llr_offset", C.c_uint64), ("ul_nof_symb", C.c_uint32),
                ("reserved", C.c_uint32), ("uci", UlUci)]


class RmSymBlock(C.Structure):
    _fields_ = [("long_cb", C.c_uint32), ("rv", C.c_uint32), ("codeword", C.c_uint32), ("e_offset", C.c_uint32),
                ("e_len", C.c_uint32), ("work_offset", C.c_uint32)]


class TbDesc(C.Structure):
    _fields_ = [("tbs", C.c_uint32), ("qm", C.c_uint32), ("rv", C.c_uint32), ("nof_e_bits", C.c_uint32),
                ("softbuffer", C.c_uint32), ("e_bits", C.c_void_p), ("data", C.c_void_p), ("ret", C.c_int32),
                ("avg_iterations", C.c_float)]


class TxBlock(C.Structure):
    _fields_ = [("long_cb", C.c_uint32), ("rv", C.c_uint32), ("e_len", C.c_uint32), ("reserved", C.c_uint32),
                ("bits_offset", C.c_uint64), ("e_offset", C.c_uint64)]


class TbSymDesc(C.Structure):
    _fields_ = [("tbs", C.c_uint32), ("qm", C.c_uint32), ("rv", C.c_uint32), ("nof_e_bits", C.c_uint32),
                ("softbuffer", C.c_uint32), ("nof_symbols", C.c_uint32), ("c_init", C.c_uint32),
                ("ul_nof_symb", C.c_uint32), ("uci", UlUci), ("symbols", C.c_void_p), ("data", C.c_void_p),
                ("ret", C.c_int32), ("avg_iterations", C.c_float)]


EXPORTS = [
    "srslte_b200_ctx_create", "srslte_b200_ctx_destroy", "srslte_b200_ctx_set_stream",
    "srslte_b200_ctx_synchronize", "srslte_b200_last_error", "srslte_b200_launch_count",
    "srslte_b200_ctx_enable_timing", "srslte_b200_ctx_kernel_time", "srslte_b200_ctx_set_exact", "srslte_b200_ctx_set_variant_bits",
    "srslte_b200_ctx_fallback_count", "srslte_b200_ctx_tier_counts", "srslte_b200_host_alloc", "srslte_b200_host_free", "srslte_b200_cb_index", "srslte_b200_cb_size",
    "srslte_b200_nof_windows", "srslte_b200_working_len", "srslte_b200_rm_rx_table",
    "srslte_b200_tdec_batch_dev", "srslte_b200_tdec_batch_host", "srslte_b200_rm_rx_batch_dev",
    "srslte_b200_demod_descramble_dev", "srslte_b200_demod_rm_rx_batch_dev", "srslte_b200_tcod_rm_tx_batch_dev",
    "srslte_b200_harq_pool_create", "srslte_b200_harq_pool_destroy", "srslte_b200_harq_reset", "srslte_b200_harq_reset_many",
    "srslte_b200_harq_cb_crc", "srslte_b200_decode_tb_batch", "srslte_b200_decode_tb_sym_batch",
    "srslte_b200_uci_q_prime_ri_ack", "srslte_b200_uci_q_prime_cqi",
    "srslte_b200_group_create", "srslte_b200_group_destroy", "srslte_b200_group_size", "srslte_b200_group_ctx",
    "srslte_b200_group_tdec_batch_host", "srslte_b200_group_h2d_probe", "srslte_b200_h2d_probe",
    "srslte_b200_group_set_weights", "srslte_b200_group_calibrate",
]

_lib = None


def lib():
    """Load libsrslte_b200.so (raises if it has not been built: run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                           " -- this package has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
    L.srslte_b200_ctx_create.argtypes = [C.POINTER(vp), i32]
    L.srslte_b200_ctx_destroy.argtypes = [vp]
    L.srslte_b200_ctx_destroy.restype = None
    L.srslte_b200_ctx_set_stream.argtypes = [vp, vp]
    L.srslte_b200_ctx_synchronize.argtypes = [vp]
    L.srslte_b200_last_error.argtypes = [vp]
    L.srslte_b200_last_error.restype = C.c_char_p
    L.srslte_b200_launch_count.argtypes = [vp]
    L.srslte_b200_launch_count.restype = C.c_uint64
    L.srslte_b200_ctx_enable_timing.argtypes = [vp, i32]
    L.srslte_b200_ctx_kernel_time.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(u32)]
    L.srslte_b200_ctx_set_exact.argtypes = [vp, i32]
    L.srslte_b200_ctx_set_variant_bits.argtypes = [vp, u32]
    L.srslte_b200_ctx_fallback_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.srslte_b200_ctx_tier_counts.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.srslte_b200_host_alloc.argtypes = [C.c_size_t]
    L.srslte_b200_host_alloc.restype = vp
    L.srslte_b200_host_free.argtypes = [vp]
    L.srslte_b200_host_free.restype = None
    L.srslte_b200_cb_index.argtypes = [u32]
    L.srslte_b200_cb_size.argtypes = [u32]
    L.srslte_b200_nof_windows.argtypes = [u32]
    L.srslte_b200_nof_windows.restype = u32
    L.srslte_b200_working_len.argtypes = [u32]
    L.srslte_b200_working_len.restype = u32
    L.srslte_b200_rm_rx_table.argtypes = [u32, u32, i32, vp]
    L.srslte_b200_tdec_batch_dev.argtypes = [vp, C.POINTER(TdecBatch), vp, vp, vp, vp]
    L.srslte_b200_tdec_batch_host.argtypes = [vp, C.POINTER(TdecBatch), vp, vp, vp, vp]
    L.srslte_b200_rm_rx_batch_dev.argtypes = [vp, C.POINTER(RmBlock), u32, vp, vp]
    L.srslte_b200_demod_descramble_dev.argtypes = [vp, C.POINTER(Codeword), u32, vp, vp]
    L.srslte_b200_demod_rm_rx_batch_dev.argtypes = [vp, C.POINTER(Codeword), u32, C.POINTER(RmSymBlock), u32, vp, vp]
    L.srslte_b200_tcod_rm_tx_batch_dev.argtypes = [vp, C.POINTER(TxBlock), u32, vp, vp]
    L.srslte_b200_harq_pool_create.argtypes = [vp, u32, u32, C.POINTER(vp)]
    L.srslte_b200_harq_pool_destroy.argtypes = [vp, vp]
    L.srslte_b200_harq_pool_destroy.restype = None
    L.srslte_b200_harq_reset.argtypes = [vp, vp, u32]
    L.srslte_b200_harq_reset_many.argtypes = [vp, vp, vp, u32]
    L.srslte_b200_harq_cb_crc.argtypes = [vp, u32, vp, u32]
    L.srslte_b200_decode_tb_batch.argtypes = [vp, vp, C.POINTER(TbDesc), u32, u32]
    L.srslte_b200_decode_tb_sym_batch.argtypes = [vp, vp, C.POINTER(TbSymDesc), u32, u32]
    L.srslte_b200_uci_q_prime_ri_ack.argtypes = [u32, u32, u32, u32, C.c_float]
    L.srslte_b200_uci_q_prime_ri_ack.restype = u32
    L.srslte_b200_uci_q_prime_cqi.argtypes = [u32, u32, u32, u32, C.c_float, u32]
    L.srslte_b200_uci_q_prime_cqi.restype = u32
    L.srslte_b200_group_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), u32]
    L.srslte_b200_group_destroy.argtypes = [vp]
    L.srslte_b200_group_destroy.restype = None
    L.srslte_b200_group_size.argtypes = [vp]
    L.srslte_b200_group_size.restype = u32
    L.srslte_b200_group_ctx.argtypes = [vp, u32]
    L.srslte_b200_group_ctx.restype = vp
    L.srslte_b200_group_tdec_batch_host.argtypes = [vp, C.POINTER(TdecBatch), vp, vp, vp, vp]
    L.srslte_b200_group_h2d_probe.argtypes = [vp, vp, C.c_size_t, u32, C.POINTER(C.c_double)]
    L.srslte_b200_h2d_probe.argtypes = [vp, vp, C.c_size_t, u32, C.POINTER(C.c_double)]
    L.srslte_b200_group_set_weights.argtypes = [vp, C.POINTER(C.c_double)]
    L.srslte_b200_group_calibrate.argtypes = [vp, C.POINTER(C.c_double)]
    _lib = L
    return L


class B200Error(RuntimeError):
    pass


class Group:
    """Several GPUs of one box from ONE process (srslte_b200_group_t): a batch is cut into contiguous shards, one per
    device, each run from its own host thread; no inter-device traffic."""

    def __init__(self, devices):
        self._L = lib()
        self._h = C.c_void_p()
        devs = list(range(devices)) if isinstance(devices, int) else list(devices)
        arr = (C.c_int * len(devs))(*devs)
        rc = self._L.srslte_b200_group_create(C.byref(self._h), arr, len(devs))
        if rc:
            raise B200Error(f"srslte_b200_group_create({devs}) failed ({rc}); there is no CPU fallback")

    def close(self):
        if self._h:
            self._L.srslte_b200_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self._L.srslte_b200_group_size(self._h))

    def tdec_batch_host(self, llr, K, nof_iterations, crc_mode=CRC_NONE, natural=True, out=None, nit=None, ok=None):
        assert llr.dtype == np.int16 and llr.ndim == 2 and llr.flags.c_contiguous
        n, in_stride = llr.shape
        kmax = int(K) if np.isscalar(K) else int(np.max(K)) if n else 0
        out_stride = kmax // 8
        if out is None:
            out = np.zeros((n, out_stride), np.uint8)
        nit = np.zeros(n, np.uint8) if nit is None else nit
        ok = np.zeros(n, np.uint8) if ok is None else ok
        b, keep = Context._batch(n, K, natural, in_stride, out_stride, nof_iterations, crc_mode)
        rc = self._L.srslte_b200_group_tdec_batch_host(self._h, C.byref(b), llr.ctypes.data, out.ctypes.data,
                                                       nit.ctypes.data, ok.ctypes.data)
        if rc:
            raise B200Error(f"srslte_b200_group_tdec_batch_host -> {rc}")
        return out, nit, ok

    def set_weights(self, weights=None):
        """share of a batch per device (None: equal shares)"""
        arr = None if weights is None else (C.c_double * len(self))(*weights)
        if self._L.srslte_b200_group_set_weights(self._h, arr):
            raise B200Error("srslte_b200_group_set_weights: bad weights")

    def calibrate(self):
        """measure the devices' concurrent host-to-device rates and use them as shares; returns GB/s per device"""
        g = (C.c_double * len(self))()
        rc = self._L.srslte_b200_group_calibrate(self._h, g)
        if rc:
            raise B200Error(f"srslte_b200_group_calibrate -> {rc}")
        return [float(x) for x in g]

    def h2d_probe(self, host_ptr, bytes_per_device, reps=4):
        """GB/s per device while ALL devices of the group copy from pinned host memory at the same time."""
        g = (C.c_double * len(self))()
        rc = self._L.srslte_b200_group_h2d_probe(self._h, C.c_void_p(host_ptr), bytes_per_device, reps, g)
        if rc:
            raise B200Error(f"srslte_b200_group_h2d_probe -> {rc}")
        return [float(x) for x in g]


def rm_rx_table(K, rv, sb_layout=True):
    t = np.zeros(3 * K + 12, np.uint16)
    rc = lib().srslte_b200_rm_rx_table(K, rv, int(sb_layout), t.ctypes.data)
    if rc:
        raise B200Error(f"srslte_b200_rm_rx_table({K},{rv}) -> {rc}")
    return t


def working_len(K):
    return int(lib().srslte_b200_working_len(K))


class PinnedArray:
    """numpy view over cudaHostAlloc memory (inputs of the *_host entry that need no staging)."""

    def __init__(self, shape, dtype):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._p = lib().srslte_b200_host_alloc(max(nbytes, 1))
        if not self._p:
            raise B200Error("cudaHostAlloc failed")
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._p:
            self.array = None
            lib().srslte_b200_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One decoder context on one GPU (srslte_b200_ctx_t)."""

    def __init__(self, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        rc = self._L.srslte_b200_ctx_create(C.byref(self._h), device)
        if rc:
            raise B200Error(f"srslte_b200_ctx_create(device={device}) failed ({rc}): no usable CUDA device; "
                            "there is no CPU fallback")

    def close(self):
        if self._h:
            self._L.srslte_b200_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc:
            raise B200Error(f"{what} -> {rc}: {self._L.srslte_b200_last_error(self._h).decode()}")

    def set_stream(self, cuda_stream_ptr):
        self._check(self._L.srslte_b200_ctx_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), "set_stream")

    def synchronize(self):
        self._check(self._L.srslte_b200_ctx_synchronize(self._h), "synchronize")

    def enable_timing(self, on=True):
        self._check(self._L.srslte_b200_ctx_enable_timing(self._h, int(on)), "enable_timing")

    def kernel_time(self, kind):
        """(total ms, launches) of kernel kind 0..4 since timing was enabled (synchronizes)."""
        ms, n = C.c_double(), C.c_uint32()
        self._check(self._L.srslte_b200_ctx_kernel_time(self._h, kind, C.byref(ms), C.byref(n)), "kernel_time")
        return ms.value, n.value

    def set_exact(self, on=True):
        """force the exact saturating variant of the window decoders (the fast variant is the default)."""
        self._check(self._L.srslte_b200_ctx_set_exact(self._h, int(on)), "set_exact")

    def set_variant_bits(self, bits):
        """tests / measurements: restrict the (bit-identical) decoder variants, see include/srslte_b200.h"""
        self._check(self._L.srslte_b200_ctx_set_variant_bits(self._h, int(bits)), "set_variant_bits")

    @property
    def tier_counts(self):
        """(warp, half iteration) pairs so far in the pure / static / tracked / exact variant"""
        v = (C.c_uint64 * 4)()
        self._check(self._L.srslte_b200_ctx_tier_counts(self._h, v), "tier_counts")
        return [int(x) for x in v]

    @property
    def fallback_count(self):
        v = C.c_uint64()
        self._check(self._L.srslte_b200_ctx_fallback_count(self._h, C.byref(v)), "fallback_count")
        return int(v.value)

    @property
    def launch_count(self):
        return int(self._L.srslte_b200_launch_count(self._h))

    def h2d_probe(self, host_ptr, nbytes, reps=4):
        """GB/s of `reps` host-to-device copies of nbytes from (pinned) host_ptr on this context's copy stream."""
        g = C.c_double()
        self._check(self._L.srslte_b200_h2d_probe(self._h, C.c_void_p(host_ptr), nbytes, reps, C.byref(g)), "h2d_probe")
        return g.value

    @staticmethod
    def _batch(n, K, natural, in_stride, out_stride, nof_iterations, crc_mode):
        b = TdecBatch()
        b.n_cb = n
        keep = None
        if np.isscalar(K):
            b.long_cb = None
            b.uniform_long_cb = int(K)
        else:
            keep = np.ascontiguousarray(K, dtype=np.uint32)
            assert keep.shape == (n,)
            b.long_cb = keep.ctypes.data_as(C.POINTER(C.c_uint32))
            b.uniform_long_cb = 0
        b.input_format = INPUT_NATURAL if natural else INPUT_WORKING
        b.in_stride = in_stride
        b.out_stride = out_stride
        b.nof_iterations = nof_iterations
        b.crc_mode = crc_mode
        return b, keep

    # ---- host-pointer entry: the call a user of the reference's API makes -------------------------
    def tdec_batch_host(self, llr, K, nof_iterations, crc_mode=CRC_NONE, natural=True, out=None):
        """llr: int16 [n, in_stride] (numpy, host).  Returns (bytes [n, out_stride], n_iter [n], crc_ok [n])."""
        assert llr.dtype == np.int16 and llr.ndim == 2 and llr.flags.c_contiguous
        n, in_stride = llr.shape
        kmax = int(K) if np.isscalar(K) else int(np.max(K)) if n else 0
        out_stride = kmax // 8
        if out is None:
            out = np.zeros((n, out_stride), np.uint8)
        nit = np.zeros(n, np.uint8)
        ok = np.zeros(n, np.uint8)
        b, keep = self._batch(n, K, natural, in_stride, out_stride, nof_iterations, crc_mode)
        rc = self._L.srslte_b200_tdec_batch_host(self._h, C.byref(b), llr.ctypes.data, out.ctypes.data,
                                                 nit.ctypes.data, ok.ctypes.data)
        self._check(rc, "srslte_b200_tdec_batch_host")
        return out, nit, ok

    # ---- device-pointer entry: raw CUDA pointers (e.g. torch tensors' data_ptr()) -----------------
    def tdec_batch_dev(self, llr_ptr, n, in_stride, K, nof_iterations, out_ptr, out_stride, nit_ptr=0, crc_ptr=0,
                       crc_mode=CRC_NONE, natural=True):
        b, keep = self._batch(n, K, natural, in_stride, out_stride, nof_iterations, crc_mode)
        rc = self._L.srslte_b200_tdec_batch_dev(self._h, C.byref(b), C.c_void_p(llr_ptr), C.c_void_p(out_ptr),
                                                C.c_void_p(nit_ptr), C.c_void_p(crc_ptr))
        self._check(rc, "srslte_b200_tdec_batch_dev")

    # ---- transport blocks (sch.c decode_tb semantics, device-resident HARQ soft buffers) ----------
    def harq_pool(self, n_softbuffers, max_cb):
        return HarqPool(self, n_softbuffers, max_cb)

    def decode_tb_batch(self, pool, tbs, max_iterations):
        """tbs: list of dicts {tbs, qm, rv, e_bits (int16 ndarray), softbuffer}.
        Returns list of (ret, data bytes [tbs/8 + 6], avg_iterations)."""
        n = len(tbs)
        arr = (TbDesc * n)()
        outs, keep = [], []
        for i, t in enumerate(tbs):
            e = np.ascontiguousarray(t["e_bits"], dtype=np.int16)
            o = np.zeros(t["tbs"] // 8 + 8, np.uint8)
            keep.append(e)
            outs.append(o)
            arr[i] = TbDesc(t["tbs"], t["qm"], t["rv"], e.size, t["softbuffer"], e.ctypes.data, o.ctypes.data, 0, 0.0)
        rc = self._L.srslte_b200_decode_tb_batch(self._h, pool._p, arr, n, max_iterations)
        self._check(rc, "srslte_b200_decode_tb_batch")
        return [(int(arr[i].ret), outs[i], float(arr[i].avg_iterations)) for i in range(n)]

    def decode_tb_sym_batch(self, pool, tbs, max_iterations):
        """tbs: list of dicts(tbs, qm, rv, nof_e_bits, softbuffer, c_init, symbols=complex64 array[, ul_nof_symb,
        uci=dict(q_prime_ack, q_prime_ri, q_prime_cqi, ri_len)]).
        Returns [(ret, data bytes, avg_iterations)] like decode_tb_batch."""
        n = len(tbs)
        arr = (TbSymDesc * n)()
        keep, outs = [], []
        for i, d in enumerate(tbs):
            sym = np.ascontiguousarray(d["symbols"], dtype=np.complex64)
            out = np.zeros(d["tbs"] // 8 + 8, np.uint8)
            keep.append(sym)
            outs.append(out)
            arr[i] = TbSymDesc(d["tbs"], d["qm"], d["rv"], d["nof_e_bits"], d["softbuffer"], sym.shape[0], d["c_init"],
                               d.get("ul_nof_symb", 0), _uci(d), sym.ctypes.data, out.ctypes.data, 0, 0.0)
        rc = self._L.srslte_b200_decode_tb_sym_batch(self._h, pool._p, arr, n, max_iterations)
        self._check(rc, "srslte_b200_decode_tb_sym_batch")
        return [(int(arr[i].ret), outs[i], float(arr[i].avg_iterations)) for i in range(n)]

    def tcod_rm_tx_batch_dev(self, blocks, bits_ptr, e_ptr):
        """TX mirror: blocks = list of (K, rv, e_len, bits_offset, e_offset); bits / e: device, one bit per byte."""
        arr = (TxBlock * len(blocks))()
        for i, (K, rv, el, bo, eo) in enumerate(blocks):
            arr[i] = TxBlock(K, rv, el, 0, bo, eo)
        rc = self._L.srslte_b200_tcod_rm_tx_batch_dev(self._h, arr, len(blocks), C.c_void_p(bits_ptr), C.c_void_p(e_ptr))
        self._check(rc, "srslte_b200_tcod_rm_tx_batch_dev")

    def rm_rx_batch_dev(self, blocks, e_ptr, work_ptr):
        """blocks: list of (K, rv, e_offset, e_len, work_offset)."""
        arr = (RmBlock * len(blocks))()
        for i, (K, rv, eo, el, wo) in enumerate(blocks):
            arr[i] = RmBlock(K, rv, eo, el, wo)
        rc = self._L.srslte_b200_rm_rx_batch_dev(self._h, arr, len(blocks), C.c_void_p(e_ptr), C.c_void_p(work_ptr))
        self._check(rc, "srslte_b200_rm_rx_batch_dev")


    # ---- front end: soft demodulation + descrambling (device pointers) ---------------------------
    @staticmethod
    def _codewords(cws):
        arr = (Codeword * len(cws))()
        for i, c in enumerate(cws):
            arr[i] = Codeword(c["qm"], c["nof_symbols"], c["c_init"], c.get("nof_bits", c["qm"] * c["nof_symbols"]),
                              c.get("sym_offset", 0), c.get("llr_offset", 0), c.get("ul_nof_symb", 0), 0, _uci(c))
        return arr

    def demod_descramble_dev(self, cws, symbols_ptr, e_ptr):
        """cws: list of dicts(qm, nof_symbols, c_init[, nof_bits, sym_offset, llr_offset])."""
        arr = self._codewords(cws)
        rc = self._L.srslte_b200_demod_descramble_dev(self._h, arr, len(cws), C.c_void_p(symbols_ptr), C.c_void_p(e_ptr))
        self._check(rc, "srslte_b200_demod_descramble_dev")

    def demod_rm_rx_batch_dev(self, cws, blocks, symbols_ptr, work_ptr):
        """blocks: list of (K, rv, codeword, e_offset, e_len, work_offset)."""
        arr = self._codewords(cws)
        bl = (RmSymBlock * len(blocks))()
        for i, b in enumerate(blocks):
            bl[i] = RmSymBlock(*b)
        rc = self._L.srslte_b200_demod_rm_rx_batch_dev(self._h, arr, len(cws), bl, len(blocks), C.c_void_p(symbols_ptr),
                                                       C.c_void_p(work_ptr))
        self._check(rc, "srslte_b200_demod_rm_rx_batch_dev")


class HarqPool:
    """Device-resident HARQ soft buffers (srslte_b200_harq_pool_t)."""

    def __init__(self, ctx, n_softbuffers, max_cb):
        self._ctx = ctx
        self._p = C.c_void_p()
        self.max_cb = max_cb
        rc = ctx._L.srslte_b200_harq_pool_create(ctx._h, n_softbuffers, max_cb, C.byref(self._p))
        ctx._check(rc, "srslte_b200_harq_pool_create")

    def reset(self, softbuffer):
        self._ctx._check(self._ctx._L.srslte_b200_harq_reset(self._ctx._h, self._p, softbuffer), "harq_reset")

    def cb_crc(self, softbuffer, n):
        out = np.zeros(n, np.uint8)
        rc = self._ctx._L.srslte_b200_harq_cb_crc(self._p, softbuffer, out.ctypes.data, n)
        self._ctx._check(rc, "harq_cb_crc")
        return out

    def close(self):
        if self._p:
            self._ctx._L.srslte_b200_harq_pool_destroy(self._ctx._h, self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CompatTdec:
    """The reference's own entry points (srslte_tdec_init / force_not_sb / run_all / free, turbodecoder.h:97-135) on a
    raw 18 264-byte srslte_tdec_t, as turbodecoder_test.c drives them: natural-order LLRs, one block per call."""

    def __init__(self, max_long_cb=6144):
        self.L = lib()
        i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
        u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        self.L.srslte_tdec_init.argtypes = [C.c_void_p, C.c_uint32]
        self.L.srslte_tdec_free.argtypes = [C.c_void_p]
        self.L.srslte_tdec_free.restype = None
        self.L.srslte_tdec_force_not_sb.argtypes = [C.c_void_p]
        self.L.srslte_tdec_force_not_sb.restype = None
        self.L.srslte_tdec_run_all.argtypes = [C.c_void_p, i16p, u8p, C.c_uint32, C.c_uint32]
        self.h = C.create_string_buffer(18264)
        if self.L.srslte_tdec_init(self.h, max_long_cb) != 0:
            raise RuntimeError("srslte_tdec_init failed (no CUDA device? this library has no CPU path)")
        self.L.srslte_tdec_force_not_sb(self.h)

    def run_all(self, llr1, nof_iterations, K):
        out = np.zeros(K // 8, np.uint8)
        rc = self.L.srslte_tdec_run_all(self.h, np.ascontiguousarray(llr1, np.int16).copy(), out, nof_iterations, K)
        if rc != 0:
            raise RuntimeError(f"srslte_tdec_run_all returned {rc}")
        return out

    def close(self):
        if self.h is not None:
            self.L.srslte_tdec_free(self.h)
            self.h = None
