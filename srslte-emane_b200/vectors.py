"""Synthetic test-vector generation for the turbo-decode path (numpy, CPU).

This is the TX mirror needed to make inputs: CRC attach, LTE turbo encoder (36.212 5.1.3.2),
circular-buffer rate matching and a BPSK/AWGN int16 LLR model with the conventions of the
reference's own test harness (lib/src/phy/fec/test/turbodecoder_test.c:207-252: the "-e" value is
used as Es/N0 with sigma = sqrt(1/10^(e/10)) applied as a standard deviation, LLR = 100 * rx).
It is not part of the decode path and contains no decoder.
"""
import numpy as np

CRC24A = 0x1864CFB
CRC24B = 0x1800063

ALL_K = (list(range(40, 513, 8)) + list(range(528, 1025, 16)) + list(range(1056, 2049, 32))
         + list(range(2112, 6145, 64)))

# 36.212 Table 5.1.3-3 (f1, f2) in the order of ALL_K
_F1 = [3, 7, 19, 7, 7, 11, 5, 11, 7, 41, 103, 15, 9, 17, 9, 21, 101, 21, 57, 23, 13, 27, 11, 27, 85, 29, 33, 15, 17,
       33, 103, 19, 19, 37, 19, 21, 21, 115, 193, 21, 133, 81, 45, 23, 243, 151, 155, 25, 51, 47, 91, 29, 29, 247,
       29, 89, 91, 157, 55, 31, 17, 35, 227, 65, 19, 37, 41, 39, 185, 43, 21, 155, 79, 139, 23, 217, 25, 17, 127, 25,
       239, 17, 137, 215, 29, 15, 147, 29, 59, 65, 55, 31, 17, 171, 67, 35, 19, 39, 19, 199, 21, 211, 21, 43, 149,
       45, 49, 71, 13, 17, 25, 183, 55, 127, 27, 29, 29, 57, 45, 31, 59, 185, 113, 31, 17, 171, 209, 253, 367, 265,
       181, 39, 27, 127, 143, 43, 29, 45, 157, 47, 13, 111, 443, 51, 51, 451, 257, 57, 313, 271, 179, 331, 363, 375,
       127, 31, 33, 43, 33, 477, 35, 233, 357, 337, 37, 71, 71, 37, 39, 127, 39, 39, 31, 113, 41, 251, 43, 21, 43,
       45, 45, 161, 89, 323, 47, 23, 47, 263]
_F2 = [10, 12, 42, 16, 18, 20, 22, 24, 26, 84, 90, 32, 34, 108, 38, 120, 84, 44, 46, 48, 50, 52, 36, 56, 58, 60, 62,
       32, 198, 68, 210, 36, 74, 76, 78, 120, 82, 84, 86, 44, 90, 46, 94, 48, 98, 40, 102, 52, 106, 72, 110, 168,
       114, 58, 118, 180, 122, 62, 84, 64, 66, 68, 420, 96, 74, 76, 234, 80, 82, 252, 86, 44, 120, 92, 94, 48, 98,
       80, 102, 52, 106, 48, 110, 112, 114, 58, 118, 60, 122, 124, 84, 64, 66, 204, 140, 72, 74, 76, 78, 240, 82,
       252, 86, 88, 60, 92, 846, 48, 28, 80, 102, 104, 954, 96, 110, 112, 114, 116, 354, 120, 610, 124, 420, 64, 66,
       136, 420, 216, 444, 456, 468, 80, 164, 504, 172, 88, 300, 92, 188, 96, 28, 240, 204, 104, 212, 192, 220, 336,
       228, 232, 236, 120, 244, 248, 168, 64, 130, 264, 134, 408, 138, 280, 142, 480, 146, 444, 120, 152, 462, 234,
       158, 80, 96, 902, 166, 336, 170, 86, 174, 176, 178, 120, 182, 184, 186, 94, 190, 480]
QPP = {k: (a, b) for k, a, b in zip(ALL_K, _F1, _F2)}

_COLPERM = np.array([0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                     1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31])


def nof_subblocks(K):
    """Window count of the reference's AUTO/AVX2 16-bit decoder for this K (16, 8, or 0 = generic)."""
    if K % 16 == 0 and K > 800:
        return 16
    if K % 8 == 0 and K > 400:
        return 8
    return 0


def qpp_perm(K):
    f1, f2 = QPP[K]
    i = np.arange(K, dtype=np.int64)
    return ((f1 * i + f2 * i * i) % K).astype(np.int64)


def crc24_bits(poly, bits):
    """bits: [..., n] of 0/1 -> CRC register [...] (MSB first, init 0)."""
    bits = np.asarray(bits, dtype=np.uint32)
    crc = np.zeros(bits.shape[:-1], dtype=np.uint32)
    p = np.uint32(poly & 0xFFFFFF)
    for i in range(bits.shape[-1]):
        top = ((crc >> 23) & 1) ^ bits[..., i]
        crc = ((crc << 1) & np.uint32(0xFFFFFF)) ^ (top * p)
    return crc


def attach_crc(poly, payload_bits):
    """append the 24 CRC bits (MSB first) to payload_bits [..., n] -> [..., n+24]."""
    crc = crc24_bits(poly, payload_bits)
    tail = ((crc[..., None] >> np.arange(23, -1, -1, dtype=np.uint32)) & 1).astype(np.uint8)
    return np.concatenate([np.asarray(payload_bits, np.uint8), tail], axis=-1)


def turbo_encode(bits):
    """bits [n, K] (0/1) -> coded [n, 3K+12] in the 3i+j order of srslte_tcod_encode, tail last."""
    bits = np.asarray(bits, dtype=np.uint8)
    n, K = bits.shape
    perm = qpp_perm(K)
    out = np.zeros((n, 3 * K + 12), np.uint8)
    out[:, 0:3 * K:3] = bits

    def rsc(seq, col):
        r0 = np.zeros(n, np.uint8); r1 = r0.copy(); r2 = r0.copy()
        par = np.zeros((n, K), np.uint8)
        for i in range(K):
            fb = seq[:, i] ^ r2 ^ r1
            par[:, i] = r2 ^ r0 ^ fb
            r2, r1, r0 = r1, r0, fb
        out[:, col:3 * K:3] = par
        tail = np.zeros((n, 6), np.uint8)
        for j in range(3):
            b = r2 ^ r1
            tail[:, 2 * j] = b
            fb = b ^ r2 ^ r1          # == 0: the register is flushed
            tail[:, 2 * j + 1] = r2 ^ r0 ^ fb
            r2, r1, r0 = r1, r0, fb
        return tail

    out[:, 3 * K:3 * K + 6] = rsc(bits, 1)
    out[:, 3 * K + 6:3 * K + 12] = rsc(bits[:, perm], 2)
    return out


def rm_select_table(K, rv):
    """natural coded index (into the 3K+12 vector of turbo_encode) read by rate-matched position i,
    for one wrap of the circular buffer (length 3K+12)."""
    D = K + 4
    R = (D - 1) // 32 + 1
    KP = 32 * R
    ND = KP - D
    jp = np.arange(3 * KP)
    nat = np.full(3 * KP, -1, np.int64)
    a = jp < KP
    d = (jp[a] % R) * 32 + _COLPERM[jp[a] // R] - ND
    nat[a] = np.where(d >= 0, 3 * d, -1)
    t = jp - KP
    b = (~a) & (t % 2 == 0)
    q = t[b] // 2
    d = (q % R) * 32 + _COLPERM[q // R] - ND
    nat[b] = np.where(d >= 0, 3 * d + 1, -1)
    c = (~a) & (t % 2 == 1)
    q = (t[c] - 1) // 2
    d = (_COLPERM[q // R] + 32 * (q % R) + 1) % KP - ND
    nat[c] = np.where(d >= 0, 3 * d + 2, -1)
    k0 = R * (2 * int(np.ceil(np.float32(3 * KP) / np.float32(8 * R))) * rv + 2)
    order = np.roll(nat, -k0)
    return order[order >= 0]


def rate_match(coded, E, rv):
    """coded [n, 3K+12] -> e [n, E] (36.212 5.1.4.1 bit selection, no soft-buffer limit)."""
    n, N = coded.shape
    K = (N - 12) // 3
    tab = rm_select_table(K, rv)
    idx = tab[np.arange(E) % N]
    return coded[:, idx]


def sb_layout_from_natural(nat, K):
    """natural [n, 3K+12] int16 -> the sub-block soft-buffer layout the decoder of this K expects
    (reference: rm_turbo.c:246-260); identity for generic K."""
    W = nof_subblocks(K)
    if W == 0:
        return np.ascontiguousarray(nat)
    n = nat.shape[0]
    L = K // W
    out = np.zeros((n, 3 * (K + 32) + 12), nat.dtype)
    pos = np.arange(K)
    st = (pos % L) * W + pos // L
    for j in range(3):
        out[:, j * (K + 32) + st] = nat[:, 3 * pos + j]
    out[:, 3 * (K + 32):] = nat[:, 3 * K:]
    return out


def harness_sigma(ebno_db):
    """sigma used by turbodecoder_test for its -e argument (turbodecoder_test.c:207-220): the harness turns
    "Eb/N0" into Es/N0 with the code rate 1/3 and then uses sqrt(1 / EsN0) as the STANDARD DEVIATION of the
    noise added to the +-1 symbols: -e 1.5 -> 1.457, -e 4.0 -> 1.092."""
    esno_db = ebno_db + 10.0 * np.log10(1.0 / 3.0)
    return float(np.sqrt(1.0 / 10.0 ** (esno_db / 10.0)))


def awgn_llr(coded_bits, sigma, scale=100.0, rng=None):
    """BPSK (+1 for bit 1) + sigma * N(0,1), quantised as (int16)(scale * rx) (truncation)."""
    rng = rng or np.random.default_rng(0)
    tx = 2.0 * coded_bits.astype(np.float32) - 1.0
    rx = tx + np.float32(sigma) * rng.standard_normal(tx.shape, dtype=np.float32)
    v = np.trunc(np.float32(scale) * rx)
    return np.clip(v, -32768, 32767).astype(np.int16)


def make_blocks(n, K, sigma, scale=100.0, seed=0, crc=True):
    """n random code blocks of size K (payload + CRC24B when crc) -> (bits [n,K], llr [n,3K+12])."""
    rng = np.random.default_rng(seed)
    if crc and K > 24:
        payload = rng.integers(0, 2, (n, K - 24), dtype=np.uint8)
        bits = attach_crc(CRC24B, payload)
    else:
        bits = rng.integers(0, 2, (n, K), dtype=np.uint8)
    coded = turbo_encode(bits)
    return bits, awgn_llr(coded, sigma, scale, rng)


def pack_bits(bits):
    return np.packbits(np.asarray(bits, np.uint8), axis=-1)
