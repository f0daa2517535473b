// lte_tables.h -- host-side LTE constants for the B200 turbo-decode path.
//
// 36.212 Table 5.1.3-3 (code-block sizes and QPP coefficients), decoder-regime selection and
// rate-dematching index tables.  Reference behaviour being matched (paths under /root/reference):
//   lib/src/phy/fec/cbsegm.c:30-135        code-block sizes, segmentation
//   lib/src/phy/fec/tc_interl_lte.c:38-62  f1/f2
//   lib/src/phy/fec/turbodecoder.c:394-406 window count per K (AVX2 AUTO build)
//   lib/src/phy/fec/rm_turbo.c:160-260     receive index tables
#pragma once
#include <cstdint>
#include <vector>

namespace b200 {

constexpr int      kNofCbSizes  = 188;
constexpr uint32_t kMaxK        = 6144;
constexpr uint32_t kSoftbufLen  = 18600;  // int16 per code block (softbuffer.h:50)
constexpr uint32_t kCrc24A      = 0x1864CFB;
constexpr uint32_t kCrc24B      = 0x1800063;

struct QppEntry {
  uint16_t K, f1, f2;
};
extern const QppEntry kQpp[kNofCbSizes];

int  cb_index_ceil(uint32_t long_cb);  // first table index with K >= long_cb, -1 if none
int  cb_index_exact(uint32_t K);       // -1 if K is not a valid code-block size
int  cb_size(uint32_t idx);            // -1 if out of range
int  nof_windows(uint32_t K);          // 16, 8, or 0 (generic scalar decoder)

// length of one code block in the decoder's working layout:
//   W > 0: 3*(K+32)+12 (three sub-block-interleaved streams with 32-sample pads, then tail)
//   W = 0: 3*K+12 (natural 3i+j order)
uint32_t working_len(uint32_t K);

struct CbSegm {
  uint32_t F, C, K1, K2, K1_idx, K2_idx, C1, C2, tbs;
};
int cbsegm(CbSegm* s, uint32_t tbs);

// table[i] = working-layout index that accumulates rate-matched sample i, i < 3K+12.
// sb_layout=false gives natural 3i+j indices for every K.
void rm_rx_table(uint32_t K, uint32_t rv, bool sb_layout, std::vector<uint16_t>& table);

// 36.211 7.2 scrambling sequence c(n) = x1(n + 1600) xor x2(n + 1600) for n < len (sequence.c:46-75):
// x1_packed bit (n & 31) of word n / 32 = x1(n + 1600); x2(n + 1600) = parity(x2_mask[n] & c_init), because x2 is a
// linear function of its 31 seed bits.
void gold_tables(uint32_t len, std::vector<uint32_t>& x1_packed, std::vector<uint32_t>& x2_mask);

void crc24_table(uint32_t poly, uint32_t table[256]);
// What bit p of a K-bit block contributes to its CRC24 register (init 0): x^(K + 23 - p) mod P.  The CRC is linear,
// so the register is the xor of the contributions of the set bits IN ANY ORDER: the window decoders accumulate it
// while they produce the hard decisions, without assembling the bit string.
// out[((which * 2 + dir) * L + k) * W + d], which: 0 = CRC24A, 1 = CRC24B; dir 0: trellis position d*L + k of DEC1
// (natural order), dir 1: of DEC2 (bit pi(d*L + k)).  Window decoders only (W = 16 or 8).
void crc_pos_tables(uint32_t K, std::vector<uint32_t>& out);
uint32_t crc24_bytes(uint32_t poly, const uint8_t* data, uint32_t nbytes);

}  // namespace b200
