"""Small end-to-end run of every kernel (crash / hang check; compute-sanitizer where the pool allows it): decode (all three regimes, CRC
and no CRC), rate de-matching, front end (plain, fused, UL de-interleaver), TB entry, TX mirror."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
pkg = ge.load_package(); vec = pkg.vectors
ctx = pkg.Context(0)
rng = np.random.default_rng(1)
for K, n in ((6144, 5), (5824, 3), (816, 9), (512, 11), (408, 3), (400, 67), (40, 3)):
    bits, llr = vec.make_blocks(n, K, 1.0, 100, seed=K)
    ctx.tdec_batch_host(llr, K, 3)
    ctx.tdec_batch_host(llr, K, 4, crc_mode=pkg.CRC_24B)
    llr2 = (llr.astype(np.int32) * 40).clip(-32768, 32767).astype(np.int16)   # exact / tracked tiers
    ctx.tdec_batch_host(llr2, K, 3)
# front end + fused rm + TB from symbols (PDSCH and PUSCH)
for qm, ul in ((2, 0), (4, 12), (6, 0), (8, 0)):
    nsym = 1440
    sym = ((rng.standard_normal(nsym) + 1j * rng.standard_normal(nsym)) * 0.8).astype(np.complex64)
    s_d = torch.from_numpy(sym.view(np.float32)).cuda()
    e_d = torch.zeros(qm * nsym, dtype=torch.int16, device="cuda")
    ctx.demod_descramble_dev([dict(qm=qm, nof_symbols=nsym, c_init=77, ul_nof_symb=ul)], s_d.data_ptr(), e_d.data_ptr())
    work = torch.zeros((2, 18624), dtype=torch.int16, device="cuda")
    E = qm * nsym // 2
    ctx.demod_rm_rx_batch_dev([dict(qm=qm, nof_symbols=nsym, c_init=77, ul_nof_symb=ul)],
                              [(1024, 0, 0, 0, E, 0), (1056, 2, 0, E, E, 18624)], s_d.data_ptr(), work.data_ptr())
# the vector paths at their edges: a codeword that ends inside a group, a descrambled length that ends inside one,
# several codewords packed on 128-bit boundaries at the very end of the buffers
for qm in (2, 4, 6, 8):
    cws, so, lo = [], 0, 0
    for nsym, cut in ((1441, 0), (7, 3), (1000, 50), (250, 0)):
        cws.append(dict(qm=qm, nof_symbols=nsym, c_init=99 + nsym, nof_bits=qm * nsym - cut, sym_offset=so, llr_offset=lo))
        so = (so + nsym + 1) & ~1
        lo = (lo + qm * nsym + 7) & ~7
    # the last codeword ends exactly at the end of both buffers
    s_d = torch.randn(2 * (cws[-1]["sym_offset"] + 250), dtype=torch.float32, device="cuda")
    e_d = torch.zeros(cws[-1]["llr_offset"] + qm * 250, dtype=torch.int16, device="cuda")
    ctx.demod_descramble_dev(cws, s_d.data_ptr(), e_d.data_ptr())
    work = torch.zeros((2, 18624), dtype=torch.int16, device="cuda")
    E = (qm * 250) // 2 // qm * qm
    ctx.demod_rm_rx_batch_dev(cws, [(512, 0, 3, 0, E, 0), (512, 1, 3, E, qm * 250 - E, 18624)], s_d.data_ptr(), work.data_ptr())
pool = ctx.harq_pool(2, 13)
sym = ((rng.standard_normal(15000) + 1j * rng.standard_normal(15000)) * 0.8).astype(np.complex64)
ctx.decode_tb_sym_batch(pool, [dict(tbs=75376, qm=6, rv=0, nof_e_bits=90000, softbuffer=0, c_init=5, symbols=sym)], 2)
pool.reset(1)
ctx.decode_tb_batch(pool, [dict(tbs=2216, qm=4, rv=0, e_bits=rng.integers(-300, 300, 4800).astype(np.int16), softbuffer=1)], 2)
pool.close()
b = torch.from_numpy(rng.integers(0, 2, 6144 + 40, dtype=np.uint8)).cuda()
e = torch.zeros(20000 + 200, dtype=torch.uint8, device="cuda")
ctx.tcod_rm_tx_batch_dev([(6144, 1, 20000, 0, 0), (40, 3, 200, 6144, 20000)], b.data_ptr(), e.data_ptr())
ctx.synchronize()
print("sanitize run done")
