"""Ad-hoc GPU parity probe (development aid): CUDA path vs the oracle port on noisy blocks."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import oracle_libs as ol
pkg = ge.load_package(); vec = pkg.vectors
ctx = pkg.Context(0)
Ks = [int(a) for a in sys.argv[1:]] or [6144, 5824, 1024, 816, 800, 512, 408, 400, 104, 40]
bad_total = 0
for K in Ks:
    for sigma, scale in [(vec.harness_sigma(1.5), 100), (vec.harness_sigma(4.0), 100), (0.9, 700)]:
        bits, llr = vec.make_blocks(6, K, sigma, scale, seed=K)
        for nit in [1, 2, 3, 4, 8]:
            t0 = time.time()
            got, n_iter, _ = ctx.tdec_batch_host(llr, K, nit)
            dt = time.time() - t0
            want = ol.port_run_all(llr, K, nit)
            nbad = int((got != want).any(axis=1).sum())
            bad_total += nbad
            if nbad:
                i = int(np.argmax((got != want).any(axis=1)))
                diff = np.unpackbits(got[i] ^ want[i])
                print(f"K={K} sigma={sigma:.3f} scale={scale} nit={nit}: {nbad}/6 blocks differ; block {i}: {int(diff.sum())} bit errors, first at {int(np.argmax(diff))}, ({dt*1e3:.1f} ms)")
        # working-layout input
        sb = vec.sb_layout_from_natural(llr, K)
        got, _, _ = ctx.tdec_batch_host(sb, K, 5, natural=False)
        want = ol.port_run_all(sb, K, 5, natural=False)
        nbad = int((got != want).any(axis=1).sum()); bad_total += nbad
        if nbad: print(f"K={K} working-layout input: {nbad}/6 differ")
    # CRC early termination
    bits, llr = vec.make_blocks(12, K, vec.harness_sigma(4.0), 100, seed=K + 1)
    if K > 40:
        got, n_iter, ok = ctx.tdec_batch_host(llr, K, 10, crc_mode=pkg.CRC_24B)
        P = ol.port()
        for i in range(12):
            by, so, _ = ol.port_trace(llr[i], K, 10)
            crcs = [P.port_crc_bytes(ol.CRC24B, by[it].copy(), K) for it in range(10)]
            stop = next((it + 1 for it in range(10) if crcs[it] == 0), 10)
            exp_ok = int(crcs[stop - 1] == 0)
            if n_iter[i] != stop or ok[i] != exp_ok or not np.array_equal(got[i], by[stop - 1]):
                bad_total += 1
                print(f"K={K} CRC mode block {i}: n_iter {n_iter[i]} vs {stop}, ok {ok[i]} vs {exp_ok}, bytes eq {np.array_equal(got[i], by[stop-1])}")
    print(f"K={K} done, cumulative mismatches {bad_total}", flush=True)
print("TOTAL MISMATCHES", bad_total)
print("fallbacks (warp x half-iteration) so far:", ctx.fallback_count)
