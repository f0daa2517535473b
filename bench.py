#!/usr/bin/env python
"""bench.py -- batched LTE turbo decode on B200: information Gbit/s (K=6144, nof_iterations=4).

One "step" = one pass of the hot path over one batch of synthetic code blocks:
natural-order int16 LLRs -> working layout -> 4 half iterations of int16 max-log-MAP -> hard decision
(srslte_tdec_run_all semantics, the reference's turbodecoder_test config scaled to a batch).

  value        device-resident throughput (inputs already in HBM), CUDA events, max over ranks
  e2e          same metric through the host-pointer C-ABI call (pinned host input, H2D + D2H inside)
  roofline     dominant kernel (16-window decoder) against the measured integer-pipe peak
               (+ roofline_hbm: its algorithmic bytes against the measured HBM copy bandwidth)
  cpu_baseline the reference's own AVX2 decoder (oracle/_ref) on the box's host cores, bounded sample

`--impl reference` times the reference's CPU implementation alone on all host cores.
Multi-GPU: one process per GPU under torchrun, code blocks sharded, no data-path collective
(torch.distributed is only used for the barrier and the max-over-ranks of the timing).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 6144
NOF_ITERATIONS = 4          # srsLTE HALF iterations
EBNO_HARNESS = 1.5          # turbodecoder_test "-e 1.5" (sigma = 1.457 on +-1; never converges, F7)
LLR_SCALE = 100.0
IN_LEN = 3 * K + 12
# DRAM bytes of the decode kernel per code block, from the committed ncu capture of this workload (profiles/r02*_summary.txt:
# dram__bytes_read.sum + dram__bytes_write.sum of one 65 536-block launch / 65 536); NOT measured by bench.py itself
NCU_DRAM_BYTES_PER_BLOCK = 249.2e3
NCU_DRAM_SOURCE = "profiles/r02k_summary.txt (ncu --set full: 12.94 GB read + 3.39 GB written per 65536-block launch)"
INT_PEAK_THREAD_INSTR_PER_CLK_SM = 64.0   # measured: profiles/r01_int_peak*.txt (VIADD.16x2 / VIMNMX.S16x2), ONE pipe
TWO_PIPE_THREAD_INSTR_PER_CLK_SM = 117.6  # measured: profiles/r01_pipe_mix.txt, "vaddmax + vadd" (both pipes busy)


def int_ops_per_block(k=K, w=16, nit=NOF_ITERATIONS):
    """int16 lane-ops per MAP call = 80*K + 1960*W."""
    return nit * (80 * k + 1960 * w)


def algo_bytes_per_block(k=K):
    """(3K+12)*2 B read + K/8 B + 2 B written."""
    return (3 * k + 12) * 2 + k // 8 + 2


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.stop_flag = False
        self.gpu = gpu_index
        self.t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.t:
            self.t.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_llr_device(n, device, seed):
    """Synthetic natural-order int16 LLRs [n, 3K+12] on the GPU: random payloads -> LTE turbo encoder ->
    BPSK +-1 -> AWGN (harness -e 1.5) -> (int16)(100*rx).  A pool of 512 distinct code words is tiled over
    the batch; the noise is independent per block."""
    import torch
    import __graft_entry__ as ge
    vec = ge.load_package().vectors
    pool = 512
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, (pool, K), dtype=np.uint8)
    coded = torch.from_numpy(vec.turbo_encode(bits)).to(device)          # [pool, 3K+12] 0/1
    sigma = vec.harness_sigma(EBNO_HARNESS)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n, IN_LEN), dtype=torch.int16, device=device)
    step = 4096
    for i in range(0, n, step):
        m = min(step, n - i)
        idx = (torch.arange(i, i + m, device=device) % pool)
        tx = coded[idx].to(torch.float32) * 2.0 - 1.0
        rx = tx + sigma * torch.randn((m, IN_LEN), device=device, generator=g)
        out[i:i + m] = torch.trunc(LLR_SCALE * rx).clamp_(-32768, 32767).to(torch.int16)
    return out


def host_cores():
    """(threads this process may run on, physical cores among them) -- the CPU arm uses one pthread per usable thread."""
    try:
        usable = sorted(os.sched_getaffinity(0))
    except Exception:
        usable = list(range(os.cpu_count() or 1))
    phys = set()
    try:
        cur = {}
        for line in open("/proc/cpuinfo"):
            if ":" in line:
                k, v = [x.strip() for x in line.split(":", 1)]
                cur[k] = v
            elif not line.strip():
                if cur and int(cur.get("processor", -1)) in usable:
                    phys.add((cur.get("physical id", "0"), cur.get("core id", cur.get("processor"))))
                cur = {}
        if cur and int(cur.get("processor", -1)) in usable:
            phys.add((cur.get("physical id", "0"), cur.get("core id", cur.get("processor"))))
    except Exception:
        pass
    return len(usable), (len(phys) or len(usable))


class CpuReference:
    """The reference's AVX2 decoder (oracle/_ref) on the host cores.  One srslte_tdec_t per pthread is created ONCE,
    outside every timed region (the reference's own turbodecoder_test initialises once and loops,
    lib/src/phy/fec/test/turbodecoder_test.c:190-266); only the srslte_tdec_run_all loops are timed."""

    def __init__(self, threads):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_libs as ol
        self.ol = ol
        self.threads = threads
        self.kind = "reference" if ol.ref() is not None else "port"
        self.pool = ol.RefPool(threads) if self.kind == "reference" else None

    def run(self, llr_np):
        """decode all blocks once; returns (wall seconds of the decode loops, summed per-thread loop seconds)"""
        if self.pool is not None:
            _, wall, busy = self.pool.run(llr_np, K, NOF_ITERATIONS, True)
            return wall, busy
        t0 = time.perf_counter()
        self.ol.port_run_all(llr_np, K, NOF_ITERATIONS, True)
        dt = time.perf_counter() - t0
        return dt, dt

    def close(self):
        if self.pool is not None:
            self.pool.close()


def cpu_reference_rate(cpu, llr_np, min_wall=1.5):
    """Gbit/s over repeated passes (about min_wall seconds of decode time), per-core microseconds per block."""
    cpu.run(llr_np[:max(cpu.threads * 4, 8)])   # warm the caches
    wall = busy = 0.0
    passes = 0
    while True:
        w, b = cpu.run(llr_np)
        wall += w
        busy += b
        passes += 1
        if wall >= min_wall:
            break
    blocks = passes * len(llr_np)
    return {"gbps": blocks * K / wall / 1e9, "wall_s": wall, "passes": passes,
            "us_per_block_per_core": 1e6 * busy / blocks}


def workload_config(n, world):
    """The `config` object of the bench line: the same for the B200 arm and for the reference arm."""
    return {
        "workload": f"batched srslte_tdec_run_all: {n} code blocks per GPU, K={K}, nof_iterations={NOF_ITERATIONS} "
                    f"(srsLTE half iterations), natural-order int16 LLR, AWGN harness -e {EBNO_HARNESS} "
                    f"(sigma 1.457), LLR scale {LLR_SCALE:g}, 16-window int16 max-log-MAP, no CRC",
        "blocks_per_gpu": n, "K": K, "nof_iterations": NOF_ITERATIONS,
        "l2": f"inputs {n * IN_LEN * 2 / 1e6:.0f} MB per step > 126 MB L2, no flush needed",
        "parallelism": f"code blocks sharded over {world} GPU(s), no collective",
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as ge
    vec = ge.load_package().vectors
    threads, phys = host_cores()
    cpu = CpuReference(threads)                  # decoder handles are created here, before any timing
    n = max(threads * 256, 2048) if cpu.kind == "reference" else 64   # distinct blocks of the sample
    # a step = the B200 arm's step (args.blocks code blocks) decoded as repeated passes over the sample, so that a step
    # lasts a few tenths of a second whatever --steps is (a single 20 ms pass measured cold threads and clock ramps)
    passes = max(1, min(args.blocks // n, 64)) if cpu.kind == "reference" else 1
    bits, llr = vec.make_blocks(n, K, vec.harness_sigma(EBNO_HARNESS), LLR_SCALE, seed=7, crc=False)

    def one_step():
        w = b = 0.0
        for _ in range(passes):
            wi, bi = cpu.run(llr)
            w += wi
            b += bi
        return w, b
    for _ in range(max(args.warmup, 1)):
        one_step()
    walls, busys = [], []
    for _ in range(args.steps):
        w, b = one_step()
        walls.append(w)
        busys.append(b)
    cpu.close()
    dt = float(np.mean(walls))
    value = n * passes * K / dt / 1e9
    us_core = 1e6 * float(np.sum(busys)) / (n * passes * len(busys))
    world = max(args.gpus, 1)
    line = {
        "impl": "reference", "metric": "turbo_decode_info_gbps_k6144_4it", "value": value, "unit": "Gbit/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        # the B200 arm's workload; every step decodes a bounded sample of it on the host cores (cpu_baseline.sample)
        "config": workload_config(args.blocks, world),
        "cpu_baseline": {"value": value, "unit": "Gbit/s", "cores": threads, "physical_cores": phys, "kind": cpu.kind,
                         "us_per_block_per_core": us_core, "per_core_gbps": K / us_core / 1e3,
                         "sample": f"{passes} passes over {n} blocks of the workload (K={K}, nof_iterations={NOF_ITERATIONS}) per step, "
                                   f"{threads} pthreads with one srslte_tdec_t each, created before the timed region; "
                                   f"only the srslte_tdec_run_all loops are timed; the reference's AVX2 16-window decoder",
                         "host": "ONE host (this box's CPU cores), whatever --gpus is: the reference has no multi-GPU or "
                                 "multi-host form, so the per-N ratios compare N GPUs with the same single host"},
        "e2e": {"value": value, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--blocks", type=int, default=65536, help="code blocks per GPU per step")
    ap.add_argument("--e2e-blocks", type=int, default=65536, help="code blocks per GPU for the host-pointer leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE.json configurations (key configs)")
    ap.add_argument("--quick-configs", action="store_true", help="smaller batches for the configs key (development)")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = None
    if world > 1:
        # one process per GPU: run on (and first-touch the pinned buffers from) the CPUs next to this rank's GPU
        try:
            import pynvml
            pynvml.nvmlInit()
            try:   # CUDA_VISIBLE_DEVICES may renumber the devices: go by UUID
                h = pynvml.nvmlDeviceGetHandleByUUID(f"GPU-{torch.cuda.get_device_properties(local).uuid}")
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(local)
            pynvml.nvmlDeviceSetCpuAffinity(h)
            affinity = f"{len(os.sched_getaffinity(0))} gpu-local cpus per rank (nvml)"
        except Exception as e:  # no NVML / cpuset restrictions: keep the inherited affinity
            affinity = f"inherited ({type(e).__name__})"
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        # only a barrier and one max-reduce of the timing go through it: gloo (the data path has no collective)
        dist_mod.init_process_group("gloo")
        dist = dist_mod

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.blocks
    ctx = pkg.Context(local)
    stream = torch.cuda.Stream(device=dev)   # an explicit stream: the kernels and the timing events share it
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    llr = make_llr_device(n, dev, seed=1234 + rank)
    out = torch.zeros((n, K // 8), dtype=torch.uint8, device=dev)
    nit = torch.zeros(n, dtype=torch.uint8, device=dev)
    crc = torch.zeros(n, dtype=torch.uint8, device=dev)

    def step():
        ctx.tdec_batch_dev(llr.data_ptr(), n, IN_LEN, K, NOF_ITERATIONS, out.data_ptr(), K // 8, nit.data_ptr(),
                           crc.data_ptr(), crc_mode=pkg.CRC_NONE, natural=True)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.enable_timing(True)
    launches0 = ctx.launch_count
    fallbacks0 = ctx.fallback_count
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    fallbacks = ctx.fallback_count - fallbacks0
    dec_ms, dec_n = ctx.kernel_time(0)
    lay_ms, lay_n = ctx.kernel_time(3)
    ctx.enable_timing(False)
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms_total], dtype=torch.float64)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * n * K / (ms_step * 1e-3) / 1e9

    # sanity: the decoder must have produced something that depends on the input
    chk = int(out[: min(n, 64)].to(torch.int32).sum().item())

    # ---- end to end through the host-pointer entry (pinned host memory, H2D + D2H inside) -------------
    ne = min(args.e2e_blocks, n)
    pin_in = pkg.PinnedArray((ne, IN_LEN), np.int16)
    pin_out = pkg.PinnedArray((ne, K // 8), np.uint8)
    pin_in.array[:] = llr[:ne].cpu().numpy()
    ctx.set_stream(0)  # the host entry uses the context's own streams
    for _ in range(2):
        ctx.tdec_batch_host(pin_in.array, K, NOF_ITERATIONS, out=pin_out.array)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        got, got_nit, _ = ctx.tdec_batch_host(pin_in.array, K, NOF_ITERATIONS, out=pin_out.array)
    e2e_dt = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_dt], dtype=torch.float64)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * ne * K / float(t.item()) / 1e9
    e2e_ranks = [e2e_dt * 1e3]
    if dist:
        allt = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allt, torch.tensor([e2e_dt * 1e3], dtype=torch.float64))
        e2e_ranks = [float(x.item()) for x in allt]
    same = bool(np.array_equal(pin_out.array[:64], out[:64].cpu().numpy()))
    # the box's host-to-device ceiling in the same run: every rank copies the same pinned input at the same time, no kernels
    barrier()
    h2d_gbs = ctx.h2d_probe(pin_in.array.ctypes.data, ne * IN_LEN * 2, reps=3)
    h2d_ranks = [h2d_gbs]
    if dist:
        allg = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allg, torch.tensor([h2d_gbs], dtype=torch.float64))
        h2d_ranks = [float(x.item()) for x in allg]

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured" if "hbm_gbs" in peaks else "fallback"
    sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    sm_max = (clocks or {}).get("sm_max_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    dec_ms_per_launch = dec_ms / max(dec_n, 1)
    ops_launch = n * int_ops_per_block()
    achieved_tops = ops_launch / (dec_ms_per_launch * 1e-3) / 1e12
    peak_tops = INT_PEAK_THREAD_INSTR_PER_CLK_SM * 2 * sms * sm_max * 1e6 / 1e12      # int16 lane-ops/s at max clock
    two_pipe_tops = TWO_PIPE_THREAD_INSTR_PER_CLK_SM * 2 * sms * sm_max * 1e6 / 1e12
    bytes_launch = n * algo_bytes_per_block()
    hbm_achieved = bytes_launch / (dec_ms_per_launch * 1e-3) / 1e9

    line = {
        "metric": "turbo_decode_info_gbps_k6144_4it", "value": value, "unit": "Gbit/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": workload_config(n, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Gbit/s", "h2d_bytes_per_step": ne * IN_LEN * 2,
                "d2h_bytes_per_step": ne * (K // 8 + 2), "blocks_per_gpu": ne, "ms_per_step": e2e_dt * 1e3,
                "matches_device_path": same, "host_affinity": affinity or "inherited",
                "ms_per_step_per_rank": e2e_ranks,
                # H2D copies alone, all ranks at once (srslte_b200_h2d_probe): what the PCIe side of the box gives
                "h2d_ceiling_gbs_per_rank": h2d_ranks,
                "h2d_gbs_per_rank_in_e2e": [ne * IN_LEN * 2 / (ms * 1e-3) / 1e9 for ms in e2e_ranks],
                "e2e_frac_of_h2d_ceiling": (ne * IN_LEN * 2 / (max(e2e_ranks) * 1e-3) / 1e9) / max(min(h2d_ranks), 1e-9)},
        "gpu_launches": launches,
        "roofline": {
            "bound": "int_alu", "kernel": "tdec_win_kernel<16, false, true> (launches without CRC)", "achieved": achieved_tops, "peak": peak_tops,
            "unit": "Tops/s (int16 lane-ops)", "frac": achieved_tops / peak_tops,
            "ops_per_block": int_ops_per_block(), "ms_per_launch": dec_ms_per_launch,
            "peak_source": f"measured {INT_PEAK_THREAD_INSTR_PER_CLK_SM:g} packed-int16x2 thread-instr/clk/SM "
                           f"(tools/int_peak.cu, profiles/r01_int_peak*.txt) x 2 lanes x {sms} SMs x {sm_max:g} MHz",
            "sm_mhz_during_run": sm_mhz,
            # NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed
            # ncu --set full capture, per block, scaled to this launch
            "traffic": NCU_DRAM_BYTES_PER_BLOCK * n, "traffic_from_profile": NCU_DRAM_SOURCE,
            "two_pipe_peak": two_pipe_tops, "frac_two_pipe": achieved_tops / two_pipe_tops,
            "note": "frac: against ONE integer pipe (64 packed thread-instr/clk/SM, what every single packed op "
                    "reaches); frac_two_pipe: against the measured rate of the kernel's own instruction mix with both "
                    "pipes busy (117.6 thread-instr/clk/SM for VIADDMNMX + VIADD.16x2, profiles/r01_pipe_mix.txt); a "
                    "fused add-max counts as two algorithmic ops",
        },
        "roofline_hbm": {
            "bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
            "bytes_per_block": algo_bytes_per_block(), "peak_source": f"MEASURED_PEAKS.json ({hbm_src})",
            "traffic": NCU_DRAM_BYTES_PER_BLOCK * n, "traffic_from_profile": NCU_DRAM_SOURCE,
            "traffic_gbs": NCU_DRAM_BYTES_PER_BLOCK * n / (dec_ms_per_launch * 1e-3) / 1e9,
            "traffic_frac_of_peak": NCU_DRAM_BYTES_PER_BLOCK * n / (dec_ms_per_launch * 1e-3) / 1e9 / hbm_peak,
        },
        "kernel_share": {"decode_ms_per_step": dec_ms / args.steps, "layout_ms_per_step": lay_ms / args.steps,
                         "decode_launches": dec_n, "layout_launches": lay_n},
        "exact_fallbacks": {"count": fallbacks, "of": args.steps * ((n + 3) // 4) * NOF_ITERATIONS,
                            "unit": "(warp, half iteration) pairs re-run with exact saturating arithmetic"},
        "checksum": chk,
    }

    if not args.no_cpu and world == 1:
        threads, phys = host_cores()
        cpu = CpuReference(threads)               # handles created before the timed passes
        sample = min(ne, max(1024, threads * 256)) if cpu.kind == "reference" else min(ne, 32)
        llr_np = pin_in.array[:sample].copy()
        r = cpu_reference_rate(cpu, llr_np, min_wall=3.0)
        cpu.close()
        line["cpu_baseline"] = {"value": r["gbps"], "unit": "Gbit/s", "cores": threads, "physical_cores": phys,
                                "kind": cpu.kind, "us_per_block_per_core": r["us_per_block_per_core"],
                                "per_core_gbps": K / r["us_per_block_per_core"] / 1e3,
                                "sample": f"{r['passes']} passes over {sample} of the same K={K} blocks, "
                                          f"nof_iterations={NOF_ITERATIONS}, {threads} pthreads with one srslte_tdec_t "
                                          f"each (created before the timed region), {r['wall_s']:.2f} s of decode loops"}
    if not args.no_configs and world == 1:
        try:
            import bench_configs
            ctx.set_stream(stream.cuda_stream)
            line["configs"] = bench_configs.run_configs(pkg, ctx, torch, dev, stream, quick=args.quick_configs)
        except Exception as e:  # the headline line must not be lost to a side measurement
            line["configs"] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
