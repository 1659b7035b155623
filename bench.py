#!/usr/bin/env python3
"""bench.py -- headline benchmark of the MPC per-block compression path on B200.

Metric (BASELINE.json): GB/s of memory blocks compressed, device-timed, whole job over N GPUs.
A step = one pass of the hot path over one batch of synthetic input: every rank compresses its
contiguous shard (1 GiB = 8 388 608 blocks of 128 B) of an N GiB synthetic dump that is already
resident in HBM.  The job's only collective -- one NCCL all-reduce of the statistics vector (N > 1) --
runs once, after the last step, inside the timed region.  Workload at N = 1 is
BASELINE.json configs[1]: "1 GB synthetic fp32 array dump (smooth values, delta-friendly)", compressed
with configs/F4.json.

  python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--config F4 --kind smooth_f32]
  torchrun ... bench.py --gpus N ...        (one rank per GPU; rank 0 prints ONE JSON line)

--impl reference times the reference's own CPU implementation (oracle/_ref, the unmodified sources
compiled by oracle/build_ref.sh; one single-threaded reference object per host core on disjoint slices)
on a bounded sample of the same workload.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BLOCK = 128
GIB = 1 << 30
METRIC = "GB/s of memory blocks compressed (device-timed), bit-exact ratio"


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            smax = mx
            if t0 - 0.05 <= t <= t1 + 0.05:
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: take every sample seen
            for t, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                except Exception:
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference classes on the host cores
# ------------------------------------------------------------------------------------------------------------
def _ref_worker(args):
    cfg_path, kind, seed, first, n, total = args
    from oracle.bridge import RefCompressor
    from tools.gen_dump import synth
    blocks = synth(kind, seed, first, n, total)
    ref = RefCompressor("VPC", cfg_path)
    t0 = time.perf_counter()
    ref.compress(blocks, want_sels=False)
    dt = time.perf_counter() - t0
    orig, comp, _ = ref.totals()
    return dt, orig, comp


def _port_worker(args):
    cfg_path, kind, seed, first, n, total = args
    from oracle.bridge import OracleMPC
    from tools.gen_dump import synth
    blocks = synth(kind, seed, first, n, total)
    o = OracleMPC(cfg_path)
    t0 = time.perf_counter()
    r = o.run(blocks, threads=1)
    dt = time.perf_counter() - t0
    return dt, r.OriginalSize, r.CompressedSize


def cpu_reference_pass(cfg_path, kind, seed, total_blocks, sample_blocks, pool, cores):
    """One bounded pass: `sample_blocks` blocks of the workload split over `cores` reference objects."""
    from oracle.bridge import have_ref
    worker, kindname = (_ref_worker, "reference") if have_ref() else (_port_worker, "port")
    per = max(1, sample_blocks // cores)
    jobs = [(cfg_path, kind, seed, i * per, per, total_blocks) for i in range(cores)]
    res = pool.map(worker, jobs)
    wall = max(r[0] for r in res)  # the pass ends when the slowest core finishes; input generation is not timed
    nbytes = per * cores * BLOCK
    orig = sum(r[1] for r in res)
    comp = sum(r[2] for r in res)
    return {"wall_s": wall, "bytes": nbytes, "kind": kindname, "max_worker_s": max(r[0] for r in res),
            "ratio": orig / comp if comp else None}


def run_reference_arm(a):
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    cfg_path = os.path.join(ROOT, "configs", a.config + ".json")
    total_blocks = a.gpus * a.bytes_per_gpu // BLOCK
    sample = a.ref_sample_blocks
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for _ in range(a.warmup):
            cpu_reference_pass(cfg_path, a.kind, a.seed, total_blocks, max(cores * 256, sample // 8), pool, cores)
        walls, nbytes, info = [], 0, None
        for _ in range(a.steps):
            info = cpu_reference_pass(cfg_path, a.kind, a.seed, total_blocks, sample, pool, cores)
            walls.append(info["wall_s"])
            nbytes += info["bytes"]
    total_s = sum(walls)
    gbs = nbytes / total_s / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * total_s / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(a),
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": info["kind"],
                         "sample": f"{info['bytes'] // BLOCK} blocks ({info['bytes'] / 2**20:.1f} MiB) of the workload per step, "
                                   f"one single-threaded reference object per core on disjoint slices"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(a):
    return {"workload": f"{a.bytes_per_gpu / GIB:g} GiB per GPU of synthetic '{a.kind}' 128-byte blocks "
                        f"(BASELINE.json configs[1]: 1 GB smooth fp32 array dump) x {a.gpus} GPU(s), MPC config configs/{a.config}.json",
            "mpc_config": a.config, "kind": a.kind, "seed": a.seed, "block_bytes": BLOCK,
            "blocks_per_gpu": a.bytes_per_gpu // BLOCK, "parallelism": f"shard{a.gpus}",
            "l2_policy": "input per step (1 GiB) is larger than the 126 MB L2; no flush needed",
            "collective": "none on the data path; one NCCL all_reduce of the 144 KB statistics vector closes the timed region (N > 1)"}


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    mpcb = importlib.import_module("cal_22-mpc_b200")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmpc_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg_path = os.path.join(ROOT, "configs", a.config + ".json")
    m = mpcb.Mpc(cfg_path, device=local)
    if a.kernel is not None:
        m.set_kernel(a.kernel)
    stream = torch.cuda.Stream()  # not the legacy default stream: its handle is 0, which the ABI reads as "own stream"
    torch.cuda.set_stream(stream)
    m.set_stream(stream.cuda_stream)

    n = a.bytes_per_gpu // BLOCK
    total = n * world
    first = rank * n  # contiguous shard [rank*n, (rank+1)*n) of the N GiB dump, SURVEY.md section 8e
    d = torch.empty(n * BLOCK, dtype=torch.uint8, device="cuda")
    m.synth_device(d.data_ptr(), first, n, total, a.kind, a.seed)
    sptr, swords = m.stats_device_ptr()

    class _Raw:  # expose the library's statistics vector to torch without a copy
        __cuda_array_interface__ = {"shape": (swords,), "typestr": "<i8", "data": (sptr, False), "version": 3}
    stats_t = torch.as_tensor(_Raw(), device="cuda")
    reduced = torch.empty_like(stats_t)

    def step():
        m.submit_device(d.data_ptr(), n, None)

    def exchange():
        # the path's only collective: ONE all-reduce of the statistics vector -- histograms,
        # totals, residue sums -- when the stream of batches ends; NCCL sum of int64 words over NVLink
        if world > 1:
            reduced.copy_(stats_t)
            dist.all_reduce(reduced)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    m.reset()
    for _ in range(a.warmup):
        step()
    exchange()
    barrier()
    m.reset()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    # The timed region is K back-to-back launches between two events on the launching stream, nothing else in the
    # stream; the kernel's average launch duration for the roofline is that span / K (it includes the launch gaps, so it
    # is an upper bound of the kernel time).  --per-launch-events brackets every launch with its own pair instead.
    m.enable_timing(False)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)] if a.per_launch_events else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ek = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(a.steps):
        if ev:
            ev[i][0].record(stream)
        m.submit_device(d.data_ptr(), n, None)
        if ev:
            ev[i][1].record(stream)
    ek.record(stream)
    exchange()
    e1.record(stream)
    barrier()
    t1 = time.perf_counter()
    m.enable_timing(True)
    total_ms = e0.elapsed_time(e1)
    kernel_ms = [s.elapsed_time(e) for s, e in ev] if ev else [e0.elapsed_time(ek) / a.steps]
    clocks = sampler.stop(t0, t1) if sampler else None
    if world > 1:
        tmax = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        total_ms = float(tmax.item())
        kmax = torch.tensor([sum(kernel_ms) / len(kernel_ms)], device="cuda", dtype=torch.float64)
        dist.all_reduce(kmax, op=dist.ReduceOp.MAX)
        kernel_avg_ms = float(kmax.item())
    else:
        kernel_avg_ms = sum(kernel_ms) / len(kernel_ms)

    # statistics of the timed region: K passes over the same shard; ratio is pass-invariant
    if world > 1:
        st = m.expand(reduced.cpu().numpy().view(np.uint64))
    else:
        st = m.finish()
    ratio = st.CompRatio

    # ---- end-to-end through the C ABI with HOST buffers (pinned), H2D inside the timed region ----
    e2e_blocks = min(n, a.e2e_bytes // BLOCK)
    host = torch.empty(e2e_blocks * BLOCK, dtype=torch.uint8).pin_memory()
    host.copy_(d[: e2e_blocks * BLOCK].cpu())
    m.set_stream(None)
    m.reset()
    for _ in range(2):
        m.submit_host_ptr(host.data_ptr(), e2e_blocks)
        m.finish()
    barrier()
    te0 = time.perf_counter()
    for _ in range(a.e2e_steps):
        m.submit_host_ptr(host.data_ptr(), e2e_blocks)
        m.finish()  # D2H of the statistics vector = the step's result
    barrier()
    e2e_s = time.perf_counter() - te0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_gbs = world * a.e2e_steps * e2e_blocks * BLOCK / e2e_s / 1e9

    if rank == 0:
        peak, peak_src = read_peaks()
        value = world * a.steps * n * BLOCK / (total_ms * 1e-3) / 1e9
        achieved = n * BLOCK / (kernel_avg_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(m.kernel_name(), {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(a),
            "kernel": m.kernel_name(), "comp_ratio": ratio,
            "clocks": clocks,
            "e2e": {"value": e2e_gbs, "unit": "GB/s", "h2d_bytes_per_step": e2e_blocks * BLOCK,
                    "d2h_bytes_per_step": int(swords) * 8, "steps": a.e2e_steps,
                    "api": "mpc_submit_host + mpc_finish (pinned host buffer, chunked double-buffered H2D)"},
            "gpu_launches": a.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": n * BLOCK, "kernel_ms": kernel_avg_ms},
        }
        if a.cpu_baseline and world >= 1:
            import multiprocessing as mp
            cores = os.cpu_count() or 1
            with mp.get_context("spawn").Pool(cores) as pool:
                info = cpu_reference_pass(cfg_path, a.kind, a.seed, total, a.ref_sample_blocks, pool, cores)
            line["cpu_baseline"] = {
                "value": info["bytes"] / info["wall_s"] / 1e9, "unit": "GB/s", "cores": cores, "kind": info["kind"],
                "sample": f"{info['bytes'] // BLOCK} blocks ({info['bytes'] / 2**20:.1f} MiB) of the same workload, one "
                          f"single-threaded reference object per core; ratio on the sample {info['ratio']}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="F4")
    ap.add_argument("--kind", default="smooth_f32")
    ap.add_argument("--seed", type=int, default=2024)
    ap.add_argument("--bytes-per-gpu", type=int, default=GIB)
    ap.add_argument("--kernel", type=int, default=None, help="0 auto, 1 generic warp kernel, 2 specialised kernel")
    ap.add_argument("--e2e-bytes", type=int, default=GIB)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-sample-blocks", type=int, default=262144,
                    help="blocks of the workload the CPU reference compresses per pass (bounded sample)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--per-launch-events", action="store_true", help="bracket every launch with its own CUDA event pair")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "ours":
        a.warmup = 3
    if a.impl == "reference":
        return run_reference_arm(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
