#!/usr/bin/env python3
"""bench.py -- headline benchmark of the MPC per-block compression path on B200.

Metric (BASELINE.json): GB/s of memory blocks compressed, device-timed, whole job over N GPUs.
A step = one pass of the hot path over one batch of synthetic input: every rank compresses its
contiguous shard (1 GiB = 8 388 608 blocks of 128 B) of an N GiB synthetic dump that is already
resident in HBM, and writes every block's (selected predictor, compressed size) -- north_star step (3).
The job's only collective -- one NCCL all-reduce of the statistics vector (N > 1), issued by the library
itself (mpc_allreduce_stats, include/mpc_capi.h) -- runs once, after the last step, inside the timed region.
Workload at N = 1 is BASELINE.json configs[1]: "1 GB synthetic fp32 array dump (smooth values,
delta-friendly)", compressed with configs/F4.json.

  python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--config F4 --kind smooth_f32]
  torchrun ... bench.py --gpus N ...        (one rank per GPU; rank 0 prints ONE JSON line)

Beside the headline the line carries (unless --no-extras): `parity` (a seeded window of every rank's shard checked
against the CPU oracle + totals = sum over ranks), `sustained` (the same launch back to back for >= 2 s with clocks),
`configs.mixed4g` (BASELINE configs[2]: 4 GiB mixed dump, F4 and P6; `configs.mixed4g_by_region` = the same classes laid out
by region, MPC and BDI / FPC / BPC), `configs.short_lines` (32- / 64-byte lines under S32 / S64: specialised vs generic
kernel), `configs.dump64g` (configs[3]: the 64 GiB dump,
sharded over the N GPUs), `configs.variants` (configs[4]: BDI / FPC / BPC / SC2 / PATTERN / CPACK over the mixed dump with the reference's
CPU rate beside each), `e2e` with the plain host->device copy ceiling of the same run, and `e2e_file` (an .npy in the
page cache through bin/compressor).

--impl reference times the reference's own CPU implementation (oracle/_ref, the unmodified sources
compiled by oracle/build_ref.sh; one single-threaded reference object per host core on disjoint slices)
on a bounded sample of the same workload.
"""
import argparse
import importlib
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
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
        sm, smax, reasons, power = [], None, set(), []
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
                try:
                    power.append(float(f[3]))
                except ValueError:
                    pass
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
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference classes on the host cores
# ------------------------------------------------------------------------------------------------------------
def _ref_worker(args):
    alg, cfg_path, kind, seed, first, n, total = args
    from oracle.bridge import RefCompressor
    from tools.gen_dump import synth
    blocks = synth(kind, seed, first, n, total)
    ref = RefCompressor(alg, cfg_path if alg == "VPC" else None)
    t0 = time.perf_counter()
    ref.compress(blocks, want_sels=False)
    dt = time.perf_counter() - t0
    orig, comp, _ = ref.totals()
    return dt, orig, comp


def _port_worker(args):
    alg, cfg_path, kind, seed, first, n, total = args
    from oracle.bridge import OracleMPC, oracle_variant
    from tools.gen_dump import synth
    blocks = synth(kind, seed, first, n, total)
    if alg == "SC2":
        from oracle.bridge import oracle_sc2
        t0 = time.perf_counter()
        sizes = oracle_sc2(blocks, 10000)
        return time.perf_counter() - t0, n * BLOCK * 8, int(sizes.astype(np.uint64).sum())
    if alg != "VPC":
        t0 = time.perf_counter()
        sizes, _ = oracle_variant(alg, blocks)
        return time.perf_counter() - t0, n * BLOCK * 8, int(sizes.astype(np.uint64).sum())
    o = OracleMPC(cfg_path)
    t0 = time.perf_counter()
    r = o.run(blocks, threads=1)
    dt = time.perf_counter() - t0
    return dt, r.OriginalSize, r.CompressedSize


def cpu_reference_pass(cfg_path, kind, seed, total_blocks, sample_blocks, pool, cores, alg="VPC"):
    """One bounded pass: `sample_blocks` blocks of the workload split over `cores` reference objects."""
    from oracle.bridge import have_ref
    worker, kindname = (_ref_worker, "reference") if have_ref() else (_port_worker, "port")
    per = max(1, sample_blocks // cores)
    jobs = [(alg, cfg_path, kind, seed, i * per, per, total_blocks) for i in range(cores)]
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
            "per_block_output": "on: every launch writes 2 B per block (size | (selected + 1) << 11), north_star step (3)",
            "l2_policy": "input per step (1 GiB) is larger than the 126 MB L2; no flush needed",
            "collective": "none on the data path; one ncclAllReduce (uint64 sum) of the 144 KB statistics vector, issued by "
                          "the library (mpc_allreduce_stats), closes the timed region (N > 1)"}


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
class Job:
    """Process-wide state of one rank."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.a = a
        self.mpcb = importlib.import_module("cal_22-mpc_b200")
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != a.gpus and self.world > 1:
            raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={self.world}")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: libmpc_b200 has no CPU path")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.stream = torch.cuda.Stream()  # not the legacy default stream: its handle is 0, which the ABI reads as "own stream"
        torch.cuda.set_stream(self.stream)
        self.peak, self.peak_src = read_peaks()
        self.uid = None

    def cfg_path(self, name):
        return os.path.join(ROOT, "configs", name + ".json")

    def context(self, cfg):
        """One library context on this rank's GPU, on the bench stream, with the job's communicator (N > 1)."""
        m = self.mpcb.Mpc(self.cfg_path(cfg), device=self.local)
        if self.a.kernel is not None:
            m.set_kernel(self.a.kernel)
        m.set_stream(self.stream.cuda_stream)
        if self.world > 1:
            # torch.distributed is the plumbing: it only ships the 128-byte NCCL id; the communicator and the
            # all-reduce are the library's own (mpc_comm_init_rank / mpc_allreduce_stats)
            t = self.torch.zeros(self.mpcb.capi.COMM_UID_BYTES, dtype=self.torch.uint8, device="cuda")
            if self.rank == 0:
                t.copy_(self.torch.frombuffer(bytearray(self.mpcb.Mpc.comm_unique_id()), dtype=self.torch.uint8))
            self.dist.broadcast(t, 0)
            m.comm_init_rank(bytes(t.cpu().numpy().tobytes()), self.world, self.rank)
        return m

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, vals):
        if self.world == 1:
            return [int(v) for v in vals]
        t = self.torch.tensor([int(v) for v in vals], device="cuda", dtype=self.torch.int64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(v) for v in t.cpu().tolist()]

    def timed_launches(self, m, d_ptr, n, packed_ptr, reps, warm=2, exchange=False):
        """`reps` back-to-back launches between two events on the launching stream (+ the all-reduce when asked):
        -> (total ms, ms up to the last kernel), each the max over ranks."""
        torch = self.torch
        for _ in range(warm):
            m.submit_device(d_ptr, n, packed_ptr)
        if exchange:
            m.allreduce_stats()
        self.barrier()
        m.reset()
        m.enable_timing(False)
        e0, ek, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        self.barrier()
        e0.record(self.stream)
        for _ in range(reps):
            m.submit_device(d_ptr, n, packed_ptr)
        ek.record(self.stream)
        if exchange:
            m.allreduce_stats()
        e1.record(self.stream)
        self.barrier()
        m.enable_timing(True)
        return self.max_over_ranks(e0.elapsed_time(e1)), self.max_over_ranks(e0.elapsed_time(ek))

    def parity_window(self, m, cfg, d, packed, n, steps, wn=65536):
        """Checker (not measured): a seeded window of this rank's shard against the CPU oracle, the local totals against the
        per-block results, and -- over the job -- the all-reduced totals against the sum of the ranks' totals."""
        from oracle.bridge import OracleMPC
        torch = self.torch
        wn = min(wn, n)
        w0 = ((0x9E3779B97F4A7C15 * (self.rank + 1) + self.a.seed) % (1 << 63)) % max(1, n - wn + 1)
        blocks = d[w0 * BLOCK:(w0 + wn) * BLOCK].cpu().numpy().reshape(-1, BLOCK)
        got_sizes, got_sels = self.mpcb.unpack(packed[w0:w0 + wn].cpu().numpy().view(np.uint16))
        r = OracleMPC(self.cfg_path(cfg)).run(blocks)
        mism = int(np.count_nonzero((got_sizes != r.sizes) | (got_sels != r.sels)))
        local = m.finish()  # this rank's own statistics (the all-reduce works on a copy)
        p = packed.view(torch.int32)
        sum_sizes = int((p & 0x7FF).to(torch.int64).sum().item() + ((p >> 16) & 0x7FF).to(torch.int64).sum().item()) if n % 2 == 0 \
            else int((packed.to(torch.int32) & 0x7FF).to(torch.int64).sum().item())
        local_ok = (local.blocks == steps * n) and (local.CompressedSize == steps * sum_sizes)
        m.allreduce_stats()  # all-reduce (a copy of) the statistics of the launches above
        red = m.reduced_stats()
        tot = self.sum_over_ranks([local.blocks, local.CompressedSize, mism, 0 if local_ok else 1])
        reduced_ok = (red.blocks == tot[0]) and (red.CompressedSize == tot[1])
        return {"blocks_checked": wn * self.world, "mismatches": tot[2], "window": "65 536 consecutive blocks per rank at a seeded offset, "
                "per-block (selected, size) vs the CPU oracle", "local_totals_equal_sum_of_per_block_results": tot[3] == 0,
                "reduced_totals_equal_sum_over_ranks": bool(reduced_ok), "ratio": red.CompRatio}, red


def int_roofline(job, kernel_name, n_blocks, kernel_ms):
    """Roofline B: counted instructions per block (ncu, profiles/roofline_traffic.json) against the
    integer issue rates tools/int_peak measures on THIS GPU in this run."""
    exe = os.path.join(ROOT, "tools", "int_peak")
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not (os.path.exists(exe) and os.path.exists(tp)):
        return None
    try:
        peaks = json.loads(subprocess.run([exe, str(job.local)], capture_output=True, text=True, timeout=120, check=True).stdout)
        prof = json.load(open(tp)).get(kernel_name, {})
        ipb = prof.get("inst_per_block")
        if not ipb:
            return {"peaks": peaks, "note": "no instruction count recorded for " + kernel_name}
        sms = peaks["sms"]
        clk = peaks["sm_clock_mhz_nominal"] * 1e6
        # per-pipe time: warp instructions of the launch / (measured warp-inst per clk per SM x SMs x clock)
        warp_inst = {k: v * n_blocks / 32.0 for k, v in ipb.items()}
        t_alu = warp_inst.get("alu", 0) / (peaks["lop3"]["warp_inst_per_clk_per_sm"] * sms * clk)
        t_fma = warp_inst.get("fma", 0) / (peaks["imad"]["warp_inst_per_clk_per_sm"] * sms * clk)
        t_issue = warp_inst.get("total", 0) / (4.0 * sms * clk)  # one warp instruction per clock per scheduler
        # shared-memory data pipe: one 128-byte wavefront per clock per SM (row-cost lookups with their bank conflicts + tile staging)
        t_smem = prof.get("smem_wavefronts_per_block", 0) * n_blocks / 32.0 / (sms * clk)
        t_int = max(t_alu, t_fma, t_issue, t_smem)
        return {"ops_per_block": ipb, "smem_wavefronts_per_block": prof.get("smem_wavefronts_per_block"), "t_smem_ms": 1e3 * t_smem,
                "source": prof.get("inst_source"),
                "peak_measured": {k: peaks[k]["warp_inst_per_clk_per_sm"] for k in ("lop3", "iadd3", "imad", "prmt", "shf", "vimnmx_u16x2", "idp4a", "mix_lop3_imad")},
                "peak_unit": "warp instructions per clock per SM (tools/int_peak, this run)",
                "t_alu_ms": 1e3 * t_alu, "t_fma_ms": 1e3 * t_fma, "t_issue_ms": 1e3 * t_issue, "t_int_ms": 1e3 * t_int,
                "frac": (1e3 * t_int) / kernel_ms if kernel_ms else None}
    except Exception as e:  # the micro-benchmark is evidence, never a reason to lose the headline
        return {"error": repr(e)}


def run_headline(job, line):
    a, torch = job.a, job.torch
    world, rank = job.world, job.rank
    m = job.context(a.config)
    n = a.bytes_per_gpu // BLOCK
    total = n * world
    first = rank * n  # contiguous shard [rank*n, (rank+1)*n) of the N GiB dump, SURVEY.md section 8e
    d = torch.empty(n * BLOCK, dtype=torch.uint8, device="cuda")
    packed = torch.zeros(n, dtype=torch.int16, device="cuda")
    m.synth_device(d.data_ptr(), first, n, total, a.kind, a.seed)

    def step():
        m.submit_device(d.data_ptr(), n, packed.data_ptr())

    m.reset()
    for _ in range(a.warmup):
        step()
    m.allreduce_stats()
    job.barrier()
    m.reset()
    sampler = ClockSampler(job.local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    job.barrier()
    # The timed region is K back-to-back launches between two events on the launching stream, nothing else in the
    # stream; the kernel's average launch duration for the roofline is that span / K (it includes the launch gaps, so it
    # is an upper bound of the kernel time).  --per-launch-events brackets every launch with its own pair instead.
    m.enable_timing(False)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)] if a.per_launch_events else None
    e0, e1, ek = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.perf_counter()
    e0.record(job.stream)
    for i in range(a.steps):
        if ev:
            ev[i][0].record(job.stream)
        step()
        if ev:
            ev[i][1].record(job.stream)
    ek.record(job.stream)
    m.allreduce_stats()  # the path's only collective: ncclAllReduce of the statistics vector, inside the library
    e1.record(job.stream)
    job.barrier()
    t1 = time.perf_counter()
    m.enable_timing(True)
    total_ms = job.max_over_ranks(e0.elapsed_time(e1))
    kernel_ms = [s.elapsed_time(e) for s, e in ev] if ev else [e0.elapsed_time(ek) / a.steps]
    kernel_avg_ms = job.max_over_ranks(sum(kernel_ms) / len(kernel_ms))
    clocks = sampler.stop(t0, t1) if sampler else None

    parity, red = job.parity_window(m, a.config, d, packed, n, a.steps)

    # the same launches without the per-block output (what round 1 reported as the headline)
    _, k_np = job.timed_launches(m, d.data_ptr(), n, None, a.steps)
    no_packed_gbs = world * a.steps * n * BLOCK / (k_np * 1e-3) / 1e9

    sustained = None
    if a.extras:
        # >= 2 s of the same launch back to back: does the integer kernel hold its clock where a matmul does not?
        reps = max(a.steps, int(a.sustained_s * 1e3 / max(kernel_avg_ms, 1e-3)))
        s2 = ClockSampler(job.local) if rank == 0 else None
        if s2:
            s2.start()
            time.sleep(0.3)
        ts0 = time.perf_counter()
        _, k_ms = job.timed_launches(m, d.data_ptr(), n, packed.data_ptr(), reps, warm=0)
        ts1 = time.perf_counter()
        c2 = s2.stop(ts0, ts1) if s2 else None
        sustained = {"launches": reps, "seconds": k_ms * 1e-3, "value": world * reps * n * BLOCK / (k_ms * 1e-3) / 1e9, "unit": "GB/s",
                     "frac": (n * (BLOCK + 2) / (k_ms / reps * 1e-3) / 1e9) / job.peak, "clocks": c2}

    # ---- end-to-end through the C ABI with HOST buffers (pinned), H2D inside the timed region ----
    e2e_blocks = min(n, a.e2e_bytes // BLOCK)
    host = torch.empty(e2e_blocks * BLOCK, dtype=torch.uint8).pin_memory()
    host.copy_(d[: e2e_blocks * BLOCK].cpu())
    m.set_stream(None)
    m.reset()
    for _ in range(2):
        m.submit_host_ptr(host.data_ptr(), e2e_blocks)
        m.finish()
    job.barrier()
    te0 = time.perf_counter()
    for _ in range(a.e2e_steps):
        m.submit_host_ptr(host.data_ptr(), e2e_blocks)
        m.finish()  # D2H of the statistics vector = the step's result
    job.barrier()
    e2e_s = job.max_over_ranks(time.perf_counter() - te0)
    e2e_gbs = world * a.e2e_steps * e2e_blocks * BLOCK / e2e_s / 1e9
    # the ceiling of that path on this box, in the same run and on all ranks at once: plain cudaMemcpyAsync of the same
    # pinned buffer to the device, nothing else
    dst = torch.empty(e2e_blocks * BLOCK, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        dst.copy_(host, non_blocking=True)
    job.barrier()
    tl0 = time.perf_counter()
    for _ in range(a.e2e_steps):
        dst.copy_(host, non_blocking=True)
    job.barrier()
    link_s = job.max_over_ranks(time.perf_counter() - tl0)
    link_gbs = world * a.e2e_steps * e2e_blocks * BLOCK / link_s / 1e9
    del dst, host
    m.set_stream(job.stream.cuda_stream)

    sptr, swords = m.stats_device_ptr()
    value = world * a.steps * n * BLOCK / (total_ms * 1e-3) / 1e9
    alg_bytes = n * (BLOCK + 2)  # 128 B read + 2 B written per block (roofline A)
    achieved = alg_bytes / (kernel_avg_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(m.kernel_name(), {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    line.update({
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": workload_config(a),
        "kernel": m.kernel_name(), "comp_ratio": red.CompRatio,
        "clocks": clocks,
        "parity": parity,
        "value_without_per_block_output": no_packed_gbs,
        "e2e": {"value": e2e_gbs, "unit": "GB/s", "h2d_bytes_per_step": e2e_blocks * BLOCK,
                "d2h_bytes_per_step": int(swords) * 8, "steps": a.e2e_steps,
                "api": "mpc_submit_host + mpc_finish (pinned host buffer, chunked double-buffered H2D)",
                "link_peak": link_gbs, "frac_of_link_peak": e2e_gbs / link_gbs if link_gbs else None,
                "link_peak_how": "cudaMemcpyAsync of the same pinned buffer on all ranks at once, same run"},
        "gpu_launches": a.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": job.peak, "unit": "GB/s", "frac": achieved / job.peak,
                     "traffic": traffic, "peak_source": job.peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_avg_ms},
    })
    if sustained:
        line["sustained"] = sustained
    if rank == 0 and a.extras:
        ri = int_roofline(job, m.kernel_name(), n, kernel_avg_ms)
        if ri:
            line["roofline"]["int"] = ri
            if ri.get("t_int_ms") is not None:
                t_hbm = alg_bytes / (job.peak * 1e9) * 1e3
                # north_star: the roofline is the slower of the two bounds
                line["roofline"]["t_hbm_ms"] = t_hbm
                line["roofline"]["bound_slower_of_two"] = "int" if ri["t_int_ms"] > t_hbm else "hbm"
                line["roofline"]["frac_of_slower_bound"] = max(ri["t_int_ms"], t_hbm) / kernel_avg_ms
    del d, packed
    return m


def run_mixed4g(job, line):
    """BASELINE configs[2] (4 GiB mixed dump, every branch) per GPU under the column-major (F4) and plane-major (P6) kernels,
    with and without the per-block output, then configs[4]: the other compressors over the same dump."""
    a, torch = job.a, job.torch
    world, rank = job.world, job.rank
    n = a.mixed_bytes // BLOCK
    d = torch.empty(n * BLOCK, dtype=torch.uint8, device="cuda")
    packed = torch.zeros(n, dtype=torch.int16, device="cuda")
    out = {"workload": f"{a.mixed_bytes / GIB:g} GiB per GPU of 'mixed_hashed' blocks (class by hash of the block index), seed 31337"}
    synth_done = False
    for cfg in ("F4", "P6"):
        m = job.context(cfg)
        if not synth_done:
            m.synth_device(d.data_ptr(), rank * n, n, n * world, "mixed_hashed", 31337)
            synth_done = True
        reps = 5
        _, k = job.timed_launches(m, d.data_ptr(), n, packed.data_ptr(), reps)
        parity, red = job.parity_window(m, cfg, d, packed, n, reps)
        _, k_np = job.timed_launches(m, d.data_ptr(), n, None, reps)
        gbs = world * reps * n * BLOCK / (k * 1e-3) / 1e9
        out[cfg] = {"value": gbs, "unit": "GB/s", "kernel": m.kernel_name(),
                    "frac": (n * (BLOCK + 2) / (k / reps * 1e-3) / 1e9) / job.peak,
                    "value_without_per_block_output": world * reps * n * BLOCK / (k_np * 1e-3) / 1e9,
                    "frac_without_per_block_output": (n * BLOCK / (k_np / reps * 1e-3) / 1e9) / job.peak,
                    "comp_ratio": red.CompRatio, "parity": {k2: parity[k2] for k2 in ("blocks_checked", "mismatches", "reduced_totals_equal_sum_over_ranks")}}
        m.close()
    line.setdefault("configs", {})["mixed4g"] = out
    if a.variants:
        line["configs"]["variants"] = run_variants(job, d, n)
    # the same eight classes laid out by region (a dump of arrays: every warp sees one class) -- the other end of the range
    byreg = {"workload": "the same classes by region ('mixed_regions': class = position in the dump), same size and seed"}
    m = job.context("F4")
    m.synth_device(d.data_ptr(), rank * n, n, n * world, "mixed_regions", 31337)
    m.close()
    for cfg in ("F4", "P6"):
        m = job.context(cfg)
        reps = 5
        _, k = job.timed_launches(m, d.data_ptr(), n, packed.data_ptr(), reps)
        parity, red = job.parity_window(m, cfg, d, packed, n, reps)
        byreg[cfg] = {"value": world * reps * n * BLOCK / (k * 1e-3) / 1e9, "unit": "GB/s", "frac": (n * (BLOCK + 2) / (k / reps * 1e-3) / 1e9) / job.peak,
                      "comp_ratio": red.CompRatio, "parity": {k2: parity[k2] for k2 in ("blocks_checked", "mismatches", "reduced_totals_equal_sum_over_ranks")}}
        m.close()
    if a.variants:
        for alg in ("BDI", "FPC", "BPC"):
            for _ in range(2):
                _, vs, ms = job.mpcb.variant_run(alg, device_ptr=d.data_ptr(), n_blocks=n, device=job.local)
            ms = job.max_over_ranks(ms)
            byreg[alg] = {"value": world * n * BLOCK / (ms * 1e-3) / 1e9, "unit": "GB/s", "frac": (n * BLOCK / (ms * 1e-3) / 1e9) / job.peak}
    line["configs"]["mixed4g_by_region"] = byreg
    run_short_lines(job, line, d, packed)
    del d, packed


def run_short_lines(job, line, d, packed):
    """32- and 64-byte lines (GPGPU-Sim sectors; the reference takes any lineSize, VPC.cpp:99-101): one GiB of the
    hash-mixed dump cut into shorter lines under configs S32 / S64, specialised kernel (four / two lines per thread) and
    the generic warp-per-block kernel beside it, a seeded window of each against the CPU oracle."""
    from oracle.bridge import OracleMPC
    torch = job.torch
    nbytes = min(GIB, d.numel())
    out = {"workload": f"{nbytes / GIB:g} GiB per GPU of 'mixed_hashed' 128-byte blocks read as 32- / 64-byte lines, per-line output on"}
    m = job.context("F4")
    m.synth_device(d.data_ptr(), job.rank * (nbytes // BLOCK), nbytes // BLOCK, job.world * (nbytes // BLOCK), "mixed_hashed", 31337)
    m.close()
    for cfg, L in (("S32", 32), ("S64", 64)):
        n = nbytes // L
        pk = torch.zeros(n, dtype=torch.int16, device="cuda")
        row = {}
        for label, kern, reps in (("spec", 0, 5), ("generic", 1, 1)):
            m = job.context(cfg)
            if kern:
                m.set_kernel(kern)
            nn = n if not kern else n // 16  # the generic kernel is ~100x slower: a sixteenth of the lines
            _, k = job.timed_launches(m, d.data_ptr(), nn, pk.data_ptr(), reps, warm=1)
            row[label] = {"value": job.world * reps * nn * L / (k * 1e-3) / 1e9, "unit": "GB/s", "kernel": m.kernel_name()}
            if not kern:
                row[label]["frac"] = (nn * (L + 2) / (k / reps * 1e-3) / 1e9) / job.peak
                wn = 65536
                w0 = (n // 3) // wn * wn
                lines_h = d[w0 * L:(w0 + wn) * L].cpu().numpy().reshape(-1, L)
                got_sizes, got_sels = job.mpcb.unpack(pk[w0:w0 + wn].cpu().numpy().view(np.uint16))
                r = OracleMPC(job.cfg_path(cfg)).run(lines_h)
                mism = int(np.count_nonzero((got_sizes != r.sizes) | (got_sels != r.sels)))
                row["parity"] = {"lines_checked": wn * job.world, "mismatches": int(job.sum_over_ranks([mism])[0])}
            m.close()
        out[cfg] = row
        del pk
    line["configs"]["short_lines"] = out


def run_variants(job, d, n):
    """BDI / FPC / BPC / SC2 on this rank's shard of the mixed dump (stateless per line: sharded like MPC), device-timed by
    the library; the unmodified reference's CPU rate on a bounded sample of the same dump beside each."""
    import ctypes as C
    import multiprocessing as mp
    mpcb = job.mpcb
    res = {"workload": "the mixed4g dump; GPU rows are whole-job GB/s (max kernel time over ranks)"}
    for alg in ("BDI", "FPC", "BPC"):
        ms = None
        for _ in range(2):
            _, vs, ms = mpcb.variant_run(alg, device_ptr=d.data_ptr(), n_blocks=n, device=job.local)
        ms = job.max_over_ranks(ms)
        tot = job.sum_over_ranks([vs.original_bits, vs.compressed_bits])
        res[alg] = {"value": job.world * n * BLOCK / (ms * 1e-3) / 1e9, "unit": "GB/s", "frac": (n * BLOCK / (ms * 1e-3) / 1e9) / job.peak,
                    "comp_ratio": tot[0] / tot[1] if tot[1] else None}
    # SC2 is two-phase: the rank that holds the first S lines builds the code table (device sort +
    # host tree), the table (<= 1024 symbols) is broadcast, every rank applies it to its shard
    total_lines = n * job.world
    S = mpcb.sc2_sampling_lines(total_lines + 1)
    if S <= n:
        lib = mpcb.lib()
        table = mpcb.capi.Sc2Table()
        ms_build = 0.0
        for _ in range(2):
            job.barrier()
            t0 = time.perf_counter()
            if job.rank == 0:
                if lib.mpc_sc2_build_table(job.local, d.data_ptr(), S, BLOCK, C.byref(table)) != 0:
                    raise SystemExit("SC2: " + lib.mpc_sc2_error().decode())
            if job.world > 1:
                buf = [bytes(table) if job.rank == 0 else None]
                job.dist.broadcast_object_list(buf, src=0)
                C.memmove(C.byref(table), buf[0], C.sizeof(table))
            ms_build = (time.perf_counter() - t0) * 1e3
            vs, msf = mpcb.VariantStats(), C.c_float()
            if lib.mpc_sc2_apply_device(job.local, d.data_ptr(), n, job.rank * n, S, BLOCK, C.byref(table), None, C.byref(vs), C.byref(msf)) != 0:
                raise SystemExit("SC2: " + lib.mpc_sc2_error().decode())
        ms_apply = job.max_over_ranks(msf.value)
        ms_build = job.max_over_ranks(ms_build)
        tot = job.sum_over_ranks([vs.original_bits, vs.compressed_bits])
        ms = ms_build + ms_apply
        res["SC2"] = {"value": job.world * n * BLOCK / (ms * 1e-3) / 1e9, "unit": "GB/s", "frac": (n * BLOCK / (ms * 1e-3) / 1e9) / job.peak,
                      "comp_ratio": tot[0] / tot[1] if tot[1] else None, "table_ms": ms_build, "lookup_ms": ms_apply, "symbols": int(table.n),
                      "includes": f"table from the first {S} lines on rank 0 (device sort + host tree, wall clock" +
                                  (", broadcast" if job.world > 1 else "") + ") + device lookup over every shard"}
    # PATTERN (the reference's analysis tool) and CPACK carry state from line to line (a cache of lines / a dictionary), so they are
    # not sharded (replicas only): rank 0 runs them on the head of its shard
    if job.rank == 0:
        npat = min(n, 1 << 23)  # 1 GiB: at most as many distinct lines as the reference's cache holds, the temporal count stays on the device
        ps, ms = None, None
        for _ in range(2):
            _, ps, ms = mpcb.pattern_run(device_ptr=d.data_ptr(), n_blocks=npat, device=job.local)
        res["PATTERN"] = {"value": npat * BLOCK / (ms * 1e-3) / 1e9, "unit": "GB/s", "frac": (npat * BLOCK / (ms * 1e-3) / 1e9) / job.peak,
                          "lines": npat, "distinct_lines": int(ps.distinct_blocks), "temporal_path": int(ps.temporal_path),
                          "where": "rank 0 only, first 1 GiB of the dump: analysis kernel + 64-bit hash sort + duplicate pass, device-timed by the library"}
        ncp = min(n, (64 << 20) // BLOCK)
        host = d[: ncp * BLOCK].cpu().numpy()
        t0 = time.perf_counter()
        _, cs = mpcb.cpack_run(host)
        dt = time.perf_counter() - t0
        res["CPACK"] = {"value": ncp * BLOCK / dt / 1e9, "unit": "GB/s", "lines": ncp,
                        "comp_ratio": cs.original_bits / cs.compressed_bits if cs.compressed_bits else None,
                        "where": "rank 0 only, first 64 MiB of the dump, ONE host thread: the dictionary persists across lines, sequential by construction"}
    if job.rank == 0 and job.a.cpu_baseline:
        cores = os.cpu_count() or 1
        with mp.get_context("spawn").Pool(cores) as pool:
            for alg in ("BDI", "FPC", "BPC", "SC2"):
                if alg not in res:
                    continue
                info = cpu_reference_pass(None, "mixed_hashed", 31337, n * job.world, job.a.ref_sample_blocks, pool, cores, alg=alg)
                res[alg]["cpu_reference"] = {"value": info["bytes"] / info["wall_s"] / 1e9, "unit": "GB/s", "cores": cores, "kind": info["kind"],
                                             "sample_blocks": info["bytes"] // BLOCK, "ratio_on_sample": info["ratio"]}
            for alg, sample in (("PATTERN", 32768), ("CPACK", 262144)):  # stateful: one reference object on one core
                if not have_ref_alg(alg):
                    continue
                info = cpu_reference_pass(None, "mixed_hashed", 31337, n * job.world, sample, pool, 1, alg=alg)
                res[alg]["cpu_reference"] = {"value": info["bytes"] / info["wall_s"] / 1e9, "unit": "GB/s", "cores": 1, "kind": info["kind"],
                                             "sample_blocks": info["bytes"] // BLOCK}
    return res


def have_ref_alg(alg):
    """PATTERN / CPACK exist only in the unmodified reference build (the plain-C oracle port times neither through this path)."""
    from oracle.bridge import have_ref
    return have_ref()


def run_dump64g(job, line):
    """BASELINE configs[3]: the 64 GiB dump sharded over the N GPUs (N = 1: it fits in one B200's HBM), one pass = one
    launch per GPU + the all-reduce of the statistics."""
    a, torch = job.a, job.torch
    world, rank = job.world, job.rank
    total_blocks = a.dump_bytes // BLOCK
    n = total_blocks // world
    try:
        d = torch.empty(n * BLOCK, dtype=torch.uint8, device="cuda")
        packed = torch.zeros(n, dtype=torch.int16, device="cuda")
    except Exception as e:  # a smaller GPU: say so instead of failing the headline
        line.setdefault("configs", {})["dump64g"] = {"skipped": repr(e)[:200]}
        return
    m = job.context(a.config)
    m.synth_device(d.data_ptr(), rank * n, n, total_blocks, a.kind, a.seed + 64)
    reps = 3
    t_all, k = job.timed_launches(m, d.data_ptr(), n, packed.data_ptr(), reps, warm=1, exchange=True)
    parity, red = job.parity_window(m, a.config, d, packed, n, reps)
    line.setdefault("configs", {})["dump64g"] = {
        "workload": f"{a.dump_bytes / GIB:g} GiB '{a.kind}' dump, {n * BLOCK / GIB:g} GiB per GPU x {world}, configs/{a.config}.json, per-block output on",
        "value": reps * total_blocks * BLOCK / (t_all * 1e-3) / 1e9, "unit": "GB/s", "ms_per_pass": t_all / reps, "scaling": "strong",
        "frac_per_gpu": (n * (BLOCK + 2) / (k / reps * 1e-3) / 1e9) / job.peak, "comp_ratio": red.CompRatio, "parity": parity}
    m.close()
    del d, packed


def run_e2e_file(job, line):
    """The drop-in CLI end to end: an .npy dump in the page cache -> bin/compressor (mmap -> pinned ring -> H2D -> kernel ->
    statistics -> CSV).  Wall time of the whole process (CUDA context creation included) and of its compress loop."""
    a = job.a
    exe = os.path.join(ROOT, "bin", "compressor")
    if job.rank != 0 or not os.path.exists(exe):
        return
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 3 * a.file_bytes else tempfile.gettempdir()
    tmp = tempfile.mkdtemp(prefix="mpc_e2e_", dir=base)
    try:
        from tools.gen_dump import synth
        n = a.file_bytes // BLOCK
        ds = os.path.join(tmp, "ds")
        os.makedirs(ds)
        path = os.path.join(ds, "dump_set.npy")
        arr = np.lib.format.open_memmap(path, mode="w+", dtype=np.uint8, shape=(n + 1, BLOCK))
        step = 1 << 20
        for lo in range(0, n, step):
            arr[lo:min(n, lo + step)] = synth(a.kind, a.seed, lo, min(step, n - lo), n)
        arr.flush()
        del arr
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            r = subprocess.run([exe, "-a", "VPC", "-i", path, "-c", job.cfg_path(a.config), "-o", tmp, "--time"], capture_output=True, text=True)
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                line["e2e_file"] = {"error": (r.stdout + r.stderr)[-300:]}
                return
            loop_s = None
            for ln in r.stderr.splitlines():
                if ln.startswith("kernel ") and " wall " in ln:
                    loop_s = float(ln.split(" wall ")[1].split(" s")[0])
            if best is None or wall < best[0]:
                best = (wall, loop_s, r.stdout.strip())
        line["e2e_file"] = {"value": n * BLOCK / best[1] / 1e9 if best[1] else None, "unit": "GB/s",
                            "what": "compress loop of bin/compressor (loader -> CompressBatch -> GetResult) over a "
                                    f"{a.file_bytes / GIB:g} GiB .npy in the page cache ({base})",
                            "process_wall_s": best[0], "process_gbs": n * BLOCK / best[0] / 1e9, "stdout": best[2]}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_ours(a):
    job = Job(a)
    line = {}
    m = run_headline(job, line)
    cfg_path = job.cfg_path(a.config)
    if a.extras:
        m.close()
        run_mixed4g(job, line)
        run_dump64g(job, line)
        if job.world == 1:
            run_e2e_file(job, line)
    if job.rank == 0:
        if a.cpu_baseline:
            import multiprocessing as mp
            cores = os.cpu_count() or 1
            total = job.world * (a.bytes_per_gpu // BLOCK)
            with mp.get_context("spawn").Pool(cores) as pool:
                info = cpu_reference_pass(cfg_path, a.kind, a.seed, total, a.ref_sample_blocks, pool, cores)
            line["cpu_baseline"] = {
                "value": info["bytes"] / info["wall_s"] / 1e9, "unit": "GB/s", "cores": cores, "kind": info["kind"],
                "sample": f"{info['bytes'] // BLOCK} blocks ({info['bytes'] / 2**20:.1f} MiB) of the same workload, one "
                          f"single-threaded reference object per core; ratio on the sample {info['ratio']}"}
        print(json.dumps(line), flush=True)
    if job.world > 1:
        job.dist.barrier()
        job.dist.destroy_process_group()
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
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="headline only (no mixed4g / dump64g / variants / sustained / e2e_file)")
    ap.add_argument("--no-variants", dest="variants", action="store_false")
    ap.add_argument("--mixed-bytes", type=int, default=4 * GIB)
    ap.add_argument("--dump-bytes", type=int, default=64 * GIB)
    ap.add_argument("--file-bytes", type=int, default=GIB)
    ap.add_argument("--sustained-s", type=float, default=2.0)
    ap.add_argument("--per-launch-events", action="store_true", help="bracket every launch with its own CUDA event pair")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "ours":
        a.warmup = 3
    if a.impl == "reference":
        return run_reference_arm(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
