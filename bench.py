#!/usr/bin/env python
"""Benchmark of the hot path: fused linear field gather + Boris push (fp64).

    python bench.py --gpus 1 --steps 100 --warmup 5              # our arm
    python bench.py --impl reference --gpus 1 --steps 3          # CPU reference arm
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N  # one rank per GPU

Workload (BASELINE.json configs[1]): the particle-in-field example scaled up --
16 Mi electrons per GPU in a 1-D EM plane wave E_y(r,t) = E0 cos(2 pi (k (r-0.5) - w t))
on a uniform grid, B = 0, linear gather at position[:,0] (Interpolators /
numpy.interp rounding), relativistic Boris push; one "step" = one gather+push
pass over every particle.  Particles shard by index across ranks (weak scaling:
16 Mi per GPU), the grid is replicated, no data-path collective.

One JSON line on stdout (rank 0).  `value` = particle pushes per second over all
ranks with state resident in HBM; `e2e` = the same metric through
BorisPush.gather_push on HOST (pinned numpy) arrays, H2D and D2H inside the timed
region; `roofline` = 96 algorithmic bytes per particle-step over the average
kernel time against the measured HBM copy peak; `cpu_baseline` = the oracle's
numpy port of the reference timed on this box's host cores.
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

Q_E, M_E = -1.6022e-19, 9.1094e-31             # particle_in_field.py:49-50 magnitudes (electron)
OMEGA, C_WAVE, E0 = 2e8, 2.998e8, 1.0          # particle_in_field.toml:15-17, particle_in_field.py:15
DT = 1e-8 / 20                                 # particle_in_field.toml:7-10
BYTES_PER_PUSH = 96                            # 6 fp64 read + 6 fp64 written per particle-step
METRIC = "Boris particle-pushes/sec (fp64, gather+push)"
WORKLOAD = ("configs[1]: pif electron ensemble in 1D EM plane wave, linear field gather "
            "(E_y on grid, B=0) + Boris push")
UNIT = "pushes/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=int, default=16 << 20, help="particles per GPU")
    ap.add_argument("--ng", type=int, default=4097, help="grid nodes")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-particles", type=int, default=1 << 20)
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay the timed steps as one CUDA graph")
    return ap.parse_args()


# ------------------------------------------------------------------ workload
def grid_points(ng):
    from oracle import grid as ogrid            # Grid.r restatement (host-side setup only)
    return ogrid.grid_points(0.0, 1.0, ng), ogrid.grid_dr(0.0, 1.0, ng)


def wave_fields(r, nlevels):
    """EMWave.update for time levels 0..nlevels-1 (particle_in_field.py:32-34)."""
    k = OMEGA / C_WAVE
    t = DT * np.arange(nlevels)[:, None]
    return E0 * np.cos(2 * np.pi * (-OMEGA * t + k * (r[None, :] - 0.5)))


def host_ensemble(n, seed):
    """pif-style ensemble: x uniform in [0.05, 0.95], y = z = 0, p = 0 (seed 2345 + rank)."""
    rng = np.random.default_rng(seed)
    pos = np.zeros((n, 3))
    pos[:, 0] = rng.uniform(0.05, 0.95, n)
    mom = np.zeros((n, 3))
    return pos, mom


class Owner:
    """The two attributes of a turboPy Simulation the pusher reads."""

    def __init__(self, r, dr):
        from types import SimpleNamespace
        self.clock = SimpleNamespace(dt=DT)
        self.grid = SimpleNamespace(r=r, dr=dr, num_points=len(r), r_min=float(r[0]), r_max=float(r[-1]))


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled
    every ~2 ms from a thread (the timed region lasts tens of milliseconds, too
    short for nvidia-smi's loop mode, which is only the fallback)."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index = index
        self.samples = []          # (sm_mhz, reasons bitmask)
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        self._smi = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch's device index -> physical index when CUDA_VISIBLE_DEVICES remaps
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None
            self._start_smi()

    def _poll(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM)
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self._handle)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle)
                self.samples.append((float(mhz), int(bits)))
            except Exception:
                break
            time.sleep(0.002)

    def _start_smi(self):
        query = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self._smi = subprocess.Popen(["nvidia-smi", f"--query-gpu={query}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read_smi, daemon=True).start()
        except OSError:
            self._smi = None

    def _read_smi(self):
        names = list(self.REASONS)
        for line in self._smi.stdout:
            cols = [c.strip() for c in line.split(",")]
            try:
                mhz, self.max_mhz = float(cols[0]), float(cols[1])
            except (ValueError, IndexError):
                continue
            bits = 0
            for name, val in zip(names, cols[2:6]):
                if val.lower().startswith("active"):
                    bits |= self.REASONS[name]
            self.samples.append((mhz, bits))

    def mark(self):
        """Index of the next sample: call at the start / end of the timed region."""
        return len(self.samples)

    def stop(self, first=0, last=None):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        if self._smi is not None:
            time.sleep(0.05)
            self._smi.terminate()
        window = self.samples[first:last] or self.samples
        if not window:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no clock samples"], "samples": 0}
        bits = 0
        for _, b in window:
            bits |= b
        return {"sm_mhz": float(np.median([m for m, _ in window])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(k for k, v in self.REASONS.items() if bits & v), "samples": len(window),
                "source": "nvml, 2 ms period, timed region only" if self._nvml else "nvidia-smi -lms 20"}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


# ------------------------------------------------------------------ CPU arms (oracle port)
def cpu_gather_push_steps(pos, mom, r, fields, steps):
    """The reference's path on the host: Interpolators gather (numpy.interp) of
    E_y at x, then BorisPush.push, whole-array numpy as the reference evaluates it."""
    from oracle import boris as oboris
    from oracle import gather as ogather
    zero3 = np.zeros(3)
    for s in range(steps):
        ey, _ = ogather.interp_b(pos[:, 0], r, fields[s % len(fields)])
        E = np.zeros_like(pos)
        E[:, 1] = ey
        oboris.push_aos(pos, mom, Q_E, M_E, E, zero3, DT)


def _cpu_worker(args):
    n, seed, ng, steps = args
    r, _ = grid_points(ng)
    fields = wave_fields(r, steps)
    pos, mom = host_ensemble(n, seed)
    t0 = time.perf_counter()
    cpu_gather_push_steps(pos, mom, r, fields, steps)
    return time.perf_counter() - t0


def cpu_baseline(ng, n, steps):
    r, _ = grid_points(ng)
    fields = wave_fields(r, steps + 1)
    pos, mom = host_ensemble(n, 99)
    cpu_gather_push_steps(pos, mom, r, fields, 1)            # warm-up
    t0 = time.perf_counter()
    cpu_gather_push_steps(pos, mom, r, fields, steps)
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} particles x {steps} steps, oracle numpy port of interp1d gather + BorisPush.push "
                      f"(single-threaded numpy, as the reference runs), {dt:.1f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation (oracle port; the
    reference itself is pure Python and is not present on the GPU box), all host
    cores: one forked worker per core on disjoint particle shards."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_worker = max(1 << 16, min(1 << 18, args.cpu_particles))
    steps, warm = max(1, args.steps), max(0, args.warmup)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        if warm:
            pool.map(_cpu_worker, [(per_worker, 7 + i, args.ng, 1) for i in range(cores)])
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(per_worker, 1000 + i, args.ng, steps) for i in range(cores)])
        dt = time.perf_counter() - t0
    value = per_worker * cores * steps / dt
    sample = (f"{cores} forked workers x {per_worker} particles x {steps} steps of the numpy port "
              f"(sample of the {args.particles}-particle-per-GPU workload)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "particles_per_gpu": args.particles, "grid_nodes": args.ng, "dt": DT,
                   "layout": "numpy (n,3) fp64 on the host", "parallelism": f"{cores} forked host workers",
                   "sample_particles_per_step": per_worker * cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ our arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from turbopy_b200 import build as tp_build
    if int(os.environ.get("LOCAL_RANK", "0")) == 0 and tp_build.needs_build():
        tp_build.build()                                           # in-tree .so missing or stale
    import turbopy_b200 as tp
    from turbopy_b200 import _lib
    from turbopy_b200.sharding import init_process_group

    rank, world, local = init_process_group()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; turbopy_b200 has no CPU fallback")
    dev = torch.device("cuda", local)
    n = args.particles
    r, dr = grid_points(args.ng)
    owner = Owner(r, dr)
    pusher = tp.BorisPush(owner, {"type": "BorisPush"})
    K, W = args.steps, max(3, args.warmup)
    nlev = K + W
    ngp = (args.ng + 1) // 2 * 2                                    # even row pitch: every row 16-B aligned (TMA)
    fields = torch.zeros((nlev, ngp), dtype=torch.float64, device=dev)
    fields[:, :args.ng] = torch.from_numpy(wave_fields(r, nlev)).to(dev)   # one E_y row per time level

    pos_h, mom_h = host_ensemble(n, 2345 + rank)
    pos, mom = tp.to_soa(pos_h, dev), tp.to_soa(mom_h, dev)

    def step(level):
        pusher.gather_push(pos, mom, Q_E, M_E, E_grid=(None, fields[level, :args.ng], None), B_grid=None,
                           axis=0, formula="interp")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for s in range(W):
        step(s)
    graph = None
    if args.graph:
        with tp.StepGraph(dev) as graph:
            for s in range(K):
                step(W + s)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.05)
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s_first = sampler.mark()
    ev0.record()
    if graph is not None:
        graph.replay()
    else:
        for s in range(K):
            step(W + s)
    ev1.record()
    barrier()
    s_last = sampler.mark()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop(s_first, s_last)
    pusher.check_status()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = n * world * K / (ms_max * 1e-3)

    # ---- roofline of the fused kernel (device time of K launches / K)
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_PUSH * n / (ms / K * 1e-3) / 1e9
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic["bytes_per_launch"] if traffic and traffic.get("particles") == n else None),
                "traffic_source": (traffic.get("source") if traffic and traffic.get("particles") == n else None),
                "peak_source": peak_src,
                "kernel": "gather_push_tma_kernel<interp(B), staged> (TMA-pipelined fused gather + Boris push)",
                "algorithmic_bytes_per_launch": BYTES_PER_PUSH * n,
                "timing": "CUDA events on the launch stream around K launches, / K"}

    # ---- end to end through the public API on HOST (pinned) arrays
    e2e = None
    if not args.no_e2e:
        pin_pos = torch.empty((n, 3), dtype=torch.float64).pin_memory()
        pin_mom = torch.empty((n, 3), dtype=torch.float64).pin_memory()
        pin_pos.copy_(torch.from_numpy(pos_h)); pin_mom.copy_(torch.from_numpy(mom_h))
        hp, hm = pin_pos.numpy(), pin_mom.numpy()
        fields_h = wave_fields(r, args.e2e_steps + 1)
        pusher.gather_push(hp, hm, Q_E, M_E, E_grid=(None, fields_h[0], None))          # warm-up
        barrier()
        t0 = time.perf_counter()
        for s in range(args.e2e_steps):
            pusher.gather_push(hp, hm, Q_E, M_E, E_grid=(None, fields_h[1 + s], None))
        barrier()
        dt_e2e = time.perf_counter() - t0
        te = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        import ctypes as C
        h2d, d2h = C.c_int64(), C.c_int64()
        _lib.lib().tp_host_last_transfer(C.byref(h2d), C.byref(d2h))
        e2e = {"value": n * world * args.e2e_steps / float(te.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d.value), "d2h_bytes_per_step": int(d2h.value),
               "steps": args.e2e_steps, "api": "BorisPush.gather_push(numpy pinned (n,3) arrays) -> "
                                              "tp_gather_push_host_f64"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.ng, args.cpu_particles, args.cpu_steps)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "particles_per_gpu": n, "grid_nodes": args.ng, "dt": DT, "layout": "SoA fp64 resident",
                       "parallelism": f"particle shards x{world}, grid replicated",
                       "l2": "inputs larger than L2 (1.6 GB of particle state per step vs 126 MB)",
                       "cuda_graph": bool(args.graph)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
