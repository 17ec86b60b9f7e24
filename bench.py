#!/usr/bin/env python
"""Benchmark of the hot path: fused linear field gather + Boris push (fp64).

    python bench.py --gpus 1 --steps 400 --warmup 5                 # our arm, BASELINE configs[1]
    python bench.py --impl reference --gpus 1 --steps 3             # the reference's own CPU path
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N     # one rank per GPU: BASELINE configs[2]

Workloads (one "step" = one gather+push pass over every particle):

* N = 1 -> BASELINE.json configs[1] ("pif"): the particle-in-field example scaled up -- 16 Mi electrons in a 1-D EM
  plane wave E_y(r,t) = E0 cos(2 pi (k (r-0.5) - w t)) on a 4097-node grid, B = 0, linear gather at position[:,0]
  (Interpolators / numpy.interp rounding), relativistic Boris push.  The K timed steps are captured once in a CUDA
  graph (StepGraph) and replayed inside the timed region (--no-graph launches them from Python instead).
* N > 1 -> BASELINE.json configs[2] ("config3"): 125 M particles PER GPU (1 B at N = 8), a replicated
  2^20+1-node grid with six field components, fused gather+push every step, kinetic-energy diagnostic every step
  call whose reduction + NCCL all-reduce (8 doubles, on the compute stream) fires every 10th step -- all INSIDE
  the timed region.  Particles are cell-ordered once before the timed region (tp.sort_by_cell; dt is half the
  grid's CFL limit, so the order outlives the run); the random-order figure is reported beside it.

One JSON line on stdout (rank 0).  `value` = particle pushes per second over all ranks with state resident in HBM;
`e2e` = the same metric through BorisPush.gather_push on HOST (pinned numpy) arrays, H2D and D2H inside the timed
region, with the measured host-link ceiling beside it; `roofline` = 96 algorithmic bytes per particle-step over the
average kernel time against the measured HBM copy peak; `cpu_baseline` = the UNMODIFIED reference (staged under
oracle/_ref) timed on this box's host cores; `configs` = the other shapes of the same kernel (six components,
large grid, random vs cell-ordered) so that the headline does not hide them.
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

Q_E, M_E, C_LIGHT = -1.6022e-19, 9.1094e-31, 2.9979e8   # particle_in_field.py:49-50 magnitudes; BorisPush's c
OMEGA, C_WAVE, E0 = 2e8, 2.998e8, 1.0                    # particle_in_field.toml:15-17, particle_in_field.py:15
DT_PIF = 1e-8 / 20                                       # particle_in_field.toml:7-10
BYTES_PER_PUSH = 96                                      # 6 fp64 read + 6 fp64 written per particle-step
METRIC = "Boris particle-pushes/sec (fp64, gather+push)"
UNIT = "pushes/s"
WORKLOADS = {
    "pif": "configs[1]: pif electron ensemble in 1D EM plane wave, linear field gather (E_y on grid, B=0) + Boris push",
    "config3": "configs[2]: particles sharded over the GPUs, 2^20+1-node replicated grid, six field components, "
               "fused gather + Boris push, KE diagnostic + NCCL all-reduce every 10 steps",
}
NG_CONFIG3 = (1 << 20) + 1
KE_INTERVAL = 10
# host-side pause between the warm-up steps and the timed region, ms (A-B switch; round 1 slept 50 ms here while the
# NVML sampler started, which let the GPU go idle in front of a 5 ms timed region)
GAP_MS = float(os.environ.get("TURBOPY_B200_BENCH_GAP_MS", "0"))
# configs[2]: the KE diagnostic rides on the push kernel (KineticEnergyDiagnostic fuse_with_push; A-B switch)
FUSE_KE = os.environ.get("TURBOPY_B200_BENCH_FUSE_KE", "1") != "0"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "pif", "config3"])
    ap.add_argument("--particles", type=int, default=None, help="particles per GPU (default: 16 Mi pif, 125 M config3)")
    ap.add_argument("--ng", type=int, default=None, help="grid nodes (default: 4097 pif, 2^20+1 config3)")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--e2e-particles", type=int, default=16 << 20, help="particles per GPU of the end-to-end leg")
    ap.add_argument("--cpu-particles", type=int, default=1 << 20)
    ap.add_argument("--cpu-steps", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the `configs` block (other shapes of the kernel)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--graph", dest="graph", action="store_true", default=True,
                    help="configs[1]: replay the K timed steps as one CUDA graph (default; north_star: timesteps are "
                         "captured in a CUDA graph)")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch every timed step from Python")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.workload == "auto":
        a.workload = "pif" if max(world, a.gpus) == 1 else "config3"
    if a.particles is None:
        a.particles = (16 << 20) if a.workload == "pif" else 125_000_000
    if a.ng is None:
        a.ng = 4097 if a.workload == "pif" else NG_CONFIG3
    return a


# ------------------------------------------------------------------ workload
def grid_points(ng):
    """Grid.r and Grid.dr exactly as turbopy/core.py:625,666-667,709 form them on [0, 1]."""
    r = 0.0 + (1.0 - 0.0) * np.linspace(0, 1, ng)
    return r, (1.0 - 0.0) / (ng - 1)


def wave_fields(r, nlevels):
    """EMWave.update for time levels 0..nlevels-1 (particle_in_field.py:32-34)."""
    k = OMEGA / C_WAVE
    t = DT_PIF * np.arange(nlevels)[:, None]
    return E0 * np.cos(2 * np.pi * (-OMEGA * t + k * (r[None, :] - 0.5)))


def config3_fields(r):
    """Six smooth analytic components: E_k = 1e5 cos(2 pi (k+1) r) V/m, B_k = cos(2 pi (k+4) r) T."""
    return [1e5 * np.cos(2 * np.pi * (k + 1) * r) for k in range(3)] + [np.cos(2 * np.pi * (k + 4) * r) for k in range(3)]


def host_ensemble(n, seed, thermal=False):
    """pif-style ensemble: x uniform in [0.05, 0.95], y = z = 0, p = 0; config3: thermal momenta (0.1 mc)."""
    rng = np.random.default_rng(seed)
    pos = np.zeros((n, 3))
    pos[:, 0] = rng.uniform(0.05, 0.95, n)
    mom = np.zeros((n, 3))
    if thermal:
        pos[:, 1:] = rng.uniform(0.05, 0.95, (n, 2))
        mom[:] = M_E * C_LIGHT * rng.normal(0.0, 0.1, (n, 3))
    return pos, mom


class Owner:
    """The two attributes of a turboPy Simulation the pusher reads."""

    def __init__(self, r, dr, dt, num_steps=1000):
        from types import SimpleNamespace
        self.clock = SimpleNamespace(dt=dt, num_steps=num_steps)
        self.grid = SimpleNamespace(r=r, dr=dr, num_points=len(r), r_min=float(r[0]), r_max=float(r[-1]))


def dt_of(workload, dr):
    return DT_PIF if workload == "pif" else 0.5 * dr / C_LIGHT      # config3: at most half a cell per step


def config_of(args, world):
    """The workload both arms run (identical in the product and the reference line); how each arm executed it
    goes into its own `run` block."""
    n, ng = args.particles, args.ng
    cfg = {"workload": WORKLOADS[args.workload], "particles_per_gpu": n, "grid_nodes": ng,
           "dt": dt_of(args.workload, grid_points(ng)[1]),
           "parallelism": f"particle shards x{world}, grid replicated",
           "l2": f"inputs larger than L2 ({BYTES_PER_PUSH * n / 1e9:.1f} GB of particle state per step vs 126 MB)"}
    if args.workload == "config3":
        cfg["field_components"] = 6
        cfg["ke_diagnostic_interval_steps"] = KE_INTERVAL
    return cfg


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


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture of this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get(workload, t if workload == "pif" and "bytes_per_launch" in t else None)
    except Exception:
        return None


# ------------------------------------------------------------------ CPU arms: the reference's own code
def reference_tools(ng, dt):
    """(kind, step function) -- the UNMODIFIED reference when it is present (/root/reference in the build
    container, the staged oracle/_ref on the GPU box): stock ``Interpolators.interpolate1D`` (scipy interp1d ->
    numpy.interp, turbopy/computetools.py:459-481) for the gather and stock ``BorisPush.push`` (:407-441) on numpy
    (n,3) arrays.  Otherwise the oracle's numpy port of the same two functions (kind "port")."""
    from oracle import refload
    if refload.available():
        turbopy = refload.load()
        from turbopy import computetools as stock           # the stock classes themselves, whatever the registry holds
        sim = refload.make_owner(turbopy, grid={"N": ng, "r_min": 0.0, "r_max": 1.0},
                                 clock={"start_time": 0.0, "end_time": dt * 1000, "num_steps": 1000})
        sim.clock.dt = dt
        push = stock.BorisPush(sim, {"type": "BorisPush"}).push
        interp = stock.Interpolators(sim, {"type": "Interpolators"}).interpolate1D
        r = sim.grid.r

        def step(pos, mom, comps):
            F = [np.zeros(len(pos)) if c is None else interp(r, c)(pos[:, 0]) for c in comps]
            push(pos, mom, Q_E, M_E, np.stack(F[:3], axis=1), np.stack(F[3:], axis=1))
        return "reference", step, r
    from oracle import boris as oboris, gather as ogather
    r, _ = grid_points(ng)

    def step(pos, mom, comps):
        F = [np.zeros(len(pos)) if c is None else ogather.interp_b(pos[:, 0], r, c)[0] for c in comps]
        oboris.push_aos(pos, mom, Q_E, M_E, np.stack(F[:3], axis=1), np.stack(F[3:], axis=1), dt)
    return "port", step, r


def cpu_fields(workload, r, nlevels):
    """Per time level the six grid components (None = absent) of the workload."""
    if workload == "pif":
        return [[None, ey, None, None, None, None] for ey in wave_fields(r, nlevels)]
    return [config3_fields(r)] * nlevels


def _cpu_worker(args):
    workload, n, seed, ng, steps = args
    _, dr = grid_points(ng)
    kind, step, r = reference_tools(ng, dt_of(workload, dr))
    fields = cpu_fields(workload, r, steps)
    pos, mom = host_ensemble(n, seed, thermal=workload != "pif")
    t0 = time.perf_counter()
    for s in range(steps):
        step(pos, mom, fields[s])
    return time.perf_counter() - t0, kind


def cpu_baseline(workload, ng, n, steps):
    """One core (numpy's elementwise kernels are single-threaded, as the reference runs), bounded sample."""
    _cpu_worker((workload, min(n, 1 << 16), 7, ng, 1))             # warm-up (imports, page faults)
    dt, kind = _cpu_worker((workload, n, 99, ng, steps))
    return {"value": n * steps / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"{n} particles x {steps} steps of the {'unmodified reference' if kind == 'reference' else 'oracle numpy port'}: "
                      f"Interpolators.interpolate1D gather + BorisPush.push on numpy (n,3) arrays, one process, {dt:.1f} s"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, all host cores -- one forked worker
    per core on disjoint particle shards (particles are independent; numpy itself is single-threaded)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_worker = max(1 << 15, min(1 << 18, args.cpu_particles))
    if args.workload == "config3":
        # six interp1d objects over a 1 Mi-node grid are built per step (that is how Interpolators.interpolate1D
        # works, computetools.py:480); a larger shard per worker keeps that O(ng) part from dominating the sample
        per_worker = 1 << 19
    steps, warm = max(1, args.steps), max(0, args.warmup)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        if warm:
            pool.map(_cpu_worker, [(args.workload, per_worker, 7 + i, args.ng, 1) for i in range(cores)])
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(args.workload, per_worker, 1000 + i, args.ng, steps) for i in range(cores)])
        dt = time.perf_counter() - t0
    kind = res[0][1]
    value = per_worker * cores * steps / dt
    sample = (f"{cores} forked workers x {per_worker} particles x {steps} steps of the "
              f"{'unmodified reference (oracle/_ref or /root/reference)' if kind == 'reference' else 'oracle numpy port'} "
              f"(a sample of the {args.particles}-particle-per-GPU workload)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args, args.gpus),
        "run": {"layout": "numpy (n,3) fp64 on the host", "host_workers": cores,
                "sample_particles_per_step": per_worker * cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ our arm
class Bench:
    """Per-rank state of the product arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from turbopy_b200 import build as tp_build
        if int(os.environ.get("LOCAL_RANK", "0")) == 0 and tp_build.needs_build():
            tp_build.build()                                           # in-tree .so missing or stale
        from turbopy_b200.sharding import init_process_group
        self.rank, self.world, self.local = init_process_group()
        if self.world > 1:
            dist.barrier()                                             # the other ranks wait for the build
        import turbopy_b200 as tp
        from turbopy_b200 import _lib
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; turbopy_b200 has no CPU fallback")
        self.torch, self.dist, self.tp, self._lib, self.args = torch, dist, tp, _lib, args
        self.dev = torch.device("cuda", self.local)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, sampler=None):
        """K calls of fn(step) bracketed by barrier + synchronize; device time by CUDA events on the launch stream."""
        torch = self.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        first = sampler.mark() if sampler else 0
        launches0 = self._lib.launch_count()
        ev0.record()
        fn(steps)
        ev1.record()
        self.barrier()
        last = sampler.mark() if sampler else 0
        return ev0.elapsed_time(ev1), self._lib.launch_count() - launches0, (first, last)

    # -------------------------------------------------------------- configs[1]
    def setup_pif(self, n, ng, nlevels):
        torch, tp = self.torch, self.tp
        r, dr = grid_points(ng)
        pusher = tp.BorisPush(Owner(r, dr, DT_PIF), {"type": "BorisPush"})
        ngp = (ng + 1) // 2 * 2                                        # even row pitch: every row 16-B aligned (TMA)
        fields = torch.zeros((nlevels, ngp), dtype=torch.float64, device=self.dev)
        fields[:, :ng] = torch.from_numpy(wave_fields(r, nlevels)).to(self.dev)     # one E_y row per time level
        pos_h, mom_h = host_ensemble(n, 2345 + self.rank)
        pos, mom = tp.to_soa(pos_h, self.dev), tp.to_soa(mom_h, self.dev)

        def step(level):
            pusher.gather_push(pos, mom, Q_E, M_E, E_grid=(None, fields[level % nlevels, :ng], None), B_grid=None,
                               axis=0, formula="interp")
        return pusher, step

    # -------------------------------------------------------------- configs[2] (and the 6-component shapes)
    def setup_config3(self, n, ng, order, diag_interval=KE_INTERVAL):
        torch, tp = self.torch, self.tp
        r, dr = grid_points(ng)
        owner = Owner(r, dr, dt_of("config3", dr), num_steps=1 << 20)
        pusher = tp.BorisPush(owner, {"type": "BorisPush"})
        gen = torch.Generator(device=self.dev).manual_seed(3456 + self.rank)        # per-rank seed
        pos, mom = tp.soa_particles(n, self.dev), tp.soa_particles(n, self.dev)
        for c in range(3):
            pos[:, c].copy_(0.05 + 0.9 * torch.rand(n, dtype=torch.float64, device=self.dev, generator=gen))
            mom[:, c].copy_(torch.randn(n, dtype=torch.float64, device=self.dev, generator=gen) * (M_E * C_LIGHT * 0.1))
        comps = config3_fields(r)
        E = torch.from_numpy(np.stack(comps[:3], axis=1)).to(self.dev)
        B = torch.from_numpy(np.stack(comps[3:], axis=1)).to(self.dev)
        sort_ms = None
        if order == "sorted":
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            tp.sort_by_cell(pos, mom, owner.grid)
            e1.record()
            torch.cuda.synchronize(self.dev)
            sort_ms = e0.elapsed_time(e1)
        diag = None
        if diag_interval:
            diag = tp.KineticEnergyDiagnostic(owner, {"momentum": "p", "mass": M_E, "interval": diag_interval,
                                                      "fuse_with_push": FUSE_KE})
            diag.inspect_resource({"p": mom})
            diag.initialize()

        def step(_level):
            if diag is not None:
                diag.diagnose()                 # every 10th call: reduction + all-reduce of 8 doubles on this stream
            pusher.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B)
        return {"pusher": pusher, "step": step, "diag": diag, "pos": pos, "mom": mom, "E": E, "B": B, "owner": owner,
                "comps": comps, "sort_ms": sort_ms, "r": r}

    def measure(self, step, steps, warmup):
        for s in range(warmup):
            step(s)
        ms, launches, _ = self.timed(lambda k: [step(warmup + s) for s in range(k)], steps)
        return self.max_over_ranks(ms), ms, launches

    # -------------------------------------------------------------- parity of the multi-GPU path (checker, untimed)
    def parity_config3(self, st):
        """(a) the all-reduced KE row against the fp64 sum of the per-rank partial rows (<= 1e-12);
        (b) a 16 384-particle subset of this rank's shard, 3 more steps, bit for bit against the oracle (particles
        are independent, so a subset is an exact oracle).  Runs AFTER the timed region; the oracle is the checker."""
        torch, tp = self.torch, self.tp
        from oracle import boris as oboris, gather as ogather          # checker only
        mom, pos = st["mom"], st["pos"]
        part = tp.reduce_moments(mom, M_E, st["pusher"].c2).clone()
        red = part.clone()
        from turbopy_b200.sharding import allreduce_sum_
        allreduce_sum_(red)
        parts = [torch.zeros_like(part) for _ in range(self.world)]
        if self.world > 1:
            self.dist.all_gather(parts, part)
        else:
            parts = [part]
        want = np.sum(np.stack([p.cpu().numpy() for p in parts]), axis=0)
        got = red.cpu().numpy()
        ke_rel = float(abs(got[0] - want[0]) / abs(want[0]))
        m = min(16_384, pos.shape[0])
        sp, sm = pos[:m].cpu().numpy().copy(), mom[:m].cpu().numpy().copy()
        # the diagnostic riding on the push (fuse_with_push) against the standalone reduction of the same momenta
        fd = tp.KineticEnergyDiagnostic(st["owner"], {"momentum": "p", "mass": M_E, "allreduce": False,
                                                      "fuse_with_push": True, "interval": 1 << 20})
        fd.inspect_resource({"p": mom})
        fd.initialize()
        for k in range(3):
            if k == 0:
                alone = tp.reduce_moments(mom, M_E, st["pusher"].c2).cpu().numpy()
                fd.diagnose()
            st["pusher"].gather_push(pos, mom, Q_E, M_E, E_grid=st["E"], B_grid=st["B"])
        fused = fd.values()[0]
        fused_rel = float(max(abs(fused[0] - alone[0]) / abs(alone[0]), abs(fused[4] - alone[4]) / abs(alone[4])))
        fused_ok = fused_rel <= 1e-12 and fused[5] == alone[5] and fd.fused_rows == 1
        st["pusher"].check_status()
        r, dt = st["r"], st["owner"].clock.dt
        for _ in range(3):
            F = [ogather.interp_b(sp[:, 0], r, c)[0] for c in st["comps"]]
            oboris.push_aos(sp, sm, Q_E, M_E, np.stack(F[:3], axis=1), np.stack(F[3:], axis=1), dt)
        gp, gm = pos[:m].cpu().numpy(), mom[:m].cpu().numpy()
        exact = bool(np.array_equal(gp.view(np.uint64), sp.view(np.uint64)) and
                     np.array_equal(gm.view(np.uint64), sm.view(np.uint64)))
        ok = self.max_over_ranks(0.0 if (exact and ke_rel <= 1e-12 and fused_ok) else 1.0) == 0.0
        return {"ok": ok, "ke_allreduce_vs_sum_of_partials_rel": ke_rel, "ke_tolerance": 1e-12,
                "ke_fused_in_push_vs_standalone_rel": fused_rel,
                "count_all_ranks": float(got[5]), "subset_particles_per_rank": m, "subset_steps": 3,
                "subset_bit_identical_to_oracle": exact, "ranks": self.world}

    # -------------------------------------------------------------- end to end on host arrays
    def e2e(self, workload, ng):
        torch, tp, args = self.torch, self.tp, self.args
        import ctypes as C
        n = args.e2e_particles
        r, dr = grid_points(ng)
        pusher = tp.BorisPush(Owner(r, dr, dt_of(workload, dr)), {"type": "BorisPush"})
        affinity = local_affinity(self.local)
        pos_h, mom_h = host_ensemble(n, 2345 + self.rank, thermal=workload != "pif")
        pin_pos = torch.empty((n, 3), dtype=torch.float64).pin_memory()
        pin_mom = torch.empty((n, 3), dtype=torch.float64).pin_memory()
        pin_pos.copy_(torch.from_numpy(pos_h)); pin_mom.copy_(torch.from_numpy(mom_h))
        hp, hm = pin_pos.numpy(), pin_mom.numpy()
        if workload == "pif":
            fields_h = wave_fields(r, args.e2e_steps + 1)
            grids = [dict(E_grid=(None, f, None), B_grid=None) for f in fields_h]
        else:
            comps = config3_fields(r)
            grids = [dict(E_grid=np.stack(comps[:3], axis=1), B_grid=np.stack(comps[3:], axis=1))] * (args.e2e_steps + 1)
        for _ in range(2):      # warm-up; large host grid arrays are page-locked the second time the library sees them
            pusher.gather_push(hp, hm, Q_E, M_E, **grids[0])
        self.barrier()
        t0 = time.perf_counter()
        for s in range(args.e2e_steps):
            pusher.gather_push(hp, hm, Q_E, M_E, **grids[1 + s])
        self.barrier()
        dt_e2e = self.max_over_ranks(time.perf_counter() - t0)
        h2d, d2h = C.c_int64(), C.c_int64()
        self._lib.lib().tp_host_last_transfer(C.byref(h2d), C.byref(d2h))
        value = n * self.world * args.e2e_steps / dt_e2e
        # the ceiling: what this box's host link gives every GPU when all ranks copy at once (both directions)
        up, down, both = C.c_double(), C.c_double(), C.c_double()
        self.barrier()
        rc = self._lib.lib().tp_host_link_probe(256 << 20, 4, C.byref(up), C.byref(down), C.byref(both))
        self.barrier()
        link = None
        if rc == 0:
            duplex_min = -self.max_over_ranks(-both.value)              # the slowest rank bounds the job
            ceiling = self.world * duplex_min * 1e9 / (BYTES_PER_PUSH / 2)
            link = {"h2d_gbs": up.value, "d2h_gbs": down.value, "duplex_gbs_per_direction": both.value,
                    "duplex_gbs_per_direction_min_over_ranks": duplex_min, "ranks_copying_at_once": self.world,
                    "ceiling_pushes_per_s": ceiling, "frac_of_ceiling": value / ceiling,
                    "how": "tp_host_link_probe: pinned cudaMemcpyAsync, 256 MiB x 4 per direction on two streams, "
                           "all ranks at once (rank 0's figures; min over ranks for the ceiling); "
                           "ceiling = ranks x duplex GB/s / 48 B per push per direction"}
        return {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(h2d.value), "d2h_bytes_per_step": int(d2h.value),
                "steps": args.e2e_steps, "particles_per_gpu": n, "workload": WORKLOADS[workload].split(":")[0],
                "api": "BorisPush.gather_push(numpy pinned (n,3) arrays) -> tp_gather_push_host_f64",
                "host_link": link, "cpu_affinity": affinity}


def local_affinity(local_rank):
    """Bind this rank to the host cores next to its GPU before it allocates pinned memory (first touch decides the
    NUMA node of the staging pages).  Only narrows within the cpuset the container allows; reports what it did."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[local_rank]) if visible and visible.split(",")[local_rank].isdigit() else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        near = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        pick = near & allowed
        if pick and pick != allowed:
            os.sched_setaffinity(0, pick)
        return {"gpu_local_cpus": len(near), "allowed_cpus": len(allowed), "bound_to": len(pick) if pick else len(allowed),
                "narrowed": bool(pick and pick != allowed)}
    except Exception as exc:                                            # best effort
        return {"error": type(exc).__name__}


def run_b200(args):
    b = Bench(args)
    torch, tp = b.torch, b.tp
    K, W = args.steps, max(3, args.warmup)
    n, ng, workload = args.particles, args.ng, args.workload
    peak, peak_src = measured_peak()
    extras = {}
    parity = None

    sampler = ClockSampler(b.local)
    if workload == "pif":
        pusher, step = b.setup_pif(n, ng, K + W)
        sampler.start()                          # NVML thread up before the warm-up: no idle gap in front of the timed steps
        graph = None
        if args.graph:
            step(0)
            with tp.StepGraph(b.dev) as graph:
                for s in range(K):
                    step(W + s)
        for s in range(W):
            step(s)
        time.sleep(GAP_MS * 1e-3)
        ms, launches, window = b.timed((lambda k: graph.replay()) if graph is not None else
                                       (lambda k: [step(W + s) for s in range(k)]), K, sampler)
        pusher.check_status()
        kernel = ("gather_push_tma_kernel<interp(B), staged, SoA, PMASK=0x02> (TMA-pipelined fused gather + Boris push; "
                  "the instance specialised for E_y-only grids, consumers paced 0.8 us per tile)")
        layout_note = "SoA fp64 resident"
        order_note = None
    else:
        st = b.setup_config3(n, ng, "sorted")
        step, pusher = st["step"], st["pusher"]
        sampler.start()
        for s in range(W):
            step(s)
        time.sleep(GAP_MS * 1e-3)
        ms, launches, window = b.timed(lambda k: [step(W + s) for s in range(k)], K, sampler)
        pusher.check_status()
        kernel = ("gather_push_win_kernel<interp(B)> (TMA particle ring + per-tile windows of cell records; "
                  "pack_cells_kernel before it; every 10th step its <MOM> instance also accumulates the KE / moment "
                  "sums, followed by moments_final_kernel + ncclAllReduce)")
        layout_note = "SoA fp64 resident, cell-ordered once before the timed region (tp.sort_by_cell)"
        order_note = {"sort_ms_per_gpu": st["sort_ms"], "sorted_before_timed_region": True,
                      "steps_since_sort_in_timed_region": [W, W + K],
                      "order_lifetime": "dt = dr/(2c), thermal 0.1c momenta: a 960-particle tile spans ~10 cells "
                                        "right after the sort and ~0.3 cells more per step; the per-tile windows "
                                        "hold 65 cells.  Measured decay at 125 M particles "
                                        "(profiles/r2_order_decay.md): 92 % of the HBM peak in the first steps, "
                                        "83 % at step 100, 78 % at 200, 66-68 % from step 300 on (L2 path); a "
                                        "re-sort every ~200 steps keeps the average incl. the sort at ~74 %"}
        if not args.no_parity:
            parity = b.parity_config3(st)
        ke_rows = st["diag"].values()
        extras["kinetic_energy_first_last_J"] = [float(ke_rows[0][0]), float(ke_rows[-1][0])] if len(ke_rows) else None
        extras["ke_allreduces_in_timed_region"] = (K + KE_INTERVAL - 1) // KE_INTERVAL
        del st
        torch.cuda.empty_cache()
    clocks = sampler.stop(*window)
    ms_max = b.max_over_ranks(ms)
    value = n * b.world * K / (ms_max * 1e-3)
    achieved = BYTES_PER_PUSH * n / (ms / K * 1e-3) / 1e9
    traffic = ncu_traffic(workload)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic["bytes_per_launch"] if traffic and traffic.get("particles") == n else None),
                "traffic_source": (traffic.get("source") if traffic and traffic.get("particles") == n else None),
                "peak_source": peak_src, "kernel": kernel,
                "algorithmic_bytes_per_launch": BYTES_PER_PUSH * n,
                "timing": "CUDA events on the launch stream around K steps, / K (this rank); `value` uses the max over ranks"}

    # ---- the other shapes of the same kernel (short runs; not the headline)
    if not args.no_extras:
        def shape(label, st, steps=20):
            mx, _, _ = b.measure(st["step"], steps, 3)
            st["pusher"].check_status()
            nn = st["pos"].shape[0]
            extras[label] = {"value": nn * b.world * steps / (mx * 1e-3), "unit": UNIT, "ms_per_step": mx / steps,
                             "frac_of_hbm_peak_per_gpu": BYTES_PER_PUSH * nn / (mx / steps * 1e-3) / 1e9 / peak,
                             "particles_per_gpu": nn, "sort_ms_per_gpu": st["sort_ms"]}
        if workload == "pif":
            for order in ("random", "sorted"):
                st = b.setup_config3(16 << 20, 1025, order, diag_interval=0)
                shape(f"six_components_ng1025_{order}", st)
                del st
            for order in ("sorted", "random"):
                st = b.setup_config3(125_000_000, NG_CONFIG3, order)
                shape(f"config3_per_gpu_{order}", st)
                del st
                torch.cuda.empty_cache()
        else:
            st = b.setup_config3(n, ng, "random")
            shape("config3_random_order", st)
            del st
            torch.cuda.empty_cache()
        # the same shape over 500 steps with the tool keeping the order itself (input_data sort_every = 100: one
        # global sort, then a block-local re-sort every 100 steps, all inside the timed region) -- the sustained figure
        st = b.setup_config3(125_000_000 if workload == "pif" else n, NG_CONFIG3, "random")
        st["pusher"].sort_every = 100
        shape("config3_sustained_500_steps_sort_every_100", st, steps=500)
        del st
        torch.cuda.empty_cache()
        if workload != "pif":
            pusher1, step1 = b.setup_pif(16 << 20, 4097, 64)
            mx, _, _ = b.measure(step1, 100, 5)
            extras["pif_16Mi_per_gpu"] = {"value": (16 << 20) * b.world * 100 / (mx * 1e-3), "unit": UNIT,
                                          "ms_per_step": mx / 100,
                                          "note": "BASELINE configs[1] per GPU, no collective (the N = 1 headline workload)"}

    e2e = None if args.no_e2e else b.e2e(workload, ng)
    cpu = None
    if b.rank == 0 and b.world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(workload, ng, args.cpu_particles, args.cpu_steps)

    if b.rank == 0:
        config = config_of(args, b.world)
        run = {"layout": layout_note, "cuda_graph": bool(args.graph and workload == "pif")}
        if run["cuda_graph"]:
            run["cuda_graph_note"] = (f"the {K} timed steps are one captured graph of {K} gather_push launches "
                                      "(StepGraph -> tp_graph_*), replayed once inside the timed region; the warm-up "
                                      "steps are launched eagerly.  Eager launches measure 0.6 % lower "
                                      "(profiles/r2_final_ab.txt)")
        if order_note:
            run["particle_order"] = order_note
            run["collective"] = ("tp_nccl_allreduce_sum_f64 (own NCCL communicator) of 8 doubles on the compute stream, "
                                 "every 10th step")
            run["diagnostic"] = ("KineticEnergyDiagnostic, fuse_with_push: on its steps the sums are accumulated inside "
                                 "gather_push_win_kernel<MOM> from the momenta it has just loaded "
                                 "(tp_gather_push_moments_f64)" if FUSE_KE else
                                 "KineticEnergyDiagnostic: standalone reduction (moments_partial_kernel) in front of the push")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": b.world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "run": run,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "configs": extras, "parity": parity,
        }
        if workload == "config3":
            line["scaling_note"] = ("N = 1 runs configs[1] (one component, 4097-node grid, no collective); this line is "
                                    "configs[2].  The single-GPU figure of THIS workload is `configs.config3_per_gpu_sorted` "
                                    "of the N = 1 line; `configs.pif_16Mi_per_gpu` here is configs[1] on all N GPUs.")
            line["value_per_gpu"] = value / b.world
        print(json.dumps(line), flush=True)
    if b.world > 1:
        from turbopy_b200.sharding import DiagnosticComm
        if DiagnosticComm._instance is not None:
            DiagnosticComm._instance.close()
        b.dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
