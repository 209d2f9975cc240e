#!/usr/bin/env python3
"""Benchmark of the pairs MD hot path on B200.

  python bench.py --gpus N --steps K --warmup W                 our CUDA path (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W   the reference's own CPU implementation (rank 0 only)

Metric (BASELINE.json): LJ atom-steps/s.  A "step" is one iteration of the generated timestep loop of examples/md.py
(initial_integrate -> exchange+borders+cell/neighbour build every 20th step | ghost refresh -> lennard_jones ->
final_integrate, thermo every 100).  Workload at N = 1: BASELINE.json configs[1] -- synthetic FCC lattice, 100^3 cells =
4,000,000 atoms, cutoff 2.5 sigma, skin 0.3, reneighbour every 20 steps, fp64.  N > 1: weak scaling, 4M atoms per GPU
(configs[3]), regular domain partitioner, NCCL halo exchange.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

RHO, TEMP, NTYPES = 0.8442, 1.44, 4
CUT, SKIN, DT, RENEIGH, THERMO = 2.5, 0.3, 0.005, 20, 100
METRIC, UNIT = "lj_atom_steps_per_s", "atom-steps/s"
# rank grids Regular6DStencil::setConfig picks when the global box is built from per-GPU cubes
WEAK_GRIDS = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}
_JSON_OUT = sys.stdout
REF_SAMPLE_NX = 63   # oracle/_ref variant md_bench: 4 * 63^3 = 1,000,188 atoms


def env_int(name, default):
    return int(os.environ.get(name, default))


def workload_config(n_gpus, nx, reference_arm=False):
    cfg = _workload_config(n_gpus, nx)
    if reference_arm:
        cfg["reference_sample"] = (f"the reference's serial C++ is timed on a {4 * REF_SAMPLE_NX ** 3}-atom cube (nx = {REF_SAMPLE_NX}) of the same "
                                   "lattice generator per replica process, not on the 4,000,000-atom box: its cost per atom-step does not "
                                   "depend on the box size, a 4 M-atom replica would take ~2 s per step and 6 GB per process")
    return cfg


def _workload_config(n_gpus, nx):
    return {"workload": f"lj_onetype-style synthetic FCC lattice, {4 * nx ** 3} atoms per GPU ({nx}^3 cells), rho 0.8442, "
                        f"cutoff 2.5 sigma + skin 0.3, reneighbour every {RENEIGH}, thermo every {THERMO}, fp64"
                        + (f"; weak scaling over {n_gpus} GPUs, regular partitioner, NCCL halo exchange" if n_gpus > 1 else ""),
            "atoms_per_gpu": 4 * nx ** 3, "n_gpus": n_gpus,
            "l2_policy": "inputs larger than L2 (neighbour lists ~1.3 GB + 0.5 GB particle state per step vs 126 MB L2)"}


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region.  NVML is polled every ~2 ms from a thread (a 20-step timed region
    lasts ~20 ms: `nvidia-smi -lms 50` cannot see it); if the NVML binding is missing, one long-lived nvidia-smi process streams rows
    every 50 ms instead.  The sampler starts before the warm-up; only rows taken while `recording` is set are kept."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []              # (sm MHz, max MHz, set of reasons)
        self.stop_flag = threading.Event()
        self.proc = None
        self.recording = False
        self.source = None
        self.ready = threading.Event()      # set once the first query has succeeded (NVML start-up takes ~100 ms), or on failure

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        self.source = "nvml"
        nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        self.ready.set()
        while not self.stop_flag.is_set():
            if self.recording:          # back to back while the timed region runs (a query takes ~1 ms; 20 steps last ~17 ms)
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((sm, mx, {k for k, b in bits.items() if r & b}))
                time.sleep(0)
            else:
                time.sleep(0.0005)

    def _run_smi(self):
        self.source = "nvidia-smi -lms 50"
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        for line in self.proc.stdout:
            self.ready.set()
            if self.stop_flag.is_set():
                break
            c = [x.strip() for x in line.split(",")]
            if self.recording and len(c) >= 7 and c[0].replace(".", "").isdigit():
                self.rows.append((float(c[0]), float(c[1]) if c[1].replace(".", "").isdigit() else None,
                                  {n for k, n in enumerate(self.NAMES) if c[3 + k].lower().startswith("active")}))

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            try:
                self._run_smi()
            except Exception:
                pass
        self.ready.set()

    def stop(self):
        self.stop_flag.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows if r[1]]
        reasons = sorted(set().union(*[r[2] for r in self.rows])) if self.rows else []
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows), "source": self.source}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the force kernel from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "lj_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------------------------------------
def reference_cuda_sample():
    """SECONDARY baseline (BASELINE.json north_star): the reference's OWN CUDA target -- `examples/md.py gpu` through the
    reference generator, nvcc -arch sm_100a with the reference's runtime (oracle/build_ref.py, variant md_1m_cuda) -- run as the
    stock executable on this GPU; numbers are the reference's own timers.  1,048,576 atoms is the largest cube its fixed
    capacities allow (at 4 M atoms it dies with an out-of-bounds write in determine_ghost_particles: 414 k ghosts against a
    send capacity of 200 000)."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "md_1m_cuda")
    if not os.path.exists(exe):
        return None
    try:
        r = subprocess.run([exe], cwd=os.path.dirname(exe), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    except Exception as e:      # noqa: BLE001
        return {"error": str(e)[:200]}
    t = {}
    for ln in r.stdout.splitlines():
        k, _, v = ln.partition(": ")
        try:
            t[k.strip()] = float(v.split()[0])
        except (ValueError, IndexError):
            pass
    if r.returncode != 0 or "all" not in t:
        return {"error": r.stdout[-300:]}
    n, iters = 4 * 64 ** 3, 201                  # timesteps=200 -> 201 loop iterations inside the `all` timer (sim/timestep.py:67-69)
    return {"value": n * iters / (t["all"] * 1e-3), "unit": UNIT, "atoms": n, "iterations": iters, "ms_per_step": t["all"] / iters,
            "lennard_jones_ms_per_call": t.get("lennard_jones", 0.0) / iters,
            "force_kernel_atoms_per_s": n * iters / (t["lennard_jones"] * 1e-3) if t.get("lennard_jones") else None,
            "what": "reference CUDA target (md.py gpu, generated md.cu + runtime/devices/cuda.cu, nvcc -O3 sm_100a, 1 rank), stock "
                    "executable, the reference's own timers; same lattice generator, reneighbour every 20"}


def ours_at(nx, steps, warmup):
    """Our path on the same 1 M-atom cube as the reference CUDA target, for a like-for-like ratio."""
    from pairs_b200 import backend
    ctx = backend.Context(0)
    a = pow(4.0 / RHO, 1.0 / 3.0)
    ctx.init_domain([0.0, nx * a, 0.0, nx * a, 0.0, nx * a])
    n = ctx.copper_fcc_lattice(nx, nx, nx, RHO, NTYPES)
    ctx.adjust_thermo(TEMP)
    ctx.set_lj_params(NTYPES, [1.0] * (NTYPES * NTYPES), [1.0] * (NTYPES * NTYPES))
    ctx.md_run(0, warmup, DT, CUT, CUT + SKIN, CUT + SKIN, RENEIGH, THERMO)
    ctx.timers_enable(True)
    ctx.timers_reset()
    ctx.stream_timer_start()
    ctx.md_run(warmup, warmup + steps, DT, CUT, CUT + SKIN, CUT + SKIN, RENEIGH, THERMO)
    ms = ctx.stream_timer_stop()
    lj_ms, lj_calls = ctx.timer("lennard_jones")
    return {"value": n * steps / (ms * 1e-3), "ms_per_step": ms / steps, "lennard_jones_ms_per_call": lj_ms / max(lj_calls, 1)}


def dem_sample():
    """Second workload of BASELINE.json (configs[2]): examples/dem.py scaled to 998,400 spheres + 2 half-spaces on one GPU,
    particle-steps/s while falling and after settling into contact (tools/bench_dem.py, run as its own process so that its
    contexts do not share state with the timed LJ run), next to the reference's serial C++ on one host core."""
    import subprocess
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_dem.py"), "8000"], cwd=ROOT, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True, timeout=240)
        if r.returncode != 0:
            return {"error": r.stderr[-300:]}
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:      # noqa: BLE001
        return {"error": str(e)[:200]}


def cpu_reference_sample(warmup, steps, replicas):
    """Times the reference's own generated serial C++ (oracle/_ref, kind 'reference') or, if that was not built, the
    restatement oracle/pairs_oracle.c (kind 'port') on a bounded 1,000,188-atom sample of the same workload."""
    from oracle import ref
    n_atoms = 4 * REF_SAMPLE_NX ** 3
    if ref.available("md_bench"):
        from oracle import ref_worker
        res = ref_worker.bench_many("md_bench", warmup, steps, replicas)
        value = sum(r["n"] * r["steps"] / r["seconds"] for r in res)
        secs = max(r["seconds"] for r in res)
        kind = "reference"
    else:
        from oracle import port
        sim = port.md_example(REF_SAMPLE_NX, reneigh_every=RENEIGH, particle_capacity=1400000, send_capacity=400000)
        for ts in range(warmup + 1):
            sim.step(ts)
        t0 = time.perf_counter()
        for ts in range(warmup + 1, warmup + 1 + steps):
            sim.step(ts)
        secs = time.perf_counter() - t0
        value = n_atoms * steps / secs
        kind, replicas = "port", 1
    sample = (f"{n_atoms} atoms (same lattice generator, nx={REF_SAMPLE_NX}), loop iterations {warmup + 1}..{warmup + steps} "
              f"timed after iteration 0 (set-up + first list build) and {warmup} warm-up iterations; reneighbour every {RENEIGH}; "
              f"serial target, g++ -O3 -ffp-contract=off; {replicas} independent replica process(es), aggregate throughput")
    out = {"value": value, "unit": UNIT, "cores": replicas, "kind": kind, "sample": sample, "seconds": secs}
    # the same program built with the reference Makefile's flags (-O3 -mavx2 -mfma, contraction allowed): the parity build above
    # forbids contraction, which is right for bit comparisons and slightly unkind for timing
    if kind == "reference" and ref.available("md_bench_fma"):
        from oracle import ref_worker
        res2 = ref_worker.bench_many("md_bench_fma", warmup, steps, replicas)
        out["value_o3_avx2_fma"] = sum(r["n"] * r["steps"] / r["seconds"] for r in res2)
    return out


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    replicas = args.ref_replicas or max(1, min(os.cpu_count() or 1, 64))
    try:
        avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2 ** 30
        replicas = max(1, min(replicas, int(avail_gb // 2.0)))     # ~1.5 GB per 1M-atom replica
    except (ValueError, OSError):
        pass
    t0 = time.perf_counter()
    cb = cpu_reference_sample(args.warmup, args.steps, replicas)
    wall = time.perf_counter() - t0
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * cb["seconds"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus, args.nx, reference_arm=True),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "value_o3_avx2_fma") if k in cb},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": wall}
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()
    return 0


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    from pairs_b200 import backend
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    dist = None
    if world > 1:
        import torch.distributed as dist   # plumbing only: rendezvous, barrier, max-over-ranks
        dist.init_process_group("gloo", rank=rank, world_size=world)
    nx = args.nx
    gx = WEAK_GRIDS.get(world)
    if gx is None:
        raise SystemExit(f"unsupported --gpus {world}")
    lattice = pow((4.0 / RHO), (1.0 / 3.0))
    cells = [nx * g for g in gx]
    grid = [0.0, cells[0] * lattice, 0.0, cells[1] * lattice, 0.0, cells[2] * lattice]
    assert backend.rank_grid(world, grid) == gx, (backend.rank_grid(world, grid), gx)

    # N > 1: the N-rank run against the SINGLE-RANK oracle of the same global system, before anything is timed (tests/mgpu_parity.py):
    # a mismatch raises on every rank -> non-zero exit, no JSON line
    parity_nranks = None
    if world > 1 and not args.no_parity:
        from tests import mgpu_parity
        t_par = time.perf_counter()
        parity_nranks = mgpu_parity.check(backend, dist, rank, world, local)
        parity_nranks["seconds"] = time.perf_counter() - t_par

    ctx = backend.Context(local)
    ctx.init_domain(grid, world_size=world, rank=rank)
    if world > 1:
        ids = [backend.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.nccl_init(ids[0])
    t_setup = time.perf_counter()
    nlocal0 = ctx.copper_fcc_lattice(cells[0], cells[1], cells[2], RHO, NTYPES)
    ctx.adjust_thermo(TEMP)
    ctx.set_lj_params(NTYPES, [1.0] * (NTYPES * NTYPES), [1.0] * (NTYPES * NTYPES))
    n_global = 4 * cells[0] * cells[1] * cells[2]
    t_setup = time.perf_counter() - t_setup

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    W, K = args.warmup, args.steps
    run = lambda a, b: ctx.md_run(a, b, DT, CUT, CUT + SKIN, CUT + SKIN, RENEIGH, THERMO)  # noqa: E731
    run_from_host = lambda x, v, m, t, a, b: ctx.md_run_from_host(x, v, m, t, a, b, DT, CUT, CUT + SKIN, CUT + SKIN, RENEIGH, THERMO)  # noqa: E731
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()           # before the warm-up, so that it is up when the timed region begins
        sampler.ready.wait(timeout=10)
    run(0, W)
    barrier()
    if rank == 0:
        sampler.recording = True
    ctx.timers_reset()
    ctx.timers_enable(True)
    launches0 = ctx.kernel_launches()
    barrier()
    t_wall = time.perf_counter()
    ctx.stream_timer_start()
    thermo = run(W, W + K)
    ms = ctx.stream_timer_stop()
    barrier()
    t_wall = time.perf_counter() - t_wall
    ctx.timers_enable(False)
    launches = ctx.kernel_launches() - launches0
    if rank == 0:
        sampler.recording = False
        sampler.stop()
        sampler.join(timeout=3)
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    value = n_global * K / (ms * 1e-3)

    # ---- roofline of the dominant kernel (lennard_jones), measured live over the timed region ----
    lj_ms, lj_calls = ctx.timer("lennard_jones")
    nl, ng = ctx.counts()
    nn = ctx.ints("numneighs")
    kbar = float(nn.mean()) if len(nn) else 0.0
    bytes_per_atom = 4.0 * kbar + 60.0 + 24.0 * ng / max(nl, 1)     # ids + numneighs + x_i + type + flags + force write + ghost x
    peak, peak_src = measured_peak()
    # one force evaluation per step; with comm / compute overlap (N > 1) it is TWO launches (interior tiles on the main stream,
    # boundary tiles on the comm stream): the per-step kernel time is their sum, whatever the number of recorded stages
    lj_ms_per_step = lj_ms / K
    achieved = (bytes_per_atom * nl / 1e9) / (lj_ms_per_step * 1e-3) if lj_ms > 0 else None
    stages = {}
    for name in ("lennard_jones", "initial_integrate", "final_integrate", "synchronize", "exchange", "borders", "build_cell_lists",
                 "build_neighbor_lists", "compute_thermo"):
        s_ms, s_calls = ctx.timer(name)
        stages[name] = {"ms": round(s_ms, 4), "calls": s_calls}
    # with and without the reneighbouring iterations: stage timers of the iterations that rebuild the lists
    ren_ms = sum(stages[k]["ms"] for k in ("exchange", "borders", "build_cell_lists", "build_neighbor_lists") if k in stages)
    ren_calls = stages.get("build_neighbor_lists", {}).get("calls", 0)
    reneighbor = {"every": RENEIGH, "rebuilds_in_timed_region": ren_calls, "ms_per_rebuild": ren_ms / ren_calls if ren_calls else None,
                  "ms_per_step_without_rebuilds": (ms - ren_ms) / K, "value_without_rebuilds": n_global * K / ((ms - ren_ms) * 1e-3),
                  "note": "rank 0 stage timers; rebuild = exchange + borders + cell lists + neighbour lists"}
    roofline = {"bound": "hbm", "kernel": "pb_k_tile_lj", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic(), "peak_source": peak_src,
                "algorithmic_bytes_per_atom": bytes_per_atom, "mean_neighbors": kbar, "atoms_per_step": nl,
                "launches_per_step": lj_calls / K, "ms_per_step_in_kernel": lj_ms_per_step,
                "force_kernel_atoms_per_s": nl / (lj_ms_per_step * 1e-3) if lj_ms > 0 else None,
                "share_of_step": lj_ms / ms if ms > 0 else None,
                "note": "numerator = SURVEY.md 8(d): 4 B per list entry + 60 B per atom + ghost positions; the tile lists actually "
                        "hold 2 B per entry (bytes_per_atom_as_stored), the staged tiles are re-read from L2, not HBM"
                        + ("; N > 1: the interior and the boundary tiles of a step are two launches on two streams that run "
                           "CONCURRENTLY (halo refresh overlapped), their event-timed durations are summed here -- an upper bound of the "
                           "kernel time per step, hence a lower bound of frac" if world > 1 else ""),
                "bytes_per_atom_as_stored": 2.0 * kbar + 60.0 + 24.0 * ng / max(nl, 1)}

    # ---- end to end through the C-ABI with HOST buffers: upload (pinned) -> K loop iterations from ts = 0 (first list build
    #      included) with thermo read-backs -> download of positions and velocities ----
    pos = ctx.real("position")
    vel = ctx.real("linear_velocity")
    mass = ctx.real("mass")
    typ = ctx.ints("type")
    cap_out = int(len(pos) * 1.25) + 4096          # nlocal changes under migration when N > 1
    out_pos, out_vel = np.empty((cap_out, 3)), np.empty((cap_out, 3))
    for a in (pos, vel, mass, typ, out_pos, out_vel):
        ctx.host_register(a)
    # three repetitions, the median counts: the GPU boxes share their host side (PCIe switch, memory bandwidth) with other jobs,
    # and one and the same 240 MB upload has been seen to take 11 ms and 500 ms in two consecutive processes
    e2e_samples = []
    for _rep in range(3):
        barrier()
        t0 = time.perf_counter()
        th2 = run_from_host(pos, vel, mass, typ, 0, K)
        n_after = ctx.counts()[0]
        assert n_after <= cap_out, (n_after, cap_out)
        ctx.real_into("position", out_pos)
        ctx.real_into("linear_velocity", out_vel)
        barrier()
        e2e_s = time.perf_counter() - t0
        if dist is not None:
            import torch
            t = torch.tensor([e2e_s], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t[0])
        e2e_samples.append(e2e_s)
    e2e_s = sorted(e2e_samples)[1]
    h2d = (pos.nbytes + vel.nbytes + mass.nbytes + typ.nbytes) * world
    d2h = (2 * n_after * 24 + th2.nbytes) * world
    e2e = {"value": n_global * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
           "what": "pb_md_run_from_host (pinned host arrays -> device, K iterations from ts=0 incl. the first neighbour build; at N = 1 "
                   "the copies of velocities and masses overlap that build) + thermo read-backs + pb_download_real(position, "
                   "linear_velocity); median of three repetitions",
           "seconds_per_repetition": e2e_samples}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_sample(1, 20, 1)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "value_o3_avx2_fma") if k in cb}
        cpu_baseline["host_cores_available"] = os.cpu_count()
    reference_cuda = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        reference_cuda = reference_cuda_sample()
        if reference_cuda is not None and "value" in reference_cuda:
            same = ours_at(64, 200, 20)
            reference_cuda["ours_same_size"] = same
            reference_cuda["speedup_same_size"] = same["value"] / reference_cuda["value"]

    dem = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.no_dem:
        dem = dem_sample()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "warmup": args.warmup, "config": workload_config(world, nx), "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "cpu_baseline": cpu_baseline, "reference_cuda": reference_cuda, "dem": dem, "reneighbor": reneighbor, "stages_ms": stages, "atoms_global": n_global,
                "parity_nranks": parity_nranks,
                "nlocal_rank0": nl, "nghost_rank0": ng, "wall_s_timed_region": t_wall, "setup_s": t_setup,
                "thermo_last": [float(x) for x in thermo[-1]] if len(thermo) else None}
        _JSON_OUT.write(json.dumps(line) + "\n")
        _JSON_OUT.flush()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# Second workload (BASELINE.json configs[2] / configs[4]): examples/dem.py's system -- simple-cubic grid of spheres with 10 % size
# scatter between two half-spaces, linear spring-dashpot contacts with tangential history, reneighbouring EVERY iteration -- on a
# 0.8 m x 0.8 m x 0.2 m box per GPU (998,400 spheres; RegularXY partitioner: 1x2x1, 2x2x1, 2x4x1 ranks).
DEM_GRIDS = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}
DEM_METRIC, DEM_UNIT = "dem_particle_steps_per_s", "particle-steps/s"


def dem_config(world, settle):
    gx, gy = DEM_GRIDS[world]
    return {"workload": f"examples/dem.py scaled to a {0.8 * gx:g} x {0.8 * gy:g} x 0.2 m box ({998400 * world} spheres + 2 half-spaces, 998,400 per GPU), "
                        f"linear spring-dashpot with tangential contact history, dt 5e-5, exchange + ghosts + cell lists every iteration; timed in the "
                        f"SETTLED bed (after {settle} iterations), fp64" + (f"; weak scaling over {world} GPUs, RegularXY partitioner ({gx}x{gy}x1 ranks), "
                        "particles migrate with their contact history" if world > 1 else ""),
            "spheres_per_gpu": 998400, "n_gpus": world,
            "l2_policy": "inputs larger than L2 (~1.6 KB of particle state + contact rows per sphere and iteration vs 126 MB L2)"}


def dem_setup(ctx, backend, dist, rank, world, domain):
    import math
    from tests import dem_common as dc
    ctx.init_domain([0.0, domain[0], 0.0, domain[1], 0.0, domain[2]], pbc=(1, 1, 0), partitioner=1, world_size=world, rank=rank)
    if world > 1:
        ids = [backend.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.nccl_init(ids[0])
    ctx.dem_enable(dc.C)
    ctx.dem_set_params(dc.DT, math.pi, dc.KAPPA, dc.LN_DRY, dc.COLLISION_TIME, dc.RHO_P, dc.RHO_F, dc.G, dc.NTYPES, dc.FS, dc.FD)
    ctx.setup_cells(dc.CELL)


def run_dem(args):
    import math
    from pairs_b200 import backend
    from tests import dem_common as dc
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
    gx, gy = DEM_GRIDS[world]
    domain = (0.8 * gx, 0.8 * gy, 0.2)
    # N > 1: the N-rank run (particles migrate WITH their contact history) against the single-GPU run of the same particles, before
    # anything is timed (tests/scripts/mgpu_dem_check.py: 840 spheres, 1300 iterations); a mismatch raises -> non-zero exit, no line
    parity_nranks = None
    if world > 1 and not args.no_parity:
        import importlib.util
        spec = importlib.util.spec_from_file_location("mgpu_dem_check", os.path.join(ROOT, "tests", "scripts", "mgpu_dem_check.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        t_par = time.perf_counter()
        parity_nranks = mod.check(backend, dist, rank, world, local)
        if rank == 0:
            parity_nranks["seconds"] = time.perf_counter() - t_par
    ctx = backend.Context(local)
    dem_setup(ctx, backend, dist, rank, world, domain)
    assert tuple(ctx.decomposition()["nranks"]) == (gx, gy, 1)
    g = ctx.dem_sc_grid(domain[0], domain[1], domain[2], dc.SPACING, dc.DIAMETER, dc.MIN_D, dc.MAX_D, dc.V0, dc.RHO_P, dc.NTYPES)
    ns = len(g["uid"])
    n = ns + 2
    pos, vel, normal = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
    mass, radius = np.ones(n), np.zeros(n)
    uid, typ, flags, shape = (np.zeros(n, np.int32) for _ in range(4))
    pos[:ns], vel[:ns], mass[:ns], radius[:ns], uid[:ns], typ[:ns] = g["position"], g["linear_velocity"], g["mass"], g["radius"], g["uid"], g["type"]
    for k, (u, p_, nrm) in enumerate([(100000000, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0)), (100000001, domain, (0.0, 0.0, -1.0))]):
        uid[ns + k], pos[ns + k], normal[ns + k], flags[ns + k], shape[ns + k] = u, p_, nrm, 13, 1
    ctx.reserve(int(1.25 * n) + 65536)
    ctx.upload(pos, vel, mass, typ, flags, uid, shape)
    ctx.dem_upload("radius", radius)
    ctx.dem_upload("normal", normal)
    ctx.dem_stage("update_mass_and_inertia")

    def allred(x, op="sum"):
        if dist is None:
            return x
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t[0])

    def barrier(c=None):
        (c or ctx).sync()
        if dist is not None:
            dist.barrier()

    n_global = int(allred(ns))
    W, K, settle = args.warmup, args.steps, args.dem_settle
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.ready.wait(timeout=10)
    # the falling phase (no contacts yet) is reported next to the headline window
    ctx.dem_run(dc.CELL, 0, 20)
    barrier()
    ctx.stream_timer_start()
    ctx.dem_run(dc.CELL, 20, 120)
    falling_ms = allred(ctx.stream_timer_stop(), "max") / 100
    ctx.dem_run(dc.CELL, 120, settle)             # settle into a bed with dense, sticking contacts
    ctx.dem_run(dc.CELL, settle, settle + W)
    barrier()
    if rank == 0:
        sampler.recording = True
    ctx.timers_reset()
    ctx.timers_enable(True)
    launches0 = ctx.kernel_launches()
    barrier()
    ctx.stream_timer_start()
    ctx.dem_run(dc.CELL, settle + W, settle + W + K)
    ms = allred(ctx.stream_timer_stop(), "max")
    barrier()
    ctx.timers_enable(False)
    launches = ctx.kernel_launches() - launches0
    if rank == 0:
        sampler.recording = False
        sampler.stop()
        sampler.join(timeout=3)
    value = n_global * K / (ms * 1e-3)
    stages = {k: {"ms": round(ctx.timer(k)[0], 4), "calls": ctx.timer(k)[1]} for k in
              ("exchange", "borders", "build_cell_lists", "gravity", "linear_spring_dashpot", "euler", "reset_contact_history_usage_status",
               "clear_unused_contact_history")}

    # ---- roofline of the contact kernels (pb_k_dem_detect + pb_k_dem_force = the module linear_spring_dashpot), SURVEY.md 8(d):
    #      950 B + 88 B per active contact + 4 B per particle in the 27 stencil cells, measured on THIS state ----
    nl, ng = ctx.counts()
    c = ctx.dem_download_contacts(nl)
    cbar = float(c["num_contacts"].mean())
    cs, _ = ctx.cell_lists()
    dcells, ncells, _ = ctx.cells()
    occ = np.diff(cs)[1:].reshape(tuple(int(x) for x in dcells)).astype(np.float64)       # cell 0 holds the INFINITE half-spaces
    box = np.zeros_like(occ)
    padded = np.pad(occ, 1)
    for a in range(3):
        for b in range(3):
            for d in range(3):
                box += padded[a:a + occ.shape[0], b:b + occ.shape[1], d:d + occ.shape[2]]
    p27 = float((box * occ).sum() / max(occ.sum(), 1.0))          # mean over particles of the particles in their 27 cells
    bytes_per_particle = 950.0 + 88.0 * cbar + 4.0 * p27
    lsd_ms = ctx.timer("linear_spring_dashpot")[0] / K
    peak, peak_src = measured_peak()
    achieved = bytes_per_particle * nl / 1e9 / (lsd_ms * 1e-3) if lsd_ms > 0 else None
    roofline = {"bound": "hbm", "kernel": "pb_k_dem_detect + pb_k_dem_force (module linear_spring_dashpot)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_particle": bytes_per_particle, "mean_contacts": cbar, "mean_particles_in_27_cells": p27,
                "particles_per_step": nl, "ms_per_step_in_kernels": lsd_ms, "share_of_step": lsd_ms * K / ms,
                "note": "SURVEY.md 8(d): 950 B own state / force / euler streams + 88 B per live contact + 4 B per particle of the 27 stencil cells"}

    # ---- end to end with HOST buffers: the settled state goes back to the host, a fresh context gets it through the C-ABI (particle
    #      arrays, DEM arrays, contact history), runs K iterations and returns positions and velocities ----
    names = ("radius", "angular_velocity", "normal", "inv_inertia", "rotation_matrix", "rotation_quat")
    host = {"position": ctx.real("position"), "linear_velocity": ctx.real("linear_velocity"), "mass": ctx.real("mass")}
    host_i = {k: ctx.ints(k) for k in ("type", "flags", "uid", "shape")}
    host_d = {k: ctx.dem_download(k, nl) for k in names}
    ctx.close()
    ctx2 = backend.Context(local)
    dem_setup(ctx2, backend, dist, rank, world, domain)
    ctx2.reserve(int(1.25 * nl) + 65536)
    out_pos, out_vel = np.empty((int(1.25 * nl) + 65536, 3)), np.empty((int(1.25 * nl) + 65536, 3))
    barrier(ctx2)
    t0 = time.perf_counter()
    ctx2.upload(host["position"], host["linear_velocity"], host["mass"], host_i["type"], host_i["flags"], host_i["uid"], host_i["shape"])
    for k in names:
        ctx2.dem_upload(k, host_d[k])
    ctx2.dem_upload_contacts(c["num_contacts"], c["contact_lists"], c["is_sticking"], c["tangential_spring_displacement"],
                             c["impact_velocity_magnitude"])
    ctx2.dem_run(dc.CELL, settle + W, settle + W + K)
    ctx2.real_into("position", out_pos)
    ctx2.real_into("linear_velocity", out_vel)
    ctx2.sync()
    e2e_s = allred(time.perf_counter() - t0, "max")
    h2d = sum(a.nbytes for a in host.values()) + sum(a.nbytes for a in host_i.values()) + sum(a.nbytes for a in host_d.values()) + \
        sum(c[k].nbytes for k in ("num_contacts", "contact_lists", "is_sticking", "tangential_spring_displacement", "impact_velocity_magnitude"))
    n_after = ctx2.counts()[0]
    e2e = {"value": n_global * K / e2e_s, "unit": DEM_UNIT, "h2d_bytes_per_step": allred(h2d) / K, "d2h_bytes_per_step": allred(2 * n_after * 24) / K,
           "what": "pb_upload_particles + pb_dem_upload_real x 6 + pb_dem_upload_contacts (host arrays of the settled state) + pb_dem_run over K "
                   "iterations + pb_download_real(position, linear_velocity)"}
    assert int(allred(n_after)) == n_global + 2 * world, "particles lost or duplicated"

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = dem_cpu_sample(1)
    if rank == 0:
        line = {"metric": DEM_METRIC, "value": value, "unit": DEM_UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dem_config(world, settle), "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": cpu_baseline, "falling_phase": {"ms_per_step": falling_ms, "value": n_global / (falling_ms * 1e-3)},
                "stages_ms": stages, "spheres_global": n_global, "nlocal_rank0": nl, "nghost_rank0": ng, "parity_nranks": parity_nranks}
        _JSON_OUT.write(json.dumps(line) + "\n")
        _JSON_OUT.flush()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def dem_cpu_sample(replicas):
    """The reference's generated serial C++ for examples/dem.py on the same 998,402-particle box (oracle/_ref variant dem_bench):
    loop iterations 3..10, i.e. the falling phase -- the settled bed is 8000 iterations = hours of CPU time away."""
    from oracle import ref, ref_worker
    if not ref.available("dem_bench"):
        return None
    res = ref_worker.bench_many("dem_bench", 2, 8, replicas)
    return {"value": sum(r["n"] * r["steps"] / r["seconds"] for r in res), "unit": DEM_UNIT, "cores": replicas, "kind": "reference",
            "sample": f"998,402 particles, loop iterations 3..10 (falling phase, no contacts yet: the cheapest iterations of the run), serial target, "
                      f"g++ -O3 -ffp-contract=off; {replicas} independent replica process(es), aggregate throughput",
            "seconds": max(r["seconds"] for r in res)}


def run_dem_reference(args):
    if env_int("RANK", 0) != 0:
        return 0
    replicas = args.ref_replicas or max(1, min(os.cpu_count() or 1, 32))
    try:
        avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2 ** 30
        replicas = max(1, min(replicas, int(avail_gb // 3.0)))
    except (ValueError, OSError):
        pass
    t0 = time.perf_counter()
    cb = dem_cpu_sample(replicas)
    if cb is None:
        _JSON_OUT.write(json.dumps({"impl": "reference", "unavailable": "oracle/_ref variant dem_bench not built"}) + "\n")
        return 0
    line = {"impl": "reference", "metric": DEM_METRIC, "value": cb["value"], "unit": DEM_UNIT, "n_gpus": args.gpus, "steps": 8, "warmup": 2,
            "ms_per_step": 1e3 * cb["seconds"] / 8, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dem_config(args.gpus, args.dem_settle),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": DEM_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "wall_s": time.perf_counter() - t0}
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()
    return 0


def main():
    # stdout carries exactly ONE JSON line: everything else a library prints there (e.g. NCCL's version banner) goes to stderr
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=100, help="FCC cells per GPU and dimension (100 -> 4M atoms: BASELINE configs[1])")
    ap.add_argument("--ref-replicas", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dem", action="store_true", help="skip the secondary DEM workload (tools/bench_dem.py)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the N-rank parity check against the single-rank oracle")
    ap.add_argument("--workload", default="lj", choices=["lj", "dem"], help="lj: BASELINE configs[1] / [3] (the headline); dem: configs[2] / [4]")
    ap.add_argument("--dem-settle", type=int, default=8000, help="DEM: untimed iterations before the timed window (SURVEY.md 8d: 8000)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_dem_reference(args) if args.workload == "dem" else run_reference(args)
    try:
        return run_dem(args) if args.workload == "dem" else run_ours(args)
    except BaseException:
        # a failing rank must not leave its peers blocked in a collective: report and hard-exit so the launcher tears down
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)


if __name__ == "__main__":
    sys.exit(main())
