"""Runs a reference program (oracle/_ref) in a FRESH process.  TEST INFRASTRUCTURE ONLY.

Why a fresh process: the reference allocates its property arrays with posix_memalign and never initialises the
ghost part (runtime/allocate.hpp:14-45; ghost `flags` are read but never transmitted), so
its results depend on the heap being zero pages -- true for its own executable, not inside a long-lived pytest
process.  MALLOC_MMAP_THRESHOLD_ forces those allocations to come from fresh zero-filled mmaps.

  python -m oracle.ref_worker dump  <variant> <out.npz> [nsteps] [name:width,...] [intname,...]   per-thermo-step snapshots (locals only;
                                                                    the list names further real properties to record)
  python -m oracle.ref_worker bench <variant> <warmup> <steps>     prints JSON {"n": atoms, "seconds": t, "steps": k}
"""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

ENV = {"MALLOC_MMAP_THRESHOLD_": "65536", "MALLOC_PERTURB_": "0"}


def spawn(args, **kw):
    env = dict(os.environ, **ENV)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    return subprocess.Popen([sys.executable, "-m", "oracle.ref_worker", *[str(a) for a in args]], cwd=root, env=env,
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, **kw)


BASE_PROPS = (("position", 3), ("linear_velocity", 3), ("force", 3), ("mass", 1))


def dump(variant, out_path, nsteps=None, extra=(), extra_int=()):
    """Snapshots of every thermo step of `variant`, produced in a fresh process; returns list of dicts.  `extra` = further
    (property name, width) pairs to record (user-defined properties of the variant)."""
    p = spawn(["dump", variant, out_path, nsteps if nsteps is not None else -1, ",".join(f"{n}:{w}" for n, w in extra), ",".join(extra_int)])
    out, err = p.communicate()
    if p.returncode != 0:
        raise RuntimeError(f"ref_worker dump failed:\n{out}\n{err}")
    z = np.load(out_path)
    n = int(z["count"])
    names = [k for k, _ in BASE_PROPS] + ["type"] + [k for k, _ in extra] + list(extra_int)
    snaps = [{k: z[f"{k}_{i}"] for k in names} | {"nlocal": int(z["nlocal"][i]), "nghost": int(z["nghost"][i])} for i in range(n)]
    for k in ("numneighs", "neighborlists"):
        if k in z.files:
            snaps[0][k] = z[k]
    return snaps


def dump_dem(variant, out_path, max_steps, keep):
    """DEM snapshots (fresh process); returns the loaded npz."""
    p = spawn(["dump_dem", variant, out_path, max_steps, ",".join(str(k) for k in keep)])
    out, err = p.communicate()
    if p.returncode != 0:
        raise RuntimeError(f"ref_worker dump_dem failed:\n{out[-3000:]}\n{err[-3000:]}")
    return np.load(out_path)


def dump_dem_end(variant, out_path, ts, names):
    """State at the end of iteration ts of a DEM program, the listed arrays only (fresh process); returns the loaded npz."""
    p = spawn(["dump_dem_end", variant, out_path, ts, ",".join(names)])
    out, err = p.communicate()
    if p.returncode != 0:
        raise RuntimeError(f"ref_worker dump_dem_end failed:\n{out[-3000:]}\n{err[-3000:]}")
    return np.load(out_path)


def bench_many(variant, warmup, steps, replicas):
    """`replicas` concurrent fresh processes, each timing `steps` loop iterations after `warmup`; list of results."""
    procs = [spawn(["bench", variant, warmup, steps]) for _ in range(replicas)]
    res = []
    for p in procs:
        out, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"ref_worker bench failed:\n{out[-2000:]}\n{err[-2000:]}")
        res.append(json.loads(out.strip().splitlines()[-1]))
    return res


def _main(argv):
    from oracle.ref import RefProgram
    mode, variant = argv[0], argv[1]
    prog = RefProgram(variant)
    if mode == "dump":
        out_path = argv[2]
        nsteps = int(argv[3]) if len(argv) > 3 and int(argv[3]) >= 0 else None
        extra = tuple((x.split(":")[0], int(x.split(":")[1])) for x in argv[4].split(",")) if len(argv) > 4 and argv[4] else ()
        extra_int = tuple(x for x in argv[5].split(",") if x) if len(argv) > 5 else ()
        # a program without Verlet lists (variant md_cells_t1) has no neighbour-list arrays to record
        snaps = prog.run_collect_thermo(props=BASE_PROPS + extra, int_props=("type", "flags") + extra_int, steps=nsteps,
                                        with_lists="cells" not in variant)
        d = {"count": len(snaps), "nlocal": np.array([s["nlocal"] for s in snaps]), "nghost": np.array([s["nghost"] for s in snaps])}
        for i, s in enumerate(snaps):
            for k in [x for x, _ in BASE_PROPS + extra] + ["type"] + list(extra_int):
                d[f"{k}_{i}"] = s[k]
        for k in ("numneighs", "neighborlists"):
            if snaps and k in snaps[0]:
                d[k] = snaps[0][k]
        np.savez(out_path, **d)
        return 0
    if mode == "dump_dem":
        # argv: dump_dem <variant> <out.npz> <max_steps> <step,step,...>  -- state at the END of the listed iterations, plus the
        # state at module boundaries inside them (after gravity = before the contact kernel, after the contact kernel, after euler)
        out_path, max_steps = argv[2], int(argv[3])
        keep = set(int(x) for x in argv[4].split(",")) if len(argv) > 4 and argv[4] else set()
        os.chdir(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref"))    # data/planes.input is opened relative to cwd
        st = {"ts": 0, "nlocal": None}
        d = {}

        def on_event(ev, a):
            ts = st["ts"]
            if ev == "thermo":
                st["nlocal"] = a
                if ts in keep:
                    for k, v in prog.dem_state(a).items():
                        d[f"end_{ts}_{k}"] = v
                d.setdefault("nlocal", []).append(a)
                nrecv = prog.array("nrecv", np.int32, 6)
                d.setdefault("nghost", []).append(int(nrecv.sum()))
                st["ts"] = ts + 1
            elif st["nlocal"] is not None and ts in keep and ev in ("module:gravity", "module:linear_spring_dashpot", "module:euler"):
                n = st["nlocal"] + int(prog.array("nrecv", np.int32, 6).sum())      # locals + ghosts
                tag = {"module:gravity": "pre", "module:linear_spring_dashpot": "post", "module:euler": "eul"}[ev]
                for k, v in prog.dem_state(n).items():
                    d[f"{tag}_{ts}_{k}"] = v
                d[f"{tag}_{ts}_particle_cell"] = prog.array("particle_cell", np.int32, n)

        prog.lib.ref_run_limited.argtypes = [ctypes.c_int, ctypes.c_int]
        import oracle.ref as _r
        cb = _r.HOOK(lambda ev, a, _u: on_event(ev.decode(), a))
        # bounded run with hooks: set the limit, then call ref_run (limit is consumed by ref_run)
        prog.lib.ref_set_limit.argtypes = [ctypes.c_int]
        prog.lib.ref_set_limit(max_steps)
        prog.lib.ref_run(cb, None, 1)
        d["nlocal"] = np.array(d["nlocal"])
        d["nghost"] = np.array(d["nghost"])
        np.savez_compressed(out_path, **d)
        return 0
    if mode == "dump_dem_end":
        # argv: dump_dem_end <variant> <out.npz> <ts> <name,name,...>  -- the state at the END of iteration ts, the listed arrays only
        # (the full-size cases: a complete dem_state of 10^6 particles is 1.2 GB)
        out_path, last = argv[2], int(argv[3])
        names = [x for x in argv[4].split(",") if x]
        os.chdir(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref"))
        st = {"ts": 0}
        d = {}
        widths = dict(RefProgram.DEM_REAL)

        def on_event(ev, a):
            if ev != "thermo":
                return
            if st["ts"] == last:
                n = a
                d["nlocal"] = np.array([n])
                d["nghost"] = np.array([int(prog.array("nrecv", np.int32, 6).sum())])
                for k in names:
                    if k in widths:
                        d[k] = prog.prop(k, n, widths[k])
                    elif k in RefProgram.DEM_INT:
                        d[k] = prog.prop(k, n, 1, np.int32)
                    elif k == "contact_lists":
                        d[k] = prog.array(k, np.int32, n * 20).reshape(n, 20)
                    elif k == "is_sticking":
                        d[k] = prog.contact_prop(k, n * 20, 1, np.int32).reshape(n, 20)
                    elif k.startswith("cp:"):          # any contact property: cp:<name>:<width>:<int|real>
                        _, cname, width, kind = k.split(":")
                        a = prog.contact_prop(cname, n * 20, int(width), np.int32 if kind == "int" else np.float64)
                        d[cname] = a.reshape(n, 20, int(width)) if int(width) > 1 else a.reshape(n, 20)
                    else:
                        d[k] = prog.array(k, np.int32, n)
            st["ts"] += 1

        import oracle.ref as _r
        cb = _r.HOOK(lambda ev, a, _u: on_event(ev.decode(), a))
        prog.lib.ref_set_limit.argtypes = [ctypes.c_int]
        prog.lib.ref_set_limit(last + 1)
        prog.lib.ref_run(cb, None, 1)
        np.savez(out_path, **d)
        return 0
    if mode == "bench":
        warmup, steps = int(argv[2]), int(argv[3])
        os.chdir(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref"))    # DEM programs open data/planes.input
        # redirect the program's own stdout chatter away from our JSON line
        sys.stdout.flush()
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)
        prog.lib.ref_run_limited.argtypes = [ctypes.c_int, ctypes.c_int]
        prog.lib.ref_run_limited(warmup + steps + 1, 1)      # iteration 0 (set-up + first build) is never timed
        os.dup2(saved, 1)
        buf = (ctypes.c_double * (warmup + steps + 8))()
        prog.lib.ref_thermo_times.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int]
        n = prog.lib.ref_thermo_times(buf, len(buf))
        t = list(buf)[:n]
        assert n >= warmup + steps + 1, (n, warmup, steps)
        seconds = t[warmup + steps] - t[warmup]
        atoms = {"md_bench": 4 * 63 ** 3, "md_bench_fma": 4 * 63 ** 3, "md_t1": 4 * 8 ** 3, "md_t2": 4 * 12 ** 3, "md": 4 * 32 ** 3, "dem_t1": 422, "dem_vtk_t1": 422, "dem_cn_t1": 422, "dem_nl_t1": 422, "dem_rn3_t1": 422, "md_half_t1": 4 * 8 ** 3, "md_custom_t1": 4 * 8 ** 3, "md_props_t1": 4 * 8 ** 3, "md_vocab_t1": 4 * 8 ** 3,
                 "dem_bench": 160 * 160 * 39 + 2, "dem_stock_t1": 160 * 3 * 39 + 2, "md_cells_t1": 4 * 8 ** 3}[variant]
        print(json.dumps({"n": atoms, "seconds": seconds, "steps": steps, "warmup": warmup}))
        return 0
    raise SystemExit(f"unknown mode {mode}")


if __name__ == "__main__":
    sys.exit(_main(sys.argv[1:]))
