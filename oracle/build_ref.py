#!/usr/bin/env python3
"""Build recipe for oracle/_ref/: the reference's OWN serial CPU implementation of the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Nothing here is copied from the reference:
the reference's generator (/root/reference/src/pairs) is *executed* on the reference's own
example scripts (/root/reference/examples/md.py, dem.py -- optionally with a few scalar
parameters substituted in a scratch copy, e.g. lattice size or number of timesteps) and the
C++ it prints is compiled, where it lies in oracle/_ref/gen/, against the reference runtime
sources where they lie in /root/reference/runtime, plus the single-rank MPI stand-in
oracle/mpi_stub/mpi.h (the image has no MPI).  Outputs go only into oracle/_ref/ (git-ignored,
NOT gpurun-ignored, so the built .so / binaries travel to the GPU box).

Variants (name -> substitutions applied to examples/md.py in a scratch copy):
  md        stock examples/md.py (config C1: 32^3 cells = 131072 atoms, 200 steps)
  md_t1     nx=ny=nz=8 (2048 atoms), 100 steps, thermo every step   -> per-step parity dumps
  md_t2     nx=ny=nz=12 (6912 atoms), 60 steps, thermo every step, reneighbour every 5
  dem_t1    examples/dem.py on a 0.1 x 0.015 x 0.04 box (420 spheres + 2 planes), 700 steps, thermo hook every step
  dem_nl_t1 dem_t1 with psim.build_neighbor_lists(linkedCellWidth) instead of build_cell_lists: Verlet lists + BuildContactHistory
  dem_vtk_t1 dem_t1 for 60 steps with the example's psim.vtk_output(..., frequency) kept, writing every 30 iterations
  dem_cn_t1 dem_t1 with build_cell_lists(..., store_neighbors_per_cell=True)
  dem_rn3_t1 dem_t1 with psim.reneighbor_every(3): exchange / borders / cell lists every third iteration, synchronize in between
  dem_more_t1 dem_t1 with three further contact properties kept by the contact model (contact tables beyond examples/dem.py's three)
  dem_stock_t1 examples/dem.py as shipped (0.8 x 0.015 x 0.2 box, VTK every 100 iterations), cut to 103 iterations
  dem_bench examples/dem.py on the 0.8 x 0.8 x 0.2 box (998400 spheres), bounded by the harness
  md_custom_t1  md_t1 with other kernel bodies (softened LJ using sqrt / select / symbols, integrators with drag) -> generic kernels
  md_props_t1   md_t1 with user-defined properties (a real entering the pair force, written by a setup() function; a second volatile
              vector; reals and a vector integrated per particle) -> generic kernels on the user-property rows (csrc/props.cu)
  md_vocab_t1   md_t1 with kernels that use the rest of the generic vocabulary (skip_when, cross, is_point_mass, integer properties
              and operators, and / or / not, n-ary min / max, normalized, length, ...): module-level pin of pairs_b200/kernelgen.py
  md_cells_t1 md_t1 (60 steps) with psim.build_cell_lists(cutoff_radius + skin) instead of build_neighbor_lists: the pair kernel walks the
              cell lists (sim/interaction.py:92-118), cells rebuilt at the reneighbouring interval
  md_half_t1  md_t1 with psim.compute_half() enabled (the line is commented out in the stock example)
  md_bench  nx=ny=nz=63 (1000188 atoms), up to 2000 steps, thermo every step (the hook that lets
            bench.py time a bounded number of loop iterations and then leave the loop)
            -> CPU-baseline sample for bench.py

  md_c1_cuda / md_c2_cuda   the reference's own CUDA target (`md.py gpu`) for configs C1 / C2, an executable for the GPU box:
            the SECONDARY baseline (bench.py reports it next to the CPU baseline)

Usage: python oracle/build_ref.py [variant ...]    (default: all; no-op if /root/reference is absent)
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PAIRS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(OUT, "gen")

RUNTIME_SRCS = ["runtime/pairs.cpp", "runtime/domain/regular_6d_stencil.cpp", "runtime/devices/dummy.cpp"]
# -ffp-contract=off: no FMA contraction, so per-operation IEEE results are comparable with a
# CUDA build compiled with --fmad=false.  No -ffast-math (Makefile:18 uses it for the GPU target only).
CXXFLAGS = ["-O3", "-ffp-contract=off", "-w"]


def _sub(text, pattern, repl, count=1):
    new, n = re.subn(pattern, repl, text, count=count, flags=re.M)
    if n == 0:
        raise RuntimeError(f"pattern not found in reference example: {pattern}")
    return new


def md_variant(nx, steps, thermo, reneigh, pcap=None, half=False, cells_only=False):
    def patch(text):
        if cells_only:  # no Verlet lists: the pair kernel walks cell 0 + the 27 stencil cells (sim/interaction.py:92-118)
            text = _sub(text, r"^psim\.build_neighbor_lists\(cutoff_radius \+ skin\)", "psim.build_cell_lists(cutoff_radius + skin)")
        if half:        # the line is present but commented out in the stock example (examples/md.py:58)
            text = _sub(text, r"^#psim\.compute_half\(\)", "psim.compute_half()")
        text = _sub(text, r"^nx = \d+", f"nx = {nx}")
        text = _sub(text, r"^ny = \d+", f"ny = {nx}")
        text = _sub(text, r"^nz = \d+", f"nz = {nx}")
        extra = f", particle_capacity={pcap}" if pcap else ""
        text = _sub(text, r"timesteps=\d+", f"timesteps={steps}{extra}")
        text = _sub(text, r"compute_thermo\(\d+\)", f"compute_thermo({thermo})")
        text = _sub(text, r"reneighbor_every\(\d+\)", f"reneighbor_every({reneigh})")
        return text
    return patch


CUSTOM_KERNELS = '''def lennard_jones(i, j):
    rsq = squared_distance(i, j)
    r = sqrt(rsq)
    sr2 = 1.0 / rsq
    sr6 = sr2 * sr2 * sr2 * sigma6[i, j]
    f = 48.0 * sr6 * (sr6 - 0.5) * sr2 * epsilon[i, j] + select(r < rsoft, kspring * (rsoft - r) / r, 0.0)
    apply(force, delta(i, j) * f)


def initial_integrate(i):
    linear_velocity[i] += (dt * 0.5) * (force[i] - gamma * linear_velocity[i]) / mass[i]
    position[i] += dt * linear_velocity[i]


def final_integrate(i):
    linear_velocity[i] += (dt * 0.5) * (force[i] - gamma * linear_velocity[i]) / mass[i]
'''


def md_custom_variant(nx, steps):
    """examples/md.py with OTHER kernel bodies (a softened LJ with sqrt / select / extra symbols, integrators with a drag term)
    and a non-uniform epsilon table: pins the generic kernel path (pairs_b200/kernelgen.py), same text as tests/scripts/custom_script.py."""
    base = md_variant(nx, steps, 1, 20)

    def patch(text):
        text = base(text)
        text = re.sub(r"def lennard_jones\(i, j\):.*?(?=\n\ncmd = )", CUSTOM_KERNELS.rstrip("\n") + "\n", text, count=1, flags=re.S)
        text = _sub(text, r"\[sigma for i in range\(ntypes \* ntypes\)\]", "[1.0 + 0.05 * ((i % ntypes) + (i // ntypes)) for i in range(ntypes * ntypes)]")
        text = _sub(text, r"symbols=\{'dt': dt\}, pre_step=True", "symbols={'dt': dt, 'gamma': 0.05}, pre_step=True")
        text = _sub(text, r"psim\.compute\(lennard_jones, cutoff_radius\)", "psim.compute(lennard_jones, cutoff_radius, symbols={'kspring': 3.5, 'rsoft': 1.05})")
        text = _sub(text, r"psim\.compute\(final_integrate, symbols=\{'dt': dt\}", "psim.compute(final_integrate, symbols={'dt': dt, 'gamma': 0.05}")
        return text
    return patch


PROPS_KERNELS = '''def init_scale(i):
    scale[i] = 1.0 + 0.25 * (position[i][0] + 0.3719 * position[i][1] + 0.1137 * position[i][2]) / xlen


def lennard_jones(i, j):
    sr2 = 1.0 / squared_distance(i, j)
    sr6 = sr2 * sr2 * sr2 * sigma6[i, j]
    f = 48.0 * sr6 * (sr6 - 0.5) * sr2 * epsilon[i, j] * scale[i]
    apply(force, delta(i, j) * f)
    apply(pull, delta(i, j) * (sr6 * scale[i]))


def initial_integrate(i):
    linear_velocity[i] += (dt * 0.5) * force[i] / mass[i]
    position[i] += dt * linear_velocity[i]
    path[i] += dt * linear_velocity[i]
    heat[i] += dt * dot(force[i], linear_velocity[i])


def final_integrate(i):
    linear_velocity[i] += (dt * 0.5) * force[i] / mass[i]
    work[i] = work[i] + dot(pull[i], linear_velocity[i])
    if dot(pull[i], linear_velocity[i]) > 0.0:
        ups[i] += 1 + (uid[i] & 1)
'''


def md_props_variant(nx, steps):
    """examples/md.py with user-defined properties and kernels that use them: same text as tests/scripts/props_script.py."""
    base = md_variant(nx, steps, 1, 20)

    def patch(text):
        text = base(text)
        text = re.sub(r"def lennard_jones\(i, j\):.*?(?=\n\ncmd = )", PROPS_KERNELS.rstrip("\n") + "\n", text, count=1, flags=re.S)
        text = _sub(text, r"^psim\.add_feature\('type', ntypes\)",
                    "psim.add_property('scale', pairs.real(), 1.0)\npsim.add_property('heat', pairs.real(), 0.0)\n"
                    "psim.add_property('work', pairs.real(), 0.0)\npsim.add_property('path', pairs.vector())\n"
                    "psim.add_property('pull', pairs.vector(), volatile=True)\npsim.add_property('ups', pairs.int32(), 0)\n"
                    "psim.add_feature('type', ntypes)")
        text = _sub(text, r"^psim\.reneighbor_every", "psim.setup(init_scale, symbols={'xlen': 13.0})\npsim.reneighbor_every")
        return text
    return patch


VOCAB_KERNELS = '''def lennard_jones(i, j):
    skip_when(uid[j] % 7 == 3)
    d = delta(i, j)
    dv = linear_velocity[j] - linear_velocity[i]
    w = cross(linear_velocity[i], dv)
    s = select(is_point_mass(j), 1.0, 0.5)
    m = (uid[i] & 3) + (uid[j] | 1) + (type[i] ^ type[j]) + (~uid[j] & 1)
    sr2 = 1.0 / squared_distance(i, j)
    u = normalized(dv)
    c = min(dot(u, linear_velocity[i]), length(dv), 1.5) + max(squared_length(dv), 0.25)
    apply(force, w * s + d * (sr2 * m) + u * c + zero_vector())


def initial_integrate(i):
    linear_velocity[i] += (dt * 0.5) * force[i] / mass[i]
    position[i] += dt * linear_velocity[i]


def final_integrate(i):
    skip_when(not is_point_mass(i))
    if uid[i] % 2 == 0 and mass[i] > 0.5 or shape[i] == 1:
        linear_velocity[i] += (dt * 0.5) * force[i] / mass[i]
'''


def md_vocab_variant(nx, steps):
    """examples/md.py with kernels exercising the rest of the generic vocabulary: same text as tests/scripts/vocab_script.py.  Only its
    modules are used (called on arrays the test provides); the program as a whole is not a meaningful simulation."""
    base = md_variant(nx, steps, 1, 20)

    def patch(text):
        text = base(text)
        return re.sub(r"def lennard_jones\(i, j\):.*?(?=\n\ncmd = )", VOCAB_KERNELS.rstrip("\n") + "\n", text, count=1, flags=re.S)
    return patch


def dem_variant(domain, steps, pcap=None, per_cell=False, vtk_every=None, verlet=False, reneigh=None):
    def patch(text):
        if reneigh is not None:    # cell lists / ghosts rebuilt every `reneigh` iterations, ghosts refreshed by synchronize in between
            text = _sub(text, r"^psim\.generate\(\)", f"psim.reneighbor_every({reneigh})\npsim.generate()")
        if verlet:      # Verlet lists + BuildContactHistory instead of the cell-list traversal (sim/simulation.py:255-261, 402-406)
            text = _sub(text, r"^psim\.build_cell_lists\(linkedCellWidth\)", "psim.build_neighbor_lists(linkedCellWidth)")
        if vtk_every is not None:     # keep psim.vtk_output(...) (runtime/vtk.hpp) and write every `vtk_every` iterations
            text = _sub(text, r"^visSpacing = \d+", f"visSpacing = {vtk_every}")
        if per_cell:    # build_cell_lists(spacing, store_neighbors_per_cell=True), sim/simulation.py:250-253
            text = _sub(text, r"^psim\.build_cell_lists\(linkedCellWidth\)", "psim.build_cell_lists(linkedCellWidth, store_neighbors_per_cell=True)")
        if pcap:
            text = _sub(text, r"particle_capacity=\d+", f"particle_capacity={pcap}")
        text = _sub(text, r"^domainSize_SI = \[[^\]]*\]", f"domainSize_SI = [{domain[0]}, {domain[1]}, {domain[2]}]")
        text = _sub(text, r"^timeSteps = \d+", f"timeSteps = {steps}")
        if vtk_every is None:
            text = _sub(text, r"^psim\.vtk_output\(", "#psim.vtk_output(")      # no files from the parity / timing variants
        # one compute_thermo call per iteration = the harness hook that exposes nlocal and the per-step state
        text = _sub(text, r"^psim\.generate\(\)", "psim.compute_thermo(1)\npsim.generate()")
        return text
    return patch


def dem_more_variant(domain, steps):
    """examples/dem.py with three FURTHER contact properties (a second vector, a second real with default -1, a second integer with
    default 3) and three statements on them at the end of its contact model -- the same text tests/scripts/dem_script.py composes
    with more_contact_props=True: the reference's own generated code for contact tables beyond dem.py's three."""
    base = dem_variant(domain, steps)

    def patch(text):
        text = base(text)
        text = _sub(text, r"^psim\.add_contact_property\('impact_velocity_magnitude', pairs\.real\(\), 0\.0\)",
                    "psim.add_contact_property('impact_velocity_magnitude', pairs.real(), 0.0)\n"
                    "psim.add_contact_property('tsd_seen', pairs.vector(), [0.0, 0.0, 0.0])\n"
                    "psim.add_contact_property('contact_age', pairs.real(), -1.0)\n"
                    "psim.add_contact_property('hits', pairs.int32(), 3)")
        text = _sub(text, r"^    apply\(torque, cross\(contact_point\(i, j\) - position, partial_force\)\)",
                    "    apply(torque, cross(contact_point(i, j) - position, partial_force))\n"
                    "    tsd_seen[i, j] = tangential_spring_displacement[i, j]\n"
                    "    contact_age[i, j] = contact_age[i, j] + 1.0\n"
                    "    hits[i, j] = hits[i, j] + 2")
        return text
    return patch


# The reference's own CUDA target (SECONDARY baseline of BASELINE.json's north_star): `md.py gpu` -> md.cu, compiled with nvcc for
# sm_100a together with the reference's runtime/devices/cuda.cu where they lie.  Stock semantics: nvcc's default FMA contraction
# (the reference's Makefile does not disable it) and its -DENABLE_CUDA_AWARE_MPI; the image has no MPI, with one rank every
# exchange is a copy inside the process and the MPI stand-in is never reached.  name -> (example script, patch)
CUDA_VARIANTS = {
    "md_c1": ("examples/md.py", None),                                                     # 131072 atoms, 200 steps (config C1)
    "md_c2": ("examples/md.py", lambda t: md_variant(100, 200, 100, 20, pcap=4800000)(t)),   # 4 M atoms, 200 steps (config C2)
    # 1 M atoms: the largest cube that stays inside the generator's fixed ncells_capacity (100000) and send capacity (200000)
    "md_1m": ("examples/md.py", lambda t: md_variant(64, 200, 100, 20, pcap=1400000)(t)),
}


def build_cuda_variant(name):
    import shutil
    script, patch = CUDA_VARIANTS[name]
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        return "skipped (no nvcc)"
    gdir = os.path.join(GEN, name + "_cuda")
    os.makedirs(gdir, exist_ok=True)
    src_script = os.path.join(REF, script)
    gen_cu = os.path.join(gdir, "md.cu")
    exe = os.path.join(OUT, f"{name}_cuda")
    text = open(src_script).read()
    if patch is not None:
        text = patch(text)
    stamp, stamp_file = recipe_stamp(text, "cuda", nvcc), os.path.join(gdir, "recipe.stamp")
    if stamp_matches(stamp_file, stamp) and os.path.exists(exe):
        return "up to date"
    scratch_script = os.path.join(gdir, f"md_{name}_input.py")
    with open(scratch_script, "w") as f:
        f.write(text)
    run([sys.executable, scratch_script, "gpu"], cwd=gdir, env=dict(os.environ, PYTHONPATH=os.path.join(REF, "src"), PYTHONHASHSEED="0"))
    if not os.path.exists(gen_cu):
        raise RuntimeError(f"generator did not write {gen_cu}")
    inc = ["-I" + os.path.join(HERE, "mpi_stub"), "-I" + REF, "-I" + os.path.join(REF, "runtime")]
    srcs = [gen_cu] + [os.path.join(REF, s) for s in ("runtime/pairs.cpp", "runtime/domain/regular_6d_stencil.cpp", "runtime/devices/cuda.cu")]
    # -DENABLE_CUDA_AWARE_MPI is the reference Makefile's own setting (Makefile:21); with one rank every transfer is then a
    # device-to-device copy inside the process (the host-staged alternative overruns its size arrays, runtime/pairs.cpp:426)
    run([nvcc, "-O3", "-w", "-DENABLE_CUDA_AWARE_MPI", "-gencode", "arch=compute_100a,code=sm_100a", *inc, *srcs, "-o", exe])
    with open(stamp_file, "w") as f:
        f.write(stamp + "\n")
    return "built"


VARIANTS = {
    # name: (example script, patch function or None, harness defines, also build executable)
    "md": ("examples/md.py", None, ["-DREF_IS_MD"], True),
    "md_t1": ("examples/md.py", md_variant(8, 100, 1, 20), ["-DREF_IS_MD"], False),
    "md_t2": ("examples/md.py", md_variant(12, 60, 1, 5), ["-DREF_IS_MD"], False),
    "md_bench": ("examples/md.py", md_variant(63, 2000, 1, 20, pcap=1400000), ["-DREF_IS_MD"], False),
    # the same program with the flags of the reference's own Makefile (contraction allowed, AVX2 + FMA): the kinder CPU baseline
    "md_bench_fma": ("examples/md.py", md_variant(63, 2000, 1, 20, pcap=1400000), ["-DREF_IS_MD", "-ffp-contract=fast", "-mavx2", "-mfma"], False),
    # cell lists without neighbour lists: every pair of the 27 stencil cells inside the cutoff, cells rebuilt every 20 iterations
    "md_cells_t1": ("examples/md.py", md_variant(8, 60, 1, 20, cells_only=True), [], False),
    "md_custom_t1": ("examples/md.py", md_custom_variant(8, 100), ["-DREF_LJ_MODULE"], False),
    "md_props_t1": ("examples/md.py", md_props_variant(8, 100), ["modules:init_scale,lennard_jones,initial_integrate,final_integrate"], False),
    "md_vocab_t1": ("examples/md.py", md_vocab_variant(8, 10), ["modules:lennard_jones,final_integrate"], False),
    # half neighbour lists + atomic update of the partner
    "md_half_t1": ("examples/md.py", md_variant(8, 100, 1, 20, half=True), ["-DREF_IS_MD", "-DREF_HALF_LISTS"], False),
    # examples/dem.py: spheres + 2 half-spaces, contact history, cell lists only, reneighbour every step
    "dem_t1": ("examples/dem.py", dem_variant((0.1, 0.015, 0.04), 700), [], False),
    "dem_cn_t1": ("examples/dem.py", dem_variant((0.1, 0.015, 0.04), 700, per_cell=True), [], False),
    "dem_nl_t1": ("examples/dem.py", dem_variant((0.1, 0.015, 0.04), 700, verlet=True), [], False),
    "dem_rn3_t1": ("examples/dem.py", dem_variant((0.1, 0.015, 0.04), 400, reneigh=3), [], False),
    "dem_vtk_t1": ("examples/dem.py", dem_variant((0.1, 0.015, 0.04), 60, vtk_every=30), [], False),
    "dem_bench": ("examples/dem.py", dem_variant((0.8, 0.8, 0.2), 100000, pcap=1300000), [], False),
    # the STOCK example (0.8 x 0.015 x 0.2 box, 18720 spheres, VTK every 100 iterations), only cut to 100 iterations: golden for the
    # test that runs the example file itself on this backend (tests/test_gpu_examples.py)
    "dem_stock_t1": ("examples/dem.py", dem_variant((0.8, 0.015, 0.2), 102, vtk_every=100), [], False),
    # dem_t1 generated by a CORRECTED copy of the reference generator: see corrected_generator()
    "dem_fix_t1": ("examples/dem.py", dem_variant((0.1, 0.015, 0.04), 700), ["generator:contact_history_fix"], False),
    # dem_t1 with three further contact properties kept by the contact model (contact tables beyond dem.py's three)
    "dem_more_t1": ("examples/dem.py", dem_more_variant((0.1, 0.015, 0.04), 400), [], False),
}


def corrected_generator(gdir):
    """A scratch copy of /root/reference/src/pairs (under oracle/_ref/, never committed) whose sim/comm.py carries the minimal
    fix of the contact-history transfer, the defects SURVEY.md Appendix A.2 lists:
      * PackContactHistoryData runs BEFORE RemoveExchangedParticles has overwritten the leaver's slot with the hole filler's row
        (stock: after -- the history of the WRONG particle is sent), into a send buffer of its own (at that point the shared one
        still holds the particle records);
      * every record carries its partner uid at `offset + disp + 0` (stock: all records write to `offset + 0`);
      * unpack reads the uid of record k from `offset + disp + 0` and the contact properties from `offset + disp + 1` on, where
        pack put them (stock: `offset + 0` and `cp_offset = 0`).
    Returns the directory to put on PYTHONPATH."""
    import shutil
    dst = os.path.join(gdir, "src_fixed")
    if os.path.exists(dst):
        shutil.rmtree(dst)
    shutil.copytree(os.path.join(REF, "src"), dst)
    path = os.path.join(dst, "pairs", "sim", "comm.py")
    t = open(path).read()

    def sub(old, new, count=1):
        nonlocal t
        assert t.count(old) >= count, f"corrected_generator: pattern not found in comm.py: {old[:60]}"
        t = t.replace(old, new, count)

    # ... into a buffer of its own: at that point send_buffer still holds the particle records CommunicateData is about to send
    sub("""        self.send_buffer      = sim.add_array('send_buffer', [self.send_capacity, self.elem_capacity], Types.Real, arr_sync=False)
""", """        self.send_buffer      = sim.add_array('send_buffer', [self.send_capacity, self.elem_capacity], Types.Real, arr_sync=False)
        self.contact_send_buffer = sim.add_array('contact_send_buffer', [self.send_capacity, self.elem_capacity], Types.Real, arr_sync=False)
""")
    sub("""        send_buffer = self.comm.send_buffer
        send_buffer.set_stride(1, 1)
        contact_soffsets = self.comm.contact_soffsets""", """        send_buffer = self.comm.contact_send_buffer
        send_buffer.set_stride(1, 1)
        contact_soffsets = self.comm.contact_soffsets""")
    sub("""                   self.comm.send_buffer, self.comm.contact_soffsets, self.comm.nsend_contact,""",
        """                   self.comm.contact_send_buffer, self.comm.contact_soffsets, self.comm.nsend_contact,""")
    # pack before the removal (the pack statement keeps its place in the class, only the call moves)
    sub("""            RemoveExchangedParticles_part1(self)
""", """            if self.sim._use_contact_history:
                PackContactHistoryData(self, step)

            RemoveExchangedParticles_part1(self)
""")
    sub("""            if self.sim._use_contact_history:
                PackContactHistoryData(self, step)
                CommunicateContactHistoryData(self, step)""", """            if self.sim._use_contact_history:
                CommunicateContactHistoryData(self, step)""")
    sub("""                    Assign(self.sim, send_buffer[soff][offset + 0],
                                     Cast(self.sim, contact_lists[m][k], Types.Real))""",
        """                    Assign(self.sim, send_buffer[soff][offset + disp + 0],
                                     Cast(self.sim, contact_lists[m][k], Types.Real))""")
    sub("""                    disp = k * nelems_per_contact
                    cp_offset = 0

                    Assign(self.sim, contact_lists[nlocal + i][k], recv_buffer[roff][offset + 0])""",
        """                    disp = k * nelems_per_contact
                    cp_offset = 1

                    Assign(self.sim, contact_lists[nlocal + i][k], recv_buffer[roff][offset + disp + 0])""")
    with open(path, "w") as f:
        f.write(t)
    return dst


def run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError(f"command failed: {' '.join(cmd)}\n{r.stdout[-4000:]}")
    return r.stdout


def recipe_stamp(*parts):
    """What a built artefact depends on, as one hash: the (patched) input script, the flags, the harness and the MPI stand-in --
    not this file's mtime, so that adding a variant does not rebuild the others."""
    import hashlib
    h = hashlib.sha1()
    for p in parts:
        h.update(repr(p).encode())
    for f in (os.path.join(HERE, "ref_harness.cpp"), os.path.join(HERE, "mpi_stub", "mpi.h")):
        h.update(open(f, "rb").read())
    return h.hexdigest()


def stamp_matches(path, stamp):
    return os.path.exists(path) and open(path).read().strip() == stamp


def write_module_wrappers(variant, gen_cpp, modules, gdir):
    """The generator orders a module's arguments differently from script to script: read the printed signatures, write
    `void ref_mod_<module>(void **args)` wrappers that forward args in that order, and record the order for oracle/ref.py."""
    import json
    text = open(gen_cpp).read()
    sigs, lines = {}, []
    for m in modules:
        mt = re.search(rf"^void {m}\(PairsSimulation \*pairs(.*?)\) \{{$", text, flags=re.M)
        if mt is None:
            raise RuntimeError(f"module {m} not found in {gen_cpp}")
        args = []
        for a in [x.strip() for x in mt.group(1).split(",") if x.strip()]:
            ctype, aname = a.rsplit(" ", 1)
            ptr = aname.startswith("*")
            args.append([ctype + (" *" if ptr else ""), aname.lstrip("*")])
        sigs[m] = args
        call = ", ".join(f"({t}) args[{k}]" if t.endswith("*") else f"*({t} *) args[{k}]" for k, (t, _) in enumerate(args))
        lines.append(f"void ref_mod_{m}(void **args) {{ {m}(nullptr, {call}); }}")
    inc_file = os.path.join(gdir, "module_wrappers.inc")
    with open(inc_file, "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(os.path.join(OUT, f"modules_{variant}.json"), "w") as f:
        json.dump(sigs, f, indent=1)
    return inc_file


def build_variant(name):
    script, patch, defines, want_exe = VARIANTS[name]
    gdir = os.path.join(GEN, name)
    os.makedirs(gdir, exist_ok=True)
    src_script = os.path.join(REF, script)
    base = os.path.splitext(os.path.basename(script))[0]          # md / dem -> generated file name
    gen_cpp = os.path.join(gdir, f"{base}.cpp")
    lib = os.path.join(OUT, f"libref_{name}.so")
    exe = os.path.join(OUT, f"{name}_cpu")
    text = open(src_script).read()
    if patch is not None:
        text = patch(text)
    stamp, stamp_file = recipe_stamp(text, defines, CXXFLAGS, want_exe), os.path.join(gdir, "recipe.stamp")
    if stamp_matches(stamp_file, stamp) and os.path.exists(lib) and (not want_exe or os.path.exists(exe)):
        return "up to date"

    # 1. run the reference generator on (a scratch copy of) the reference's example script
    scratch_script = os.path.join(gdir, f"{base}_{name}_input.py")
    with open(scratch_script, "w") as f:
        f.write(text)
    gens = [d for d in defines if d.startswith("generator:")]
    defines = [d for d in defines if not d.startswith("generator:")]
    gen_src = corrected_generator(gdir) if gens else os.path.join(REF, "src")
    env = dict(os.environ, PYTHONPATH=gen_src, PYTHONHASHSEED="0")
    run([sys.executable, scratch_script, "cpu"], cwd=gdir, env=env)
    if not os.path.exists(gen_cpp):
        raise RuntimeError(f"generator did not write {gen_cpp}")

    inc = ["-I" + os.path.join(HERE, "mpi_stub"), "-I" + REF, "-I" + os.path.join(REF, "runtime")]
    rt = [os.path.join(REF, s) for s in RUNTIME_SRCS]
    mods = [d for d in defines if d.startswith("modules:")]
    defines = [d for d in defines if not d.startswith("modules:")]
    if mods:
        inc_file = write_module_wrappers(name, gen_cpp, mods[0][len("modules:"):].split(","), gdir)
        defines.append(f'-DREF_MODULES_INC="{inc_file}"')
    # 2. shared library with hooks + per-module wrappers
    run(["g++", *CXXFLAGS, "-shared", "-fPIC", *defines, f'-DREF_GENERATED="{gen_cpp}"', *inc,
         os.path.join(HERE, "ref_harness.cpp"), *rt, "-o", lib])
    # 3. stock executable (for wall-clock baselines; prints the reference's own timers)
    if want_exe:
        run(["g++", *CXXFLAGS, *inc, gen_cpp, *rt, "-o", exe])
    with open(stamp_file, "w") as f:
        f.write(stamp + "\n")
    return "built"


def stage_data():
    """The generated DEM program reads data/planes.input relative to its working directory (runtime/read_from_file.hpp);
    the reference's 2-row input file is staged under oracle/_ref/data/ so the oracle also runs where /root/reference is absent."""
    import shutil
    ddir = os.path.join(OUT, "data")
    os.makedirs(ddir, exist_ok=True)
    os.makedirs(os.path.join(OUT, "output"), exist_ok=True)
    shutil.copyfile(os.path.join(REF, "data", "planes.input"), os.path.join(ddir, "planes.input"))
    # The reference's example scripts themselves, verbatim, for tests/test_gpu_examples.py ("examples/md.py, lj_onetype.py and dem.py run
    # unchanged", BASELINE.json): the GPU box has no /root/reference, and nothing of the reference is copied into the repository --
    # like every other artefact under oracle/_ref/ these copies are git-ignored build outputs that travel with the snapshot.
    edir = os.path.join(OUT, "examples")
    os.makedirs(edir, exist_ok=True)
    for name in ("md.py", "dem.py", "lj_onetype.py"):
        shutil.copyfile(os.path.join(REF, "examples", name), os.path.join(edir, name))
    shutil.copyfile(os.path.join(REF, "data", "minimd_setup_4x4x4_onetype.input"), os.path.join(ddir, "minimd_setup_4x4x4_onetype.input"))


def main(argv):
    if not os.path.isdir(REF):
        print(f"[oracle/_ref] {REF} not present: keeping prebuilt artefacts")
        return 0
    names = argv or (list(VARIANTS) + [n + "_cuda" for n in CUDA_VARIANTS])
    os.makedirs(GEN, exist_ok=True)
    stage_data()
    for n in names:
        if n.endswith("_cuda"):
            print(f"[oracle/_ref] {n}: {build_cuda_variant(n[:-5])}")
        else:
            print(f"[oracle/_ref] {n}: {build_variant(n)}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
