// Harness around the reference's OWN generated serial C++ program (the output of
// /root/reference/examples/{md,dem}.py with target_cpu(), produced at build time into
// oracle/_ref/gen/ by oracle/build_ref.py).  TEST INFRASTRUCTURE ONLY: nothing under
// pairs_b200/ may link, load or call this; only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py use it, as the checker.
//
// What it does: textually includes the generated translation unit with `main` renamed and
// three runtime entry points (register_timer / stop_timer / compute_thermo,
// runtime/timing.hpp:9-19, runtime/thermo.hpp:11-51) routed through hook functions, so a
// test can (a) run the reference's whole timestep loop and observe full-precision state
// (the stock program prints 6 significant digits only, runtime/thermo.hpp:47) and
// (b) call single generated modules on caller-provided arrays.
//
// Build: g++ -O3 -ffp-contract=off -shared -fPIC -DREF_GENERATED='"gen/md.cpp"' [-DREF_IS_MD]
//        -Ioracle/mpi_stub -I/root/reference  (see oracle/build_ref.py)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "runtime/pairs.hpp"
#include "runtime/timing.hpp"
#include "runtime/thermo.hpp"

extern "C" {
// event: "module:<name>" after a timed module finished, "thermo" when compute_thermo is called
// from the timestep loop (a = nlocal), "registered" once timers are known.
typedef void (*ref_hook_fn)(const char *event, int a, void *user);
}

namespace {
pairs::PairsSimulation *g_ps = nullptr;
ref_hook_fn g_hook = nullptr;
void *g_hook_user = nullptr;
std::vector<std::string> g_timer_names;
int g_thermo_calls = 0;
int g_thermo_limit = -1;               // stop the generated loop after this many compute_thermo calls (< 0: never)
bool g_quiet = false;
std::vector<double> g_thermo_times;    // wall-clock seconds (steady clock) at every compute_thermo call
struct RefStop {};
}

namespace pairs {

void ref_hook_register_timer(PairsSimulation *ps, int id, std::string name) {
    g_ps = ps;
    if((int) g_timer_names.size() <= id) { g_timer_names.resize(id + 1); }
    g_timer_names[id] = name;
    register_timer(ps, id, name);
}

void ref_hook_stop_timer(PairsSimulation *ps, int id) {
    stop_timer(ps, id);
    g_ps = ps;
    if(g_hook != nullptr && id > 0 && id < (int) g_timer_names.size()) {
        std::string ev = "module:" + g_timer_names[id];
        g_hook(ev.c_str(), id, g_hook_user);
    }
}

double ref_hook_compute_thermo(PairsSimulation *ps, int nlocal, double xprd, double yprd, double zprd, int print) {
    g_ps = ps;
    double t = compute_thermo(ps, nlocal, xprd, yprd, zprd, g_quiet ? 0 : print);
    g_thermo_calls++;
    g_thermo_times.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count());
    if(g_hook != nullptr) { g_hook("thermo", nlocal, g_hook_user); }
    if(g_thermo_limit >= 0 && g_thermo_calls >= g_thermo_limit) { throw RefStop(); }
    return t;
}

}

#define main ref_generated_main
#define register_timer ref_hook_register_timer
#define stop_timer ref_hook_stop_timer
#define compute_thermo ref_hook_compute_thermo
#include REF_GENERATED
#undef main
#undef register_timer
#undef stop_timer
#undef compute_thermo

extern "C" {

// Runs the generated program's main() (set-up + the full timestep loop + timers/stats print).
int ref_run(ref_hook_fn hook, void *user, int quiet) {
    g_hook = hook;
    g_hook_user = user;
    g_quiet = quiet != 0;
    g_thermo_calls = 0;
    g_thermo_times.clear();
    g_timer_names.clear();
    // the reference draws particle types from rand() in its process-initial state (runtime/copper_fcc_lattice.hpp:128)
    srand(1);
    char arg0[] = "ref";
    char *argv_[] = {arg0, nullptr};
    char **argv = argv_;
    int rc = 0;
    try {
        rc = ref_generated_main(1, argv);
    } catch(const RefStop &) {
        rc = 0;   // bounded run: the generated loop was left after g_thermo_limit thermo calls (its state is leaked)
    }
    g_hook = nullptr;
    g_ps = nullptr;
    g_thermo_limit = -1;
    return rc;
}

// Bounded run for timing: programs generated with compute_thermo(1) call the hook once per loop iteration; the
// loop is left after `max_thermo_calls` iterations.  Timestamps of every iteration end are kept.
int ref_run_limited(int max_thermo_calls, int quiet) {
    g_thermo_limit = max_thermo_calls;
    return ref_run(nullptr, nullptr, quiet);
}

void ref_set_limit(int max_thermo_calls) { g_thermo_limit = max_thermo_calls; }

int ref_thermo_times(double *out, int cap) {
    const int n = (int) g_thermo_times.size();
    for(int k = 0; k < n && k < cap; k++) { out[k] = g_thermo_times[k]; }
    return n;
}

// Valid only inside a hook callback (the PairsSimulation is deleted when main returns).
void *ref_property_ptr(const char *name) {
    if(g_ps == nullptr) { return nullptr; }
    return g_ps->getPropertyByName(name).getHostPointer();
}

void *ref_array_ptr(const char *name) {
    if(g_ps == nullptr) { return nullptr; }
    return g_ps->getArrayByName(name).getHostPointer();
}

void *ref_contact_property_ptr(const char *name) {
    if(g_ps == nullptr) { return nullptr; }
    return g_ps->getContactPropertyByName(name).getHostPointer();
}

long ref_array_size(const char *name) {
    if(g_ps == nullptr) { return -1; }
    return (long) g_ps->getArrayByName(name).getSize();
}

#if defined(REF_LJ_MODULE) && !defined(REF_IS_MD)
// variants with other kernel bodies (md_custom_t1): only the pair module keeps the argument list of examples/md.py
void ref_md_lennard_jones(int neighbor_capacity, int nlocal, int *numneighs, int *neighborlists, int *flags,
                          double *position, int *type, double *force, double *sigma6, double *epsilon) {
    lennard_jones(nullptr, neighbor_capacity, nlocal, numneighs, neighborlists, flags, position, type, force, sigma6, epsilon);
}
#endif

#ifdef REF_MODULES_INC
// Generic module access: oracle/build_ref.py parses the argument lists the generator printed (their order is not the same for
// every script) and writes one wrapper per requested module, `void ref_mod_<name>(void **args)`, args in the generated order
// (scalars by address); the order is recorded next to the library (modules_<variant>.json) and oracle/ref.py passes by NAME.
#include REF_MODULES_INC
#endif

#ifdef REF_IS_MD
// Direct calls into single generated modules (signatures: generated md.cpp, which the
// reference's generator emits deterministically for examples/md.py; `pairs` is unused inside
// compute modules, code_gen/cgen.py:144-200).
void ref_md_lennard_jones(int neighbor_capacity, int nlocal, int *numneighs, int *neighborlists, int *flags,
                          double *position, int *type, double *force, double *sigma6, double *epsilon) {
    lennard_jones(nullptr, neighbor_capacity, nlocal, numneighs, neighborlists, flags, position, type, force, sigma6, epsilon);
}

void ref_md_initial_integrate(int nlocal, int *flags, double *force, double *mass, double *linear_velocity, double *position) {
    initial_integrate(nullptr, nlocal, flags, force, mass, linear_velocity, position);
}

void ref_md_final_integrate(int nlocal, int *flags, double *force, double *mass, double *linear_velocity) {
    final_integrate(nullptr, nlocal, flags, force, mass, linear_velocity);
}

void ref_md_build_cell_lists_stencil(int ncells_capacity, int *ncells, int *nstencil, int *shapes_buffer, double *subdom,
                                     int *dim_cells, int *resizes, int *stencil) {
    build_cell_lists_stencil(nullptr, ncells_capacity, ncells, nstencil, shapes_buffer, subdom, dim_cells, resizes, stencil);
}

void ref_md_build_cell_lists(int ncells, int nlocal, int nghost, int cell_capacity, int *cell_sizes, double *subdom,
                             int *dim_cells, int *particle_cell, int *resizes, int *cell_particles, int *flags, double *position) {
    build_cell_lists(nullptr, ncells, nlocal, nghost, cell_capacity, cell_sizes, subdom, dim_cells, particle_cell, resizes,
                     cell_particles, flags, position);
}

void ref_md_partition_cell_lists(int cell_capacity, int ncells, int *cell_sizes, int *nshapes, int *shapes_buffer,
                                 int *cell_particles, int *shape) {
    partition_cell_lists(nullptr, cell_capacity, ncells, cell_sizes, nshapes, shapes_buffer, cell_particles, shape);
}

#ifdef REF_HALF_LISTS      // compute_half(): the list builder additionally reads the shape ids (sim/interaction.py:107-113)
void ref_md_build_neighbor_lists(int nlocal, int ncells, int cell_capacity, int neighbor_capacity, int nstencil, int *numneighs,
                                 int *particle_cell, int *stencil, int *nshapes, int *cell_particles, int *neighborlists,
                                 int *resizes, int *flags, double *position, int *shape) {
    build_neighbor_lists(nullptr, nlocal, ncells, cell_capacity, neighbor_capacity, nstencil, numneighs, particle_cell, stencil,
                         nshapes, cell_particles, neighborlists, resizes, flags, position, shape);
}
#else
void ref_md_build_neighbor_lists(int nlocal, int ncells, int cell_capacity, int neighbor_capacity, int nstencil, int *numneighs,
                                 int *particle_cell, int *stencil, int *nshapes, int *cell_particles, int *neighborlists,
                                 int *resizes, int *flags, double *position) {
    build_neighbor_lists(nullptr, nlocal, ncells, cell_capacity, neighbor_capacity, nstencil, numneighs, particle_cell, stencil,
                         nshapes, cell_particles, neighborlists, resizes, flags, position);
}
#endif
#endif

}
