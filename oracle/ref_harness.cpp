// ref_harness.cpp -- C entry points around the reference's own dycore routines, transliterated from Fortran to C++
// by oracle/f2cpp.py into oracle/_ref/ti_ref.inc (generated at build time from /root/reference, never committed).
//
// TEST INFRASTRUCTURE ONLY: loaded by tests/test_reference_pin.py (through oracle/ref.py) to pin the hand-written oracle
// against the reference's own statements.  Hand-written here: the pool plumbing and, per entry point, the ONE call that
// atm_srk3 makes to the routine (mpas_atm_time_integration.F line cited), with the index ranges of mpas_atm_threading.F:
// one thread = the MPI-only run (cellThreadStart = 1, cellThreadEnd = nCells, ...; :114-125), n threads = the MPAS_OPENMP
// build (:100-111), every routine entered by all threads at once and synchronised by its own !$OMP BARRIERs.
// No arithmetic of the model is restated in this file.
#include <omp.h>
#include "f2cpp_rt.h"
#include "_ref/ti_ref.inc"

struct RefBlock {
    std::map<std::string, int> dims;
    std::map<std::string, CfgVal> cfgs;
    Pool state, diag, mesh, tend, tend_physics, configs, halo_scratch, dimensions;
    std::map<std::string, PoolEntry> module_arrays;
    BlockT block;
    std::string exchange_log;
    int n_threads = 1;
    RefBlock() {
        Pool* ps[] = {&state, &diag, &mesh, &tend, &tend_physics, &configs, &halo_scratch, &dimensions};
        const char* names[] = {"state", "diag", "mesh", "tend", "tend_physics", "configs", "halo_scratch", "dimensions"};
        for (int i = 0; i < 8; i++) { ps[i]->name = names[i]; ps[i]->dims = &dims; ps[i]->cfgs = &cfgs; }
    }
    Pool* pool(const std::string& n) {
        if (n == "state") return &state; if (n == "diag") return &diag; if (n == "mesh") return &mesh; if (n == "tend") return &tend;
        if (n == "tend_physics") return &tend_physics; if (n == "halo_scratch") return &halo_scratch; return nullptr;
    }
};

template <class T> static void bind_mod(FArr<T>& a, const PoolEntry& e) {
    a.p = (T*)e.p[0]; a.rank = e.rank;
    for (int d = 0; d < 3; d++) { a.lo[d] = 1; a.n[d] = e.n[d]; }
}

// module variables of atm_time_integration / mpas_atm_dimensions are globals of the generated code: load this block's
static void activate(RefBlock* b) {
    nvertlevels = b->dims.at("nVertLevels"); maxedges = b->dims.at("maxEdges"); maxedges2 = b->dims.at("maxEdges2");
    num_scalars = b->dims.at("num_scalars");
    config_apply_lbcs = b->cfgs.at("config_apply_lbcs").i != 0;                       // TI:773-775
    struct { const char* n; FArr<real>* a; } tab[] = {
        {"tend_ru_physics", &tend_ru_physics}, {"tend_rtheta_physics", &tend_rtheta_physics}, {"tend_rho_physics", &tend_rho_physics},
        {"qtot", &qtot}, {"delsq_theta", &delsq_theta}, {"delsq_w", &delsq_w}, {"delsq_divergence", &delsq_divergence},
        {"delsq_u", &delsq_u}, {"delsq_vorticity", &delsq_vorticity}, {"dpdz", &dpdz}, {"horiz_flux_array", &horiz_flux_array},
        {"scalar_old_arr", &scalar_old_arr}, {"scalar_new_arr", &scalar_new_arr}, {"s_max_arr", &s_max_arr}, {"s_min_arr", &s_min_arr},
        {"flux_array", &flux_array}, {"flux_upwind_tmp_arr", &flux_upwind_tmp_arr}, {"flux_tmp_arr", &flux_tmp_arr},
        {"wdtn_arr", &wdtn_arr}, {"rho_zz_int", &rho_zz_int}, {"ke_vertex", &ke_vertex}, {"ke_edge", &ke_edge}};
    for (auto& t : tab) {
        auto it = b->module_arrays.find(t.n);
        if (it == b->module_arrays.end()) { fprintf(stderr, "ref_harness: module array %s not bound\n", t.n); abort(); }
        bind_mod(*t.a, it->second);
    }
    bind_mod(bdymaskedge, b->module_arrays.at("bdyMaskEdge"));
}

extern "C" {
void* ref_create(void) { return new RefBlock(); }
void ref_destroy(void* h) { delete (RefBlock*)h; }
int ref_real_bytes(void) { return (int)sizeof(real); }
void ref_set_dim(void* h, const char* name, int v) { ((RefBlock*)h)->dims[name] = v; }
void ref_set_cfg_real(void* h, const char* name, double v) { CfgVal c; c.kind = 0; c.r = v; ((RefBlock*)h)->cfgs[name] = c; }
void ref_set_cfg_int(void* h, const char* name, int v) { CfgVal c; c.kind = 1; c.i = v; ((RefBlock*)h)->cfgs[name] = c; }
void ref_set_cfg_str(void* h, const char* name, const char* v) { CfgVal c; c.kind = 3; c.s = v; ((RefBlock*)h)->cfgs[name] = c; }
// pool == "module": a module variable of atm_time_integration
int ref_bind_array(void* h, const char* pool, const char* key, int lev, void* ptr, int is_int, int rank, long n0, long n1, long n2) {
    RefBlock* b = (RefBlock*)h;
    std::map<std::string, PoolEntry>* m = nullptr;
    if (std::string(pool) == "module") m = &b->module_arrays;
    else { Pool* p = b->pool(pool); if (!p) return 1; m = &p->arrays; }
    PoolEntry& e = (*m)[key];
    e.p[lev - 1] = ptr; e.is_int = is_int != 0; e.rank = rank; e.n[0] = n0; e.n[1] = n1; e.n[2] = n2;
    return 0;
}
const char* ref_exchange_log(void* h) { return ((RefBlock*)h)->exchange_log.c_str(); }

// One routine of the step.  ia / ra: the integer and real arguments in the order of the oracle's entry points.
int ref_call(void* h, const char* routine, const int* ia, const double* ra) {
    RefBlock* b = (RefBlock*)h;
    activate(b);
    const std::string r = routine;
    const int nCells = b->dims.at("nCells"), nEdges = b->dims.at("nEdges"), nVertices = b->dims.at("nVertices");
    const int nCellsSolve = b->dims.at("nCellsSolve"), nEdgesSolve = b->dims.at("nEdgesSolve"), nVerticesSolve = b->dims.at("nVerticesSolve");
    const int nVertLevels = b->dims.at("nVertLevels");
    // a single block exchanges nothing; the group names the reference asks for are recorded for the test
    ExchFn xch = [b](const std::string& g) { b->exchange_log += g + ";"; };
    // Threading as the reference does it (MPAS_OPENMP): every routine is called by all threads at once, each with its own
    // contiguous index ranges (mpas_atm_threading.F:100-111: start = tid * n / nThreads + 1, end = (tid + 1) * n / nThreads),
    // and the routines synchronise among themselves with the !$OMP BARRIERs that f2cpp.py keeps.
    int rc = 0;
    const int nT = b->n_threads;
    // (see the note on the reference's OpenMP race after the parallel region) last dynamics substep: nothing re-copies the
    // columns that get zeroed, so they are saved here and put back below
    std::vector<real> saved;
    const bool last_finish = nT > 1 && r == "rk_dynamics_substep_finish" && ia[0] >= ia[1];
    if (last_finish) {
        FArr<real> th = b->state.arr<real>("theta_m", 1, 2);
        for (int t = 1; t < nT; t++) for (int k = 1; k <= nVertLevels; k++) saved.push_back(th(k, (long)t * nCells / nT + 1));
    }
#pragma omp parallel num_threads(nT)
    {
    const int tid = omp_get_thread_num(), nthr = omp_get_num_threads();
#define RANGE(n) (int)((long)tid * (n) / nthr) + 1, (int)((long)(tid + 1) * (n) / nthr)
#define CELLS RANGE(nCells)
#define VERTS RANGE(nVertices)
#define EDGES RANGE(nEdges)
#define CELLS_SOLVE RANGE(nCellsSolve)
#define VERTS_SOLVE RANGE(nVerticesSolve)
#define EDGES_SOLVE RANGE(nEdgesSolve)
    if (r == "rk_integration_setup")                          // TI:1083
        atm_rk_integration_setup(b->state, b->diag, nVertLevels, b->dims.at("num_scalars"), CELLS, VERTS, EDGES, CELLS_SOLVE, VERTS_SOLVE, EDGES_SOLVE);
    else if (r == "compute_moist_coefficients")               // TI:1102
        atm_compute_moist_coefficients(b->dimensions, b->state, b->diag, b->mesh, CELLS, VERTS, EDGES, CELLS_SOLVE, VERTS_SOLVE, EDGES_SOLVE);
    else if (r == "compute_vert_imp_coefs")                   // TI:1117
        atm_compute_vert_imp_coefs(b->state, b->mesh, b->diag, b->configs, nVertLevels, (real)ra[0], CELLS, EDGES, CELLS_SOLVE, EDGES_SOLVE);
    else if (r == "compute_dyn_tend")                         // TI:1173
        atm_compute_dyn_tend(b->tend, b->tend_physics, b->state, b->diag, b->mesh, b->configs, nVertLevels, ia[0], (real)ra[0],
                             CELLS, VERTS, EDGES, CELLS_SOLVE, VERTS_SOLVE, EDGES_SOLVE);
    else if (r == "set_smlstep_pert_variables")               // TI:1212
        atm_set_smlstep_pert_variables(b->tend, b->mesh, CELLS_SOLVE);
    else if (r == "advance_acoustic_step")                    // TI:1285
        atm_advance_acoustic_step(b->state, b->diag, b->tend, b->mesh, b->configs, nCells, nVertLevels, (real)ra[0], ia[0],
                                  CELLS, VERTS, EDGES, CELLS_SOLVE, VERTS_SOLVE, EDGES_SOLVE);
    else if (r == "divergence_damping_3d")                    // TI:1310
        atm_divergence_damping_3d(b->state, b->diag, b->mesh, b->configs, (real)ra[0], EDGES);
    else if (r == "recover_large_step_variables")             // TI:1328
        atm_recover_large_step_variables(b->state, b->diag, b->tend, b->mesh, b->configs, (real)ra[0], ia[0], ia[1],
                                         CELLS, VERTS, EDGES, CELLS_SOLVE, VERTS_SOLVE, EDGES_SOLVE);
    else if (r == "compute_solve_diagnostics")                // TI:1451 (time level 2, rk_step given)
        atm_compute_solve_diagnostics((real)ra[0], b->state, 2, b->diag, b->mesh, b->configs, CELLS, VERTS, EDGES, Opt<int>(ia[0]));
    else if (r == "init_solve_diagnostics")                   // mpas_atm_core.F:515-527 (time level 1, no rk_step)
        atm_compute_solve_diagnostics((real)ra[0], b->state, 1, b->diag, b->mesh, b->configs, CELLS, VERTS, EDGES);
    else if (r == "init_coupled_diagnostics")                 // mpas_atm_core.F:509
        atm_init_coupled_diagnostics(b->state, 1, b->diag, b->mesh, b->configs, CELLS, VERTS, EDGES, CELLS_SOLVE, VERTS_SOLVE, EDGES_SOLVE);
    else if (r == "rk_dynamics_substep_finish")               // TI:1502
        atm_rk_dynamics_substep_finish(b->state, b->diag, nVertLevels, ia[0], ia[1], CELLS, VERTS, EDGES, CELLS_SOLVE, VERTS_SOLVE, EDGES_SOLVE);
    else if (r == "advance_scalars")                          // advance_scalars, TI:1846: horiz_flux_array, advance_density = config_split_dynamics_transport
        atm_advance_scalars("scalars", b->tend, b->state, b->diag, b->mesh, b->configs, (real)ra[0], EDGES, CELLS_SOLVE,
                            horiz_flux_array, ia[0], b->cfgs.at("config_time_integration_order").i,
                            Opt<bool>(b->cfgs.at("config_split_dynamics_transport").i != 0));
    else if (r == "advance_scalars_mono")                     // advance_scalars, TI:1897
        atm_advance_scalars_mono("scalars", b->block, b->tend, b->state, b->diag, b->mesh, b->halo_scratch, b->configs, (real)ra[0],
                                 CELLS, EDGES, CELLS_SOLVE, scalar_old_arr, scalar_new_arr, s_max_arr, s_min_arr, wdtn_arr,
                                 flux_array, flux_upwind_tmp_arr, flux_tmp_arr, xch,
                                 Opt<bool>(b->cfgs.at("config_split_dynamics_transport").i != 0), rho_zz_int);
    // ---- regional path (config_apply_lbcs).  The driving fields come from mpas_atm_get_bdy_tend / _state
    // (mpas_atm_boundaries.F:375-437, 473-674), restated by bdy_tend / bdy_state below: tendency = time level 1 of lbc_<field>,
    // state at now + delta_t = level 2 - (seconds to the end of the LBC interval - delta_t) * level 1.  One thread, as TI does
    // (`do thread=1,nThreads` loops inside a single region are serial calls with thread ranges: same result).
    else if (r.rfind("lbc_", 0) == 0) {
#pragma omp master
        {
        const real dt_end = (real)b->cfgs.at("lbc_dt_end").r;
        const int S = b->dims.at("num_scalars");
        auto bdy_state = [&](const char* f, real delta_t, std::vector<real>& out, long n) {
            FArr<real> tend = b->mesh.arr<real>(f, 1, 0), state = b->mesh.arr<real>(f, 2, 0);
            real dtl = dt_end; dtl = dtl - delta_t;
            out.resize(n);
            for (long q = 0; q < n; q++) out[q] = state.p[q] - dtl * tend.p[q];
        };
        auto view2 = [&](std::vector<real>& v, long n1, long n2) { FArr<real> a; a.bind(v.data(), 1, n1, 1, n2); return a; };
        std::vector<real> va, vb, vc;
        const int one = 1;
        if (r == "lbc_speczone_tend")                        // TI:1223-1235
            atm_bdy_adjust_dynamics_speczone_tend(b->tend, b->mesh, b->configs, nVertLevels, b->mesh.arr<real>("lbc_ru", 1, 2),
                                                  b->mesh.arr<real>("lbc_rtheta_m", 1, 2), b->mesh.arr<real>("lbc_rho_zz", 1, 2),
                                                  one, nCells, one, nEdges, one, nCellsSolve, one, nEdgesSolve);
        else if (r == "lbc_relaxzone_tend") {                // TI:1246-1261: ra[0] = time_dyn_step, ra[1] = dt
            bdy_state("lbc_ru", (real)ra[0], va, (long)nVertLevels * (nEdges + 1));
            bdy_state("lbc_rtheta_m", (real)ra[0], vb, (long)nVertLevels * (nCells + 1));
            bdy_state("lbc_rho_zz", (real)ra[0], vc, (long)nVertLevels * (nCells + 1));
            atm_bdy_adjust_dynamics_relaxzone_tend(b->configs, b->tend, b->state, b->diag, b->mesh, nVertLevels, (real)ra[1],
                                                   view2(va, nVertLevels, nEdges + 1), view2(vb, nVertLevels, nCells + 1), view2(vc, nVertLevels, nCells + 1),
                                                   one, nCells, one, nEdges, one, nCellsSolve, one, nEdgesSolve);
        } else if (r == "lbc_reset_u_ru") {                  // TI:1343-1388 (inline in atm_srk3): ra[0] = time_dyn_step
            FArr<int> bdyMaskEdge = b->mesh.arr<int>("bdyMaskEdge", 1, 1);
            FArr<real> u = b->state.arr<real>("u", 2, 2), ru = b->diag.arr<real>("ru", 1, 2);
            bdy_state("lbc_u", (real)ra[0], va, (long)nVertLevels * (nEdges + 1));
            FArr<real> dv = view2(va, nVertLevels, nEdges + 1);
            for (int iEdge = 1; iEdge <= nEdgesSolve; iEdge++)
                if (bdyMaskEdge(iEdge) > nrelaxzone) for (int k = 1; k <= nVertLevels; k++) u(k, iEdge) = dv(k, iEdge);
            bdy_state("lbc_ru", (real)ra[0], va, (long)nVertLevels * (nEdges + 1));
            dv = view2(va, nVertLevels, nEdges + 1);
            for (int iEdge = 1; iEdge <= nEdges; iEdge++)
                if (bdyMaskEdge(iEdge) > nrelaxzone) for (int k = 1; k <= nVertLevels; k++) ru(k, iEdge) = dv(k, iEdge);
        } else if (r == "lbc_adjust_scalars") {              // TI:1413-1428 / 1565-1580: ra[0] = dt, ra[1] = rk_timestep(rk_step)
            bdy_state("lbc_scalars", (real)ra[1], va, (long)S * nVertLevels * (nCells + 1));
            FArr<real> sd; sd.bind(va.data(), 1, S, 1, nVertLevels, 1, nCells + 1);
            atm_bdy_adjust_scalars(b->state, b->diag, b->mesh, b->configs, sd, nVertLevels, (real)ra[0], (real)ra[1], one, nCells, one, nCellsSolve);
        } else if (r == "lbc_zero_gradient_w")               // TI:1477-1484
            atm_zero_gradient_w_bdy(b->state, b->mesh, one, nCellsSolve);
        else if (r == "lbc_reset_speczone_values") {         // TI:1676-1695: ra[0] = dt
            bdy_state("lbc_rtheta_m", (real)ra[0], va, (long)nVertLevels * (nCells + 1));
            bdy_state("lbc_rho_zz", (real)ra[0], vb, (long)nVertLevels * (nCells + 1));
            atm_bdy_reset_speczone_values(b->state, b->diag, b->mesh, nVertLevels, view2(va, nVertLevels, nCells + 1), view2(vb, nVertLevels, nCells + 1),
                                          one, nCells, one, nCellsSolve);
        } else if (r == "lbc_set_scalars") {                 // TI:1700-1719: ra[0] = dt
            bdy_state("lbc_scalars", (real)ra[0], va, (long)S * nVertLevels * (nCells + 1));
            FArr<real> sd; sd.bind(va.data(), 1, S, 1, nVertLevels, 1, nCells + 1);
            atm_bdy_set_scalars(b->state, b->mesh, sd, nVertLevels, one, nCells, one, nCellsSolve);
        } else { fprintf(stderr, "ref_call: unknown routine %s\n", routine); rc = 1; }
        }
    }
    // init-time derivations of atm_mpas_init_block (mpas_atm_core.F:573-586), serial as in the reference
    else if (r == "compute_mesh_scaling") {
#pragma omp master
        atm_compute_mesh_scaling(b->mesh, b->configs);
    } else if (r == "compute_signs") {
#pragma omp master
        atm_compute_signs(b->mesh);
    } else if (r == "compute_damping_coefs") {
#pragma omp master
        atm_compute_damping_coefs(b->mesh, b->configs);
    } else if (r == "adv_coef_compression") {
#pragma omp master
        atm_adv_coef_compression(b->mesh);
    } else if (r == "couple_coef_3rd_order") {
#pragma omp master
        atm_couple_coef_3rd_order(b->mesh, b->configs);
    }
    else {
#pragma omp master
        { fprintf(stderr, "ref_call: unknown routine %s\n", routine); rc = 1; }
    }
    }
    // The reference's OpenMP build has a data race in its two copy routines: every thread executes
    // `theta_m_2(:,cellEnd+1) = 0` (TI:1987; `theta_m_1` at TI:7082) with ITS cellEnd, i.e. it zeroes the first column of the
    // next thread's range, which that thread may already have copied.  Intended is the garbage column nCells+1 only (what one
    // thread does).  To keep the threaded arm's results equal to the MPI-only run, the nT-1 affected columns are re-copied here.
    if (nT > 1 && (r == "rk_integration_setup" || (r == "rk_dynamics_substep_finish" && ia[0] < ia[1]))) {
        const bool setup = r == "rk_integration_setup";
        FArr<real> dst = b->state.arr<real>("theta_m", setup ? 2 : 1, 2), src = b->state.arr<real>("theta_m", setup ? 1 : 2, 2);
        for (int t = 1; t < nT; t++) {
            const long c = (long)t * nCells / nT + 1;
            for (int k = 1; k <= nVertLevels; k++) dst(k, c) = src(k, c);
        }
    }
    if (last_finish) {
        FArr<real> th = b->state.arr<real>("theta_m", 1, 2);
        size_t q = 0;
        for (int t = 1; t < nT; t++) for (int k = 1; k <= nVertLevels; k++) th(k, (long)t * nCells / nT + 1) = saved[q++];
    }
    return rc;
}
void ref_set_threads(void* h, int n) { ((RefBlock*)h)->n_threads = n < 1 ? 1 : n; }
}
