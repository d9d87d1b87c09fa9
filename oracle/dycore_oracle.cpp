// dycore_oracle.cpp -- CPU restatement of the MPAS-Atmosphere dycore step.
//
// TEST INFRASTRUCTURE ONLY.  This is the parity oracle for the CUDA library in
// mpas_model_b200/csrc: only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it.  The product never does.
//
// PARITY PIN: the reference ships no golden vectors or logs for the dycore arithmetic and this
// image has no Fortran compiler, so the pin is the reference's own SOURCE TEXT: oracle/f2cpp.py transliterates
//   /root/reference/src/core_atmosphere/dynamics/mpas_atm_time_integration.F  ("TI": every *_work routine of the step,
//   their wrappers, the pool-based routines) and src/framework/mpas_constants.F
// statement by statement into C++ at build time (oracle/_ref/, never committed), and tests/test_reference_pin.py requires
// this hand-written restatement to equal it BIT FOR BIT -- every routine on identical inputs, every scratch array,
// six non-default namelists, an irregular mesh, the single-precision build, and free-running steps.  What that pin
// does not cover is stated in DESIGN.md §2 (gfortran's code generation itself; the routines outside TI).
// This file keeps 1-based indexing so that each loop can be read against the cited lines.
// Compile with -ffp-contract=off (mirrors -Mnofma, reference Makefile:160).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <algorithm>
#include "../include/mpasb.h"

// RKIND: PRECISION=double, or PRECISION=single when built -DORACLE_SINGLE -fsingle-precision-constant (every
// un-suffixed literal is then a float, as the reference's default-kind literals are in its single build:
// reference Makefile:861-873, src/framework/mpas_kind_types.F:22-28)
#ifdef ORACLE_SINGLE
typedef float real;
#else
typedef double real;
#endif

// the namelist in RKIND (mpasb_config carries doubles across the ABI whatever the build precision)
struct OCfg {
    int config_time_integration_order, config_number_of_sub_steps, config_dynamics_split_steps, config_split_dynamics_transport,
        config_scalar_advection, config_monotonic, config_positive_definite, config_horiz_mixing, config_mix_full,
        config_rayleigh_damp_u, config_number_rayleigh_damp_u_levels, config_number_cam_damping_levels, config_apply_lbcs,
        config_print_global_minmax_vel;
    real config_epssm, config_smdiv, config_len_disp, config_coef_3rd_order, config_visc4_2dsmag, config_smagorinsky_coef,
         config_del4u_div_factor, config_h_mom_eddy_visc2, config_h_mom_eddy_visc4, config_v_mom_eddy_visc2,
         config_h_theta_eddy_visc2, config_h_theta_eddy_visc4, config_v_theta_eddy_visc2, config_apvm_upwinding,
         config_mpas_cam_coef, config_rayleigh_damp_u_timescale_days, config_relax_zone_divdamp_coef, cf1, cf2, cf3, sphere_radius;
    int on_a_sphere;
    OCfg() {}
    explicit OCfg(const mpasb_config& c) {
#define CI(n) n = c.n;
#define CR(n) n = (real)c.n;
        CI(config_time_integration_order) CI(config_number_of_sub_steps) CI(config_dynamics_split_steps) CI(config_split_dynamics_transport)
        CI(config_scalar_advection) CI(config_monotonic) CI(config_positive_definite) CI(config_horiz_mixing) CI(config_mix_full)
        CI(config_rayleigh_damp_u) CI(config_number_rayleigh_damp_u_levels) CI(config_number_cam_damping_levels) CI(config_apply_lbcs)
        CI(config_print_global_minmax_vel) CI(on_a_sphere)
        CR(config_epssm) CR(config_smdiv) CR(config_len_disp) CR(config_coef_3rd_order) CR(config_visc4_2dsmag) CR(config_smagorinsky_coef)
        CR(config_del4u_div_factor) CR(config_h_mom_eddy_visc2) CR(config_h_mom_eddy_visc4) CR(config_v_mom_eddy_visc2)
        CR(config_h_theta_eddy_visc2) CR(config_h_theta_eddy_visc4) CR(config_v_theta_eddy_visc2) CR(config_apvm_upwinding)
        CR(config_mpas_cam_coef) CR(config_rayleigh_damp_u_timescale_days) CR(config_relax_zone_divdamp_coef) CR(cf1) CR(cf2) CR(cf3) CR(sphere_radius)
#undef CI
#undef CR
    }
};

// src/framework/mpas_constants.F:43-56
static const real gravity = 9.80616, rgas = 287.0, cp = 7.0 * 287.0 / 2.0, rv = 461.6;
static const real rvord = rv / rgas, cv = cp - rgas, prandtl = 1.0;
static const real seconds_per_day = 86400.0;

// ---------------------------------------------------------------- 1-based views
struct A1 { real* p; inline real& operator()(int i) const { return p[i - 1]; } };
struct A2 { real* p; int n1; inline real& operator()(int k, int i) const { return p[(size_t)(i - 1) * n1 + (k - 1)]; } };
struct A3 { real* p; int n1, n2; inline real& operator()(int a, int b, int c) const { return p[((size_t)(c - 1) * n2 + (b - 1)) * n1 + (a - 1)]; } };
struct I1 { int* p; inline int& operator()(int i) const { return p[i - 1]; } };
struct I2 { int* p; int n1; inline int& operator()(int k, int i) const { return p[(size_t)(i - 1) * n1 + (k - 1)]; } };

enum Loc { CELL, EDGE, VERTEX, LEVS };
enum Inner { ONE, NL, NL1, ME, ME2, VD, TWO, F15, NL1_ME, S_NL, NL_TWO, THREE_ME };
enum Type { REAL, INT };
enum Target { NONE, LOCAL, T_CELL, T_EDGE, T_VERTEX };
struct FieldDef { const char* name; Loc loc; Inner inner; int levels; Type type; };
#define F(name, loc, inner, lev, type, tgt) { #name, loc, inner, lev, type },
static const FieldDef kFields[] = {
#include "../include/mpasb_fields.def"
};
#undef F
static const int kNumFields = sizeof(kFields) / sizeof(kFields[0]);

struct HaloList {      // one (neighbour, layer) pair of a kind (cell/edge/vertex)
    int nbr, layer;
    std::vector<int> send_src, recv_dst;   // 1-based local indices
};

struct Block {
    mpasb_dims d;
    OCfg c;
    std::map<std::string, std::vector<real>> rf;
    std::map<std::string, std::vector<int>> nf;
    std::map<std::string, const FieldDef*> defs;
    std::vector<HaloList> halo[3];
    int rank = 0;
    std::string err;

    int inner1(Inner in) const {
        switch (in) {
            case ONE: return 1; case NL: return d.nVertLevels; case NL1: return d.nVertLevels + 1;
            case ME: return d.maxEdges; case ME2: return d.maxEdges2; case VD: return d.vertexDegree;
            case TWO: return 2; case F15: return 15; case NL1_ME: return d.nVertLevels + 1;
            case S_NL: return d.num_scalars; case NL_TWO: return d.nVertLevels; case THREE_ME: return 3;
        }
        return 1;
    }
    int inner2(Inner in) const {
        switch (in) { case NL1_ME: return d.maxEdges; case S_NL: return d.nVertLevels; case NL_TWO: return 2; case THREE_ME: return d.maxEdges; default: return 1; }
    }
    long outer(Loc l) const {
        switch (l) { case CELL: return d.nCells + 1; case EDGE: return d.nEdges + 1; case VERTEX: return d.nVertices + 1; case LEVS: return 1; }
        return 1;
    }
    long count(const FieldDef* f) const { return (long)inner1(f->inner) * inner2(f->inner) * outer(f->loc); }
    static std::string key(const char* name, int lev) { return std::string(name) + (lev == 2 ? "@2" : ""); }

    void init() {
        for (int i = 0; i < kNumFields; i++) {
            const FieldDef* f = &kFields[i];
            defs[f->name] = f;
            for (int l = 1; l <= f->levels; l++) {
                if (f->type == REAL) rf[key(f->name, l)].assign(count(f), 0.0);
                else nf[key(f->name, l)].assign(count(f), 0);
            }
        }
    }
    real* rp(const char* name, int lev = 1) { return rf.at(key(name, lev)).data(); }
    A1 r1(const char* name) { return A1{rp(name)}; }
    A2 r2(const char* name, int lev = 1) { return A2{rp(name, lev), inner1(defs.at(name)->inner)}; }
    A3 r3(const char* name, int lev = 1) { const FieldDef* f = defs.at(name); return A3{rp(name, lev), inner1(f->inner), inner2(f->inner)}; }
    I1 i1(const char* name) { return I1{nf.at(name).data()}; }
    I2 i2(const char* name) { return I2{nf.at(name).data(), inner1(defs.at(name)->inner)}; }
};

static inline real sign1(real x) { return std::copysign(1.0, x); }   // Fortran sign(1.0, x)

// statement functions TI:5156-5161 (identical at 3634-3639, 4091-4096)
static inline real flux4(real q_im2, real q_im1, real q_i, real q_ip1, real ua) {
    return ua * (7. * (q_i + q_im1) - (q_ip1 + q_im2)) / 12.0;
}
static inline real flux3(real q_im2, real q_im1, real q_i, real q_ip1, real ua, real coef3) {
    return flux4(q_im2, q_im1, q_i, q_ip1, ua) + coef3 * std::fabs(ua) * ((q_ip1 - q_im2) - 3. * (q_i - q_im1)) / 12.0;
}

// ============================================================ TI:1930-2039
static void atm_rk_integration_setup(Block& b) {
    const int nVertLevels = b.d.nVertLevels, num_scalars = b.d.num_scalars;
    const int cellStart = 1, cellEnd = b.d.nCells, edgeStart = 1, edgeEnd = b.d.nEdges;
    A2 ru = b.r2("ru"), ru_save = b.r2("ru_save"), rw = b.r2("rw"), rw_save = b.r2("rw_save");
    A2 rtheta_p = b.r2("rtheta_p"), rtheta_p_save = b.r2("rtheta_p_save"), rho_p = b.r2("rho_p"), rho_p_save = b.r2("rho_p_save");
    A2 rho_zz_old_split = b.r2("rho_zz_old_split");
    A2 u_1 = b.r2("u", 1), u_2 = b.r2("u", 2), w_1 = b.r2("w", 1), w_2 = b.r2("w", 2);
    A2 theta_m_1 = b.r2("theta_m", 1), theta_m_2 = b.r2("theta_m", 2), rho_zz_1 = b.r2("rho_zz", 1), rho_zz_2 = b.r2("rho_zz", 2);
    A3 scalars_1 = b.r3("scalars", 1), scalars_2 = b.r3("scalars", 2);
    for (int k = 1; k <= nVertLevels; k++) theta_m_2(k, cellEnd + 1) = 0.0;                 // TI:1987
    #pragma omp parallel for
    for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++)
        for (int k = 1; k <= nVertLevels; k++) { ru_save(k, iEdge) = ru(k, iEdge); u_2(k, iEdge) = u_1(k, iEdge); }
    #pragma omp parallel for
    for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
        for (int k = 1; k <= nVertLevels; k++) {
            rtheta_p_save(k, iCell) = rtheta_p(k, iCell);
            rho_p_save(k, iCell) = rho_p(k, iCell);
            theta_m_2(k, iCell) = theta_m_1(k, iCell);
            rho_zz_2(k, iCell) = rho_zz_1(k, iCell);
            rho_zz_old_split(k, iCell) = rho_zz_1(k, iCell);
        }
        for (int k = 1; k <= nVertLevels + 1; k++) { rw_save(k, iCell) = rw(k, iCell); w_2(k, iCell) = w_1(k, iCell); }
        for (int k = 1; k <= nVertLevels; k++)
            for (int j = 1; j <= num_scalars; j++) scalars_2(j, k, iCell) = scalars_1(j, k, iCell);
    }
}

// ============================================================ TI:2042-2146
static void atm_compute_moist_coefficients(Block& b) {
    const int nVertLevels = b.d.nVertLevels, nCellsSolve = b.d.nCellsSolve;
    const int cellStart = 1, cellEnd = b.d.nCells, edgeStart = 1, edgeEnd = b.d.nEdges;
    const int moist_start = b.d.moist_start, moist_end = b.d.moist_end;
    I2 cellsOnEdge = b.i2("cellsOnEdge");
    A3 scalars = b.r3("scalars", 2);
    A2 cqw = b.r2("cqw"), cqu = b.r2("cqu"), qtot = b.r2("qtot");
    #pragma omp parallel for
    for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
        for (int k = 1; k <= nVertLevels; k++) {
            qtot(k, iCell) = 0.0;
            for (int iq = moist_start; iq <= moist_end; iq++) qtot(k, iCell) = qtot(k, iCell) + scalars(iq, k, iCell);
        }
        for (int k = 2; k <= nVertLevels; k++) {
            real qtotal = 0.5 * (qtot(k, iCell) + qtot(k - 1, iCell));
            cqw(k, iCell) = 1.0 / (1.0 + qtotal);
        }
    }
    #pragma omp parallel for
    for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++) {
        int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
        if (cell1 <= nCellsSolve || cell2 <= nCellsSolve) {
            for (int k = 1; k <= nVertLevels; k++) {
                real qtotal = 0.0;
                for (int iq = moist_start; iq <= moist_end; iq++) qtotal = qtotal + 0.5 * (scalars(iq, k, cell1) + scalars(iq, k, cell2));
                cqu(k, iEdge) = 1.0 / (1.0 + qtotal);
            }
        }
    }
}

// ============================================================ TI:2225-2366
static void atm_compute_vert_imp_coefs(Block& b, real dts) {
    const int nVertLevels = b.d.nVertLevels;
    const int cellSolveStart = 1, cellSolveEnd = b.d.nCellsSolve;
    const real epssm = b.c.config_epssm;
    A2 zz = b.r2("zz"), cqw = b.r2("cqw"), p = b.r2("exner"), t = b.r2("theta_m", 2), rb = b.r2("rho_base"),
       rtb = b.r2("rtheta_base"), pb = b.r2("exner_base"), rt = b.r2("rtheta_p"), qtot = b.r2("qtot");
    A2 cofwr = b.r2("cofwr"), cofwz = b.r2("cofwz"), coftz = b.r2("coftz"), cofwt = b.r2("cofwt"),
       a_tri = b.r2("a_tri"), alpha_tri = b.r2("alpha_tri"), gamma_tri = b.r2("gamma_tri");
    A1 cofrz = b.r1("cofrz"), rdzw = b.r1("rdzw"), fzm = b.r1("fzm"), fzp = b.r1("fzp"), rdzu = b.r1("rdzu");
    const real dtseps = .5 * dts * (1. + epssm);
    const real rcv = rgas / (cp - rgas);
    const real c2 = cp * rcv;
    for (int k = 1; k <= nVertLevels; k++) cofrz(k) = dtseps * rdzw(k);
    #pragma omp parallel for
    for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++) {
        std::vector<real> b_tri_v(nVertLevels + 1), c_tri_v(nVertLevels + 1);
        A1 b_tri{b_tri_v.data()}, c_tri{c_tri_v.data()};
        for (int k = 2; k <= nVertLevels; k++)
            cofwr(k, iCell) = .5 * dtseps * gravity * (fzm(k) * zz(k, iCell) + fzp(k) * zz(k - 1, iCell));
        coftz(1, iCell) = 0.0;
        for (int k = 2; k <= nVertLevels; k++) {
            cofwz(k, iCell) = dtseps * c2 * (fzm(k) * zz(k, iCell) + fzp(k) * zz(k - 1, iCell))
                              * rdzu(k) * cqw(k, iCell) * (fzm(k) * p(k, iCell) + fzp(k) * p(k - 1, iCell));
            coftz(k, iCell) = dtseps * (fzm(k) * t(k, iCell) + fzp(k) * t(k - 1, iCell));
        }
        coftz(nVertLevels + 1, iCell) = 0.0;
        for (int k = 1; k <= nVertLevels; k++) {
            real qtotal = qtot(k, iCell);
            cofwt(k, iCell) = .5 * dtseps * rcv * zz(k, iCell) * gravity * rb(k, iCell) / (1. + qtotal)
                              * p(k, iCell) / ((rtb(k, iCell) + rt(k, iCell)) * pb(k, iCell));
        }
        a_tri(1, iCell) = 0.;
        b_tri(1) = 1.;
        c_tri(1) = 0.;
        gamma_tri(1, iCell) = 0.;
        alpha_tri(1, iCell) = 0.;
        for (int k = 2; k <= nVertLevels; k++) {
            a_tri(k, iCell) = -cofwz(k, iCell) * coftz(k - 1, iCell) * rdzw(k - 1) * zz(k - 1, iCell)
                              + cofwr(k, iCell) * cofrz(k - 1)
                              - cofwt(k - 1, iCell) * coftz(k - 1, iCell) * rdzw(k - 1);
            b_tri(k) = 1.
                       + cofwz(k, iCell) * (coftz(k, iCell) * rdzw(k) * zz(k, iCell)
                                            + coftz(k, iCell) * rdzw(k - 1) * zz(k - 1, iCell))
                       - coftz(k, iCell) * (cofwt(k, iCell) * rdzw(k)
                                            - cofwt(k - 1, iCell) * rdzw(k - 1))
                       + cofwr(k, iCell) * (cofrz(k) - cofrz(k - 1));
            c_tri(k) = -cofwz(k, iCell) * coftz(k + 1, iCell) * rdzw(k) * zz(k, iCell)
                       - cofwr(k, iCell) * cofrz(k)
                       + cofwt(k, iCell) * coftz(k + 1, iCell) * rdzw(k);
        }
        for (int k = 2; k <= nVertLevels; k++) {
            alpha_tri(k, iCell) = 1. / (b_tri(k) - a_tri(k, iCell) * gamma_tri(k - 1, iCell));
            gamma_tri(k, iCell) = c_tri(k) * alpha_tri(k, iCell);
        }
    }
}

// ============================================================ TI:2427-2508
static void atm_set_smlstep_pert_variables(Block& b) {
    const int nVertLevels = b.d.nVertLevels;
    const int cellSolveStart = 1, cellSolveEnd = b.d.nCellsSolve;
    I1 nEdgesOnCell = b.i1("nEdgesOnCell");
    I2 edgesOnCell = b.i2("edgesOnCell");
    A2 edgesOnCell_sign = b.r2("edgesOnCell_sign");
    A1 fzm = b.r1("fzm"), fzp = b.r1("fzp");
    A3 zb_cell = b.r3("zb_cell"), zb3_cell = b.r3("zb3_cell");
    A2 zz = b.r2("zz"), w_tend = b.r2("tend_w"), u_tend = b.r2("tend_u");
    #pragma omp parallel for
    for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++) {
        for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
            int iEdge = edgesOnCell(i, iCell);
            for (int k = 2; k <= nVertLevels; k++) {
                real flux = edgesOnCell_sign(i, iCell) * (fzm(k) * u_tend(k, iEdge) + fzp(k) * u_tend(k - 1, iEdge));
                w_tend(k, iCell) = w_tend(k, iCell)
                                   - (zb_cell(k, i, iCell) + sign1(u_tend(k, iEdge)) * zb3_cell(k, i, iCell)) * flux;
            }
        }
        for (int k = 2; k <= nVertLevels; k++)
            w_tend(k, iCell) = (fzm(k) * zz(k, iCell) + fzp(k) * zz(k - 1, iCell)) * w_tend(k, iCell);
    }
}

// ============================================================ TI:2646-2984
static void atm_advance_acoustic_step(Block& b, real dts, int small_step) {
    const int nVertLevels = b.d.nVertLevels, nCellsSolve = b.d.nCellsSolve;
    const int cellStart = 1, cellEnd = b.d.nCells, edgeStart = 1, edgeEnd = b.d.nEdges;
    const int cellSolveStart = 1, cellSolveEnd = b.d.nCellsSolve;
    const real epssm = b.c.config_epssm;
    A2 rho_zz = b.r2("rho_zz", 2), theta_m = b.r2("theta_m", 1), w = b.r2("w", 2);      // TI:2573-2577
    A2 ru_p = b.r2("ru_p"), rw_p = b.r2("rw_p"), rtheta_pp = b.r2("rtheta_pp"), rtheta_pp_old = b.r2("rtheta_pp_old");
    A2 zz = b.r2("zz"), exner = b.r2("exner"), cqu = b.r2("cqu"), ruAvg = b.r2("ruAvg"), wwAvg = b.r2("wwAvg");
    A2 rho_pp = b.r2("rho_pp"), cofwt = b.r2("cofwt"), coftz = b.r2("coftz"), zxu = b.r2("zxu");
    A2 a_tri = b.r2("a_tri"), alpha_tri = b.r2("alpha_tri"), gamma_tri = b.r2("gamma_tri"), dss = b.r2("dss");
    A2 tend_ru = b.r2("tend_u"), tend_rho = b.r2("tend_rho"), tend_rt = b.r2("tend_theta"), tend_rw = b.r2("tend_w");
    A2 cofwr = b.r2("cofwr"), cofwz = b.r2("cofwz"), rw = b.r2("rw"), rw_save = b.r2("rw_save");
    A1 fzm = b.r1("fzm"), fzp = b.r1("fzp"), rdzw = b.r1("rdzw"), invDcEdge = b.r1("invDcEdge"),
       invAreaCell = b.r1("invAreaCell"), cofrz = b.r1("cofrz"), dvEdge = b.r1("dvEdge");
    A1 specZoneMaskCell = b.r1("specZoneMaskCell"), specZoneMaskEdge = b.r1("specZoneMaskEdge");
    I1 nEdgesOnCell = b.i1("nEdgesOnCell");
    I2 cellsOnEdge = b.i2("cellsOnEdge"), edgesOnCell = b.i2("edgesOnCell");
    A2 edgesOnCell_sign = b.r2("edgesOnCell_sign");

    const real rcv = rgas / (cp - rgas);
    const real c2 = cp * rcv;
    const real resm = (1.0 - epssm) / (1.0 + epssm);

    if (small_step != 1) {
        #pragma omp parallel for
        for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++) {
            int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
            if (cell1 <= nCellsSolve || cell2 <= nCellsSolve) {
                for (int k = 1; k <= nVertLevels; k++) {
                    real pgrad = ((rtheta_pp(k, cell2) - rtheta_pp(k, cell1)) * invDcEdge(iEdge)) / (.5 * (zz(k, cell2) + zz(k, cell1)));
                    pgrad = cqu(k, iEdge) * 0.5 * c2 * (exner(k, cell1) + exner(k, cell2)) * pgrad;
                    pgrad = pgrad + 0.5 * zxu(k, iEdge) * gravity * (rho_pp(k, cell1) + rho_pp(k, cell2));
                    ru_p(k, iEdge) = ru_p(k, iEdge) + dts * (tend_ru(k, iEdge) - (1.0 - specZoneMaskEdge(iEdge)) * pgrad);
                }
                for (int k = 1; k <= nVertLevels; k++) ruAvg(k, iEdge) = ruAvg(k, iEdge) + ru_p(k, iEdge);
            }
        }
    } else {
        #pragma omp parallel for
        for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++) {
            int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
            if (cell1 <= nCellsSolve || cell2 <= nCellsSolve) {
                for (int k = 1; k <= nVertLevels; k++) ru_p(k, iEdge) = dts * tend_ru(k, iEdge);
                for (int k = 1; k <= nVertLevels; k++) ruAvg(k, iEdge) = ru_p(k, iEdge);
            }
        }
    }
    if (small_step == 1) {
        #pragma omp parallel for
        for (int iCell = cellStart; iCell <= cellEnd; iCell++)
            for (int k = 1; k <= nVertLevels; k++) rtheta_pp_old(k, iCell) = 0.0;
    } else {
        #pragma omp parallel for
        for (int iCell = cellStart; iCell <= cellEnd; iCell++)
            for (int k = 1; k <= nVertLevels; k++) rtheta_pp_old(k, iCell) = rtheta_pp(k, iCell);
    }
    // !$OMP BARRIER  TI:2844
    #pragma omp parallel for
    for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++) {
        std::vector<real> ts_v(nVertLevels + 1), rs_v(nVertLevels + 1);
        A1 ts{ts_v.data()}, rs{rs_v.data()};
        if (small_step == 1) {
            for (int k = 1; k <= nVertLevels; k++) {
                wwAvg(k, iCell) = 0.0; rho_pp(k, iCell) = 0.0; rtheta_pp(k, iCell) = 0.0; rw_p(k, iCell) = 0.0;
            }
            wwAvg(nVertLevels + 1, iCell) = 0.0;
            rw_p(nVertLevels + 1, iCell) = 0.0;
        }
        if (specZoneMaskCell(iCell) == 0.0) {
            for (int k = 1; k <= nVertLevels; k++) { ts(k) = 0.0; rs(k) = 0.0; }
            for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
                int iEdge = edgesOnCell(i, iCell);
                int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
                for (int k = 1; k <= nVertLevels; k++) {
                    real flux = edgesOnCell_sign(i, iCell) * dts * dvEdge(iEdge) * ru_p(k, iEdge) * invAreaCell(iCell);
                    rs(k) = rs(k) - flux;
                    ts(k) = ts(k) - flux * 0.5 * (theta_m(k, cell2) + theta_m(k, cell1));
                }
            }
            for (int k = 1; k <= nVertLevels; k++) {
                rs(k) = rho_pp(k, iCell) + dts * tend_rho(k, iCell) + rs(k)
                        - cofrz(k) * resm * (rw_p(k + 1, iCell) - rw_p(k, iCell));
                ts(k) = rtheta_pp(k, iCell) + dts * tend_rt(k, iCell) + ts(k)
                        - resm * rdzw(k) * (coftz(k + 1, iCell) * rw_p(k + 1, iCell)
                                            - coftz(k, iCell) * rw_p(k, iCell));
            }
            for (int k = 2; k <= nVertLevels; k++) wwAvg(k, iCell) = wwAvg(k, iCell) + 0.5 * (1.0 - epssm) * rw_p(k, iCell);
            for (int k = 2; k <= nVertLevels; k++) {
                rw_p(k, iCell) = rw_p(k, iCell) + dts * tend_rw(k, iCell)
                                 - cofwz(k, iCell) * ((zz(k, iCell) * ts(k)
                                                       - zz(k - 1, iCell) * ts(k - 1))
                                                      + resm * (zz(k, iCell) * rtheta_pp(k, iCell)
                                                                - zz(k - 1, iCell) * rtheta_pp(k - 1, iCell)))
                                 - cofwr(k, iCell) * ((rs(k) + rs(k - 1))
                                                      + resm * (rho_pp(k, iCell) + rho_pp(k - 1, iCell)))
                                 + cofwt(k, iCell) * (ts(k) + resm * rtheta_pp(k, iCell))
                                 + cofwt(k - 1, iCell) * (ts(k - 1) + resm * rtheta_pp(k - 1, iCell));
            }
            for (int k = 2; k <= nVertLevels; k++)
                rw_p(k, iCell) = (rw_p(k, iCell) - a_tri(k, iCell) * rw_p(k - 1, iCell)) * alpha_tri(k, iCell);
            for (int k = nVertLevels; k >= 1; k--)
                rw_p(k, iCell) = rw_p(k, iCell) - gamma_tri(k, iCell) * rw_p(k + 1, iCell);
            for (int k = 2; k <= nVertLevels; k++) {
                rw_p(k, iCell) = (rw_p(k, iCell) + (rw_save(k, iCell) - rw(k, iCell)) - dts * dss(k, iCell) *
                                  (fzm(k) * zz(k, iCell) + fzp(k) * zz(k - 1, iCell))
                                  * (fzm(k) * rho_zz(k, iCell) + fzp(k) * rho_zz(k - 1, iCell))
                                  * w(k, iCell)) / (1.0 + dts * dss(k, iCell))
                                 - (rw_save(k, iCell) - rw(k, iCell));
            }
            for (int k = 2; k <= nVertLevels; k++) wwAvg(k, iCell) = wwAvg(k, iCell) + 0.5 * (1.0 + epssm) * rw_p(k, iCell);
            for (int k = 1; k <= nVertLevels; k++) {
                rho_pp(k, iCell) = rs(k) - cofrz(k) * (rw_p(k + 1, iCell) - rw_p(k, iCell));
                rtheta_pp(k, iCell) = ts(k) - rdzw(k) * (coftz(k + 1, iCell) * rw_p(k + 1, iCell)
                                                         - coftz(k, iCell) * rw_p(k, iCell));
            }
        } else {
            for (int k = 1; k <= nVertLevels; k++) {
                rho_pp(k, iCell) = rho_pp(k, iCell) + dts * tend_rho(k, iCell);
                rtheta_pp(k, iCell) = rtheta_pp(k, iCell) + dts * tend_rt(k, iCell);
                rw_p(k, iCell) = rw_p(k, iCell) + dts * tend_rw(k, iCell);
                wwAvg(k, iCell) = wwAvg(k, iCell) + 0.5 * (1.0 + epssm) * rw_p(k, iCell);
            }
        }
    }
}

// ============================================================ TI:2987-3075
static void atm_divergence_damping_3d(Block& b, real dts) {
    const int nVertLevels = b.d.nVertLevels, nCellsSolve = b.d.nCellsSolve;
    const int edgeStart = 1, edgeEnd = b.d.nEdges;
    A2 theta_m = b.r2("theta_m", 1), ru_p = b.r2("ru_p"), rtheta_pp = b.r2("rtheta_pp"), rtheta_pp_old = b.r2("rtheta_pp_old");
    A1 specZoneMaskEdge = b.r1("specZoneMaskEdge");
    I2 cellsOnEdge = b.i2("cellsOnEdge");
    const real smdiv = b.c.config_smdiv, config_len_disp = b.c.config_len_disp;
    const real rdts = 1.0 / dts;
    const real coef_divdamp = 2.0 * smdiv * config_len_disp * rdts;
    #pragma omp parallel for
    for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++) {
        int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
        if (cell1 <= nCellsSolve || cell2 <= nCellsSolve) {
            for (int k = 1; k <= nVertLevels; k++) {
                real divCell1 = -(rtheta_pp(k, cell1) - rtheta_pp_old(k, cell1));
                real divCell2 = -(rtheta_pp(k, cell2) - rtheta_pp_old(k, cell2));
                ru_p(k, iEdge) = ru_p(k, iEdge) + coef_divdamp * (divCell2 - divCell1) * (1.0 - specZoneMaskEdge(iEdge))
                                                  / (theta_m(k, cell1) + theta_m(k, cell2));
            }
        }
    }
}

// ============================================================ TI:3189-3431
static void atm_recover_large_step_variables(Block& b, real dt, int ns, int rk_step) {
    const int nVertLevels = b.d.nVertLevels, nCells = b.d.nCells;
    const int cellStart = 1, cellEnd = b.d.nCells, edgeStart = 1, edgeEnd = b.d.nEdges;
    A2 wwAvg = b.r2("wwAvg"), rw_save = b.r2("rw_save"), w = b.r2("w", 2), rw = b.r2("rw"), rw_p = b.r2("rw_p");
    A2 rtheta_p = b.r2("rtheta_p"), rtheta_pp = b.r2("rtheta_pp"), rtheta_p_save = b.r2("rtheta_p_save"),
       rt_diabatic_tend = b.r2("rt_diabatic_tend"), rho_p = b.r2("rho_p"), rho_p_save = b.r2("rho_p_save"),
       rho_pp = b.r2("rho_pp"), rho_zz = b.r2("rho_zz", 2), rho_base = b.r2("rho_base");
    A2 ruAvg = b.r2("ruAvg"), ru_save = b.r2("ru_save"), ru_p = b.r2("ru_p"), u = b.r2("u", 2), ru = b.r2("ru");
    A2 exner = b.r2("exner"), exner_base = b.r2("exner_base"), rtheta_base = b.r2("rtheta_base"),
       pressure_p = b.r2("pressure_p"), zz = b.r2("zz"), theta_m = b.r2("theta_m", 2);
    A1 fzm = b.r1("fzm"), fzp = b.r1("fzp");
    A3 zb_cell = b.r3("zb_cell"), zb3_cell = b.r3("zb3_cell");
    A2 edgesOnCell_sign = b.r2("edgesOnCell_sign");
    I2 cellsOnEdge = b.i2("cellsOnEdge"), edgesOnCell = b.i2("edgesOnCell");
    I1 nEdgesOnCell = b.i1("nEdgesOnCell");
    const real cf1 = b.c.cf1, cf2 = b.c.cf2, cf3 = b.c.cf3;
    const real rcv = rgas / (cp - rgas);
    const real p0 = 1.0e+05;
    for (int k = 1; k <= nVertLevels; k++) rho_zz(k, nCells + 1) = 1.0;                    // TI:3282-3284
    const real invNs = 1 / (real)ns;
    #pragma omp parallel for
    for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
        for (int k = 1; k <= nVertLevels; k++) {
            rho_p(k, iCell) = rho_p_save(k, iCell) + rho_pp(k, iCell);
            rho_zz(k, iCell) = rho_p(k, iCell) + rho_base(k, iCell);
        }
        rw(1, iCell) = 0.0;
        w(1, iCell) = 0.0;
        for (int k = 2; k <= nVertLevels; k++) {
            wwAvg(k, iCell) = rw_save(k, iCell) + (wwAvg(k, iCell) * invNs);
            rw(k, iCell) = rw_save(k, iCell) + rw_p(k, iCell);
            w(k, iCell) = rw(k, iCell) / (fzm(k) * zz(k, iCell) + fzp(k) * zz(k - 1, iCell));
        }
        rw(nVertLevels + 1, iCell) = 0.0;
        w(nVertLevels + 1, iCell) = 0.0;
    }
    if (rk_step == 3) {
        #pragma omp parallel for
        for (int iCell = cellStart; iCell <= cellEnd; iCell++)
            for (int k = 1; k <= nVertLevels; k++) {
                rtheta_p(k, iCell) = rtheta_p_save(k, iCell) + rtheta_pp(k, iCell)
                                     - dt * rho_zz(k, iCell) * rt_diabatic_tend(k, iCell);
                theta_m(k, iCell) = (rtheta_p(k, iCell) + rtheta_base(k, iCell)) / rho_zz(k, iCell);
                exner(k, iCell) = std::pow(zz(k, iCell) * (rgas / p0) * (rtheta_p(k, iCell) + rtheta_base(k, iCell)), rcv);
                pressure_p(k, iCell) = zz(k, iCell) * rgas * (exner(k, iCell) * rtheta_p(k, iCell) + rtheta_base(k, iCell)
                                                               * (exner(k, iCell) - exner_base(k, iCell)));
            }
    } else {
        #pragma omp parallel for
        for (int iCell = cellStart; iCell <= cellEnd; iCell++)
            for (int k = 1; k <= nVertLevels; k++) {
                rtheta_p(k, iCell) = rtheta_p_save(k, iCell) + rtheta_pp(k, iCell);
                theta_m(k, iCell) = (rtheta_p(k, iCell) + rtheta_base(k, iCell)) / rho_zz(k, iCell);
            }
    }
    // !$OMP BARRIER TI:3356
    #pragma omp parallel for
    for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++) {
        int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
        for (int k = 1; k <= nVertLevels; k++) {
            ruAvg(k, iEdge) = ru_save(k, iEdge) + (ruAvg(k, iEdge) * invNs);
            ru(k, iEdge) = ru_save(k, iEdge) + ru_p(k, iEdge);
            u(k, iEdge) = 2. * ru(k, iEdge) / (rho_zz(k, cell1) + rho_zz(k, cell2));
        }
    }
    // !$OMP BARRIER TI:3375
    #pragma omp parallel for
    for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
        for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
            int iEdge = edgesOnCell(i, iCell);
            real flux = (cf1 * ru(1, iEdge) + cf2 * ru(2, iEdge) + cf3 * ru(3, iEdge));
            w(1, iCell) = w(1, iCell) + edgesOnCell_sign(i, iCell) *
                          (zb_cell(1, i, iCell) + sign1(flux) * zb3_cell(1, i, iCell)) * flux;
            for (int k = 2; k <= nVertLevels; k++) {
                flux = (fzm(k) * ru(k, iEdge) + fzp(k) * ru(k - 1, iEdge));
                w(k, iCell) = w(k, iCell) + edgesOnCell_sign(i, iCell) *
                              (zb_cell(k, i, iCell) + sign1(flux) * zb3_cell(k, i, iCell)) * flux;
            }
        }
        w(1, iCell) = w(1, iCell) / (cf1 * rho_zz(1, iCell) + cf2 * rho_zz(2, iCell) + cf3 * rho_zz(3, iCell));
        for (int k = 2; k <= nVertLevels; k++)
            w(k, iCell) = w(k, iCell) / (fzm(k) * rho_zz(k, iCell) + fzp(k) * rho_zz(k - 1, iCell));
    }
}

// ============================================================ TI:3575-3855
static void atm_advance_scalars(Block& b, real dt, int rk_step) {
    const int nVertLevels = b.d.nVertLevels, num_scalars = b.d.num_scalars;
    const int edgeStart = 1, edgeEnd = b.d.nEdges, cellSolveStart = 1, cellSolveEnd = b.d.nCellsSolve;
    const int config_time_integration_order = b.c.config_time_integration_order;
    const bool advance_density = b.c.config_split_dynamics_transport != 0;
    const real coef_3rd_order = b.c.config_coef_3rd_order;
    A3 scalar_old = b.r3("scalars", 1), scalar_new = b.r3("scalars", 2), scalar_tend_save = b.r3("scalars_tend"),
       horiz_flux_arr = b.r3("horiz_flux_arr");
    A2 rho_zz_old = b.r2("rho_zz", 1), rho_zz_new = b.r2("rho_zz", 2), uhAvg = b.r2("ruAvg"), wwAvg = b.r2("wwAvg");
    A1 invAreaCell = b.r1("invAreaCell"), fnm = b.r1("fzm"), fnp = b.r1("fzp"), rdnw = b.r1("rdzw");
    I2 edgesOnCell = b.i2("edgesOnCell"), advCellsForEdge = b.i2("advCellsForEdge");
    I1 nEdgesOnCell = b.i1("nEdgesOnCell"), nAdvCellsForEdge = b.i1("nAdvCellsForEdge");
    A2 edgesOnCell_sign = b.r2("edgesOnCell_sign"), adv_coefs = b.r2("adv_coefs"), adv_coefs_3rd = b.r2("adv_coefs_3rd");

    real weight_time_new = 1.;
    if (!advance_density) {
        weight_time_new = 1.;
    } else {
        if ((rk_step == 1) && config_time_integration_order == 3) weight_time_new = 1. / 3;
        if ((rk_step == 1) && config_time_integration_order == 2) weight_time_new = 1. / 2;
        if (rk_step == 2) weight_time_new = 1. / 2;
        if (rk_step == 3) weight_time_new = 1.;
    }
    const real weight_time_old = 1. - weight_time_new;

    #pragma omp parallel for
    for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++) {
        if (nAdvCellsForEdge(iEdge) == 10) {
            std::vector<real> sw2((size_t)nVertLevels * 10);
            A2 scalar_weight2{sw2.data(), nVertLevels};
            int ica[11];
            for (int j = 1; j <= 10; j++)
                for (int k = 1; k <= nVertLevels; k++)
                    scalar_weight2(k, j) = adv_coefs(j, iEdge) + sign1(uhAvg(k, iEdge)) * adv_coefs_3rd(j, iEdge);
            for (int j = 1; j <= 10; j++) ica[j] = advCellsForEdge(j, iEdge);
            for (int k = 1; k <= nVertLevels; k++)
                for (int iScalar = 1; iScalar <= num_scalars; iScalar++)
                    horiz_flux_arr(iScalar, k, iEdge) =
                        scalar_weight2(k, 1) * scalar_new(iScalar, k, ica[1]) +
                        scalar_weight2(k, 2) * scalar_new(iScalar, k, ica[2]) +
                        scalar_weight2(k, 3) * scalar_new(iScalar, k, ica[3]) +
                        scalar_weight2(k, 4) * scalar_new(iScalar, k, ica[4]) +
                        scalar_weight2(k, 5) * scalar_new(iScalar, k, ica[5]) +
                        scalar_weight2(k, 6) * scalar_new(iScalar, k, ica[6]) +
                        scalar_weight2(k, 7) * scalar_new(iScalar, k, ica[7]) +
                        scalar_weight2(k, 8) * scalar_new(iScalar, k, ica[8]) +
                        scalar_weight2(k, 9) * scalar_new(iScalar, k, ica[9]) +
                        scalar_weight2(k, 10) * scalar_new(iScalar, k, ica[10]);
        } else {
            for (int k = 1; k <= nVertLevels; k++)
                for (int iScalar = 1; iScalar <= num_scalars; iScalar++) horiz_flux_arr(iScalar, k, iEdge) = 0.0;
            for (int j = 1; j <= nAdvCellsForEdge(iEdge); j++) {
                int iAdvCell = advCellsForEdge(j, iEdge);
                for (int k = 1; k <= nVertLevels; k++)
                    for (int iScalar = 1; iScalar <= num_scalars; iScalar++) {
                        real scalar_weight = adv_coefs(j, iEdge) + sign1(uhAvg(k, iEdge)) * adv_coefs_3rd(j, iEdge);
                        horiz_flux_arr(iScalar, k, iEdge) = horiz_flux_arr(iScalar, k, iEdge)
                                                            + scalar_weight * scalar_new(iScalar, k, iAdvCell);
                    }
            }
        }
    }
    // !$OMP BARRIER TI:3754
    #pragma omp parallel for
    for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++) {
        std::vector<real> stc((size_t)num_scalars * nVertLevels), wd((size_t)num_scalars * (nVertLevels + 1));
        A2 scalar_tend_column{stc.data(), num_scalars}, wdtn{wd.data(), num_scalars};
        for (int k = 1; k <= nVertLevels; k++)
            for (int iScalar = 1; iScalar <= num_scalars; iScalar++) {
                scalar_tend_column(iScalar, k) = 0.0;
                scalar_tend_save(iScalar, k, iCell) = 0.0;          // #ifndef DO_PHYSICS, TI:3781-3783
            }
        for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
            int iEdge = edgesOnCell(i, iCell);
            for (int k = 1; k <= nVertLevels; k++)
                for (int iScalar = 1; iScalar <= num_scalars; iScalar++)
                    scalar_tend_column(iScalar, k) = scalar_tend_column(iScalar, k)
                        - edgesOnCell_sign(i, iCell) * uhAvg(k, iEdge) * horiz_flux_arr(iScalar, k, iEdge);
        }
        for (int k = 1; k <= nVertLevels; k++)
            for (int iScalar = 1; iScalar <= num_scalars; iScalar++)
                scalar_tend_column(iScalar, k) = scalar_tend_column(iScalar, k) * invAreaCell(iCell)
                                                 + scalar_tend_save(iScalar, k, iCell);
        for (int iScalar = 1; iScalar <= num_scalars; iScalar++) {
            wdtn(iScalar, 1) = 0.0;
            wdtn(iScalar, 2) = wwAvg(2, iCell) * (fnm(2) * scalar_new(iScalar, 2, iCell) + fnp(2) * scalar_new(iScalar, 2 - 1, iCell));
            wdtn(iScalar, nVertLevels) = wwAvg(nVertLevels, iCell) *
                                         (fnm(nVertLevels) * scalar_new(iScalar, nVertLevels, iCell)
                                          + fnp(nVertLevels) * scalar_new(iScalar, nVertLevels - 1, iCell));
            wdtn(iScalar, nVertLevels + 1) = 0.0;
        }
        for (int k = 3; k <= nVertLevels - 1; k++)
            for (int iScalar = 1; iScalar <= num_scalars; iScalar++)
                wdtn(iScalar, k) = flux3(scalar_new(iScalar, k - 2, iCell), scalar_new(iScalar, k - 1, iCell),
                                         scalar_new(iScalar, k, iCell), scalar_new(iScalar, k + 1, iCell),
                                         wwAvg(k, iCell), coef_3rd_order);
        for (int k = 1; k <= nVertLevels; k++)
            for (int iScalar = 1; iScalar <= num_scalars; iScalar++) {
                real rho_zz_new_inv = 1.0 / (weight_time_old * rho_zz_old(k, iCell) + weight_time_new * rho_zz_new(k, iCell));
                scalar_new(iScalar, k, iCell) = (scalar_old(iScalar, k, iCell) * rho_zz_old(k, iCell)
                    + dt * (scalar_tend_column(iScalar, k) - rdnw(k) * (wdtn(iScalar, k + 1) - wdtn(iScalar, k)))) * rho_zz_new_inv;
            }
    }
}

// ============================================================ TI:4012-4734, split at its two halo exchanges
// part A: TI:4129-4143 (then exchange 'dynamics:scalars_old', TI:4155)
static void mono_pre_update(Block& b, real dt) {
    const int nVertLevels = b.d.nVertLevels, num_scalars = b.d.num_scalars;
    A3 scalars_old = b.r3("scalars", 1), scalar_tend = b.r3("scalars_tend");
    A2 rho_zz_old = b.r2("rho_zz", 1);
    #pragma omp parallel for
    for (int iCell = 1; iCell <= b.d.nCellsSolve; iCell++)
        for (int k = 1; k <= nVertLevels; k++)
            for (int iScalar = 1; iScalar <= num_scalars; iScalar++) {
                scalar_tend(iScalar, k, iCell) = 0.0;                     // #ifndef DO_PHYSICS TI:4135-4137
                scalars_old(iScalar, k, iCell) = scalars_old(iScalar, k, iCell) + dt * scalar_tend(iScalar, k, iCell) / rho_zz_old(k, iCell);
                scalar_tend(iScalar, k, iCell) = 0.0;
            }
}
// part B: TI:4167-4210 density re-integration
static void mono_rho_zz_int(Block& b, real dt) {
    const int nVertLevels = b.d.nVertLevels;
    if (!b.c.config_split_dynamics_transport) return;
    A2 rho_zz_int = b.r2("rho_zz_int"), uhAvg = b.r2("ruAvg"), wwAvg = b.r2("wwAvg"), rho_zz_old = b.r2("rho_zz", 1);
    A2 edgesOnCell_sign = b.r2("edgesOnCell_sign");
    A1 dvEdge = b.r1("dvEdge"), invAreaCell = b.r1("invAreaCell"), rdnw = b.r1("rdzw");
    I2 edgesOnCell = b.i2("edgesOnCell");
    I1 nEdgesOnCell = b.i1("nEdgesOnCell");
    #pragma omp parallel for
    for (int iCell = 1; iCell <= b.d.nCellsSolve; iCell++) {
        for (int k = 1; k <= nVertLevels; k++) rho_zz_int(k, iCell) = 0.0;
        for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
            int iEdge = edgesOnCell(i, iCell);
            for (int k = 1; k <= nVertLevels; k++)
                rho_zz_int(k, iCell) = rho_zz_int(k, iCell) - edgesOnCell_sign(i, iCell)
                                       * uhAvg(k, iEdge) * dvEdge(iEdge) * invAreaCell(iCell);
        }
        for (int k = 1; k <= nVertLevels; k++)
            rho_zz_int(k, iCell) = rho_zz_old(k, iCell) + dt * (rho_zz_int(k, iCell) - rdnw(k) * (wwAvg(k + 1, iCell) - wwAvg(k, iCell)));
    }
}
// part C: per scalar, TI:4225-4553 (then exchange 'dynamics:scale', TI:4568)
static void mono_scalar_phase1(Block& b, real dt, int iScalar) {
    const int nVertLevels = b.d.nVertLevels, nCells = b.d.nCells, nCellsSolve = b.d.nCellsSolve;
    const int cellStart = 1, cellEnd = nCells, edgeStart = 1, edgeEnd = b.d.nEdges, cellSolveStart = 1, cellSolveEnd = nCellsSolve;
    const bool local_advance_density = b.c.config_split_dynamics_transport != 0;
    const real coef_3rd_order = b.c.config_coef_3rd_order;
    const real eps = 1.e-20;
    const int SCALE_IN = 1, SCALE_OUT = 2;
    A3 scalars_old = b.r3("scalars", 1), scalars_new = b.r3("scalars", 2), scale_arr = b.r3("scale_arr");
    A2 scalar_old = b.r2("scalar_old"), scalar_new = b.r2("scalar_new"), s_max = b.r2("s_max"), s_min = b.r2("s_min"),
       wdtn = b.r2("wdtn"), flux_arr = b.r2("flux_arr"), flux_upwind_tmp = b.r2("flux_upwind_tmp"), flux_tmp = b.r2("flux_tmp"),
       rho_zz_int = b.r2("rho_zz_int"), rho_zz_old = b.r2("rho_zz", 1), rho_zz_new = b.r2("rho_zz", 2),
       uhAvg = b.r2("ruAvg"), wwAvg = b.r2("wwAvg");
    A1 invAreaCell = b.r1("invAreaCell"), dvEdge = b.r1("dvEdge"), fnm = b.r1("fzm"), fnp = b.r1("fzp"), rdnw = b.r1("rdzw");
    I2 cellsOnEdge = b.i2("cellsOnEdge"), cellsOnCell = b.i2("cellsOnCell"), edgesOnCell = b.i2("edgesOnCell"),
       advCellsForEdge = b.i2("advCellsForEdge");
    I1 nEdgesOnCell = b.i1("nEdgesOnCell"), nAdvCellsForEdge = b.i1("nAdvCellsForEdge");
    A2 edgesOnCell_sign = b.r2("edgesOnCell_sign"), adv_coefs = b.r2("adv_coefs"), adv_coefs_3rd = b.r2("adv_coefs_3rd");

    #pragma omp parallel for
    for (int iCell = cellStart; iCell <= cellEnd; iCell++)
        for (int k = 1; k <= nVertLevels; k++) {
            scalar_old(k, iCell) = scalars_old(iScalar, k, iCell);
            scalar_new(k, iCell) = scalars_new(iScalar, k, iCell);
        }
    for (int k = 1; k <= nVertLevels; k++) { scalar_old(k, nCells + 1) = 0.0; scalar_new(k, nCells + 1) = 0.0; }
    // !$OMP BARRIER TI:4243
    #pragma omp parallel for
    for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++) {
        wdtn(1, iCell) = 0.0;
        wdtn(nVertLevels + 1, iCell) = 0.0;
        int k = 1;
        s_max(k, iCell) = std::max(scalar_old(1, iCell), scalar_old(2, iCell));
        s_min(k, iCell) = std::min(scalar_old(1, iCell), scalar_old(2, iCell));
        k = 2;
        wdtn(k, iCell) = wwAvg(k, iCell) * (fnm(k) * scalar_new(k, iCell) + fnp(k) * scalar_new(k - 1, iCell));
        s_max(k, iCell) = std::max(std::max(scalar_old(k - 1, iCell), scalar_old(k, iCell)), scalar_old(k + 1, iCell));
        s_min(k, iCell) = std::min(std::min(scalar_old(k - 1, iCell), scalar_old(k, iCell)), scalar_old(k + 1, iCell));
        for (k = 3; k <= nVertLevels - 1; k++) {
            wdtn(k, iCell) = flux3(scalar_new(k - 2, iCell), scalar_new(k - 1, iCell),
                                   scalar_new(k, iCell), scalar_new(k + 1, iCell),
                                   wwAvg(k, iCell), coef_3rd_order);
            s_max(k, iCell) = std::max(std::max(scalar_old(k - 1, iCell), scalar_old(k, iCell)), scalar_old(k + 1, iCell));
            s_min(k, iCell) = std::min(std::min(scalar_old(k - 1, iCell), scalar_old(k, iCell)), scalar_old(k + 1, iCell));
        }
        k = nVertLevels;
        wdtn(k, iCell) = wwAvg(k, iCell) * (fnm(k) * scalar_new(k, iCell) + fnp(k) * scalar_new(k - 1, iCell));
        s_max(k, iCell) = std::max(scalar_old(k, iCell), scalar_old(k - 1, iCell));
        s_min(k, iCell) = std::min(scalar_old(k, iCell), scalar_old(k - 1, iCell));
        for (int i = 1; i <= nEdgesOnCell(iCell); i++)
            for (k = 1; k <= nVertLevels; k++) {
                s_max(k, iCell) = std::max(s_max(k, iCell), scalar_old(k, cellsOnCell(i, iCell)));
                s_min(k, iCell) = std::min(s_min(k, iCell), scalar_old(k, cellsOnCell(i, iCell)));
            }
    }
    // !$OMP BARRIER TI:4348 -- high-order horizontal flux
    #pragma omp parallel for
    for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++) {
        int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
        if (cell1 <= nCellsSolve || cell2 <= nCellsSolve) {
            if (nAdvCellsForEdge(iEdge) == 10) {
                int ica[11]; real swa[11][3];
                for (int jj = 1; jj <= 10; jj++) {
                    ica[jj] = advCellsForEdge(jj, iEdge);
                    swa[jj][1] = adv_coefs(jj, iEdge) + adv_coefs_3rd(jj, iEdge);
                    swa[jj][2] = adv_coefs(jj, iEdge) - adv_coefs_3rd(jj, iEdge);
                }
                for (int k = 1; k <= nVertLevels; k++) {
                    int ii = (uhAvg(k, iEdge) > 0) ? 1 : 2;
                    flux_arr(k, iEdge) = uhAvg(k, iEdge) * (
                        swa[1][ii] * scalar_new(k, ica[1]) + swa[2][ii] * scalar_new(k, ica[2]) +
                        swa[3][ii] * scalar_new(k, ica[3]) + swa[4][ii] * scalar_new(k, ica[4]) +
                        swa[5][ii] * scalar_new(k, ica[5]) + swa[6][ii] * scalar_new(k, ica[6]) +
                        swa[7][ii] * scalar_new(k, ica[7]) + swa[8][ii] * scalar_new(k, ica[8]) +
                        swa[9][ii] * scalar_new(k, ica[9]) + swa[10][ii] * scalar_new(k, ica[10]));
                }
            } else {
                for (int k = 1; k <= nVertLevels; k++) flux_arr(k, iEdge) = 0.0;
                for (int i = 1; i <= nAdvCellsForEdge(iEdge); i++) {
                    int iCell = advCellsForEdge(i, iEdge);
                    for (int k = 1; k <= nVertLevels; k++) {
                        real scalar_weight = uhAvg(k, iEdge) * (adv_coefs(i, iEdge) + sign1(uhAvg(k, iEdge)) * adv_coefs_3rd(i, iEdge));
                        flux_arr(k, iEdge) = flux_arr(k, iEdge) + scalar_weight * scalar_new(k, iCell);
                    }
                }
            }
        } else {
            for (int k = 1; k <= nVertLevels; k++) flux_arr(k, iEdge) = 0.0;
        }
    }
    // !$OMP BARRIER TI:4417 -- upwind update, vertical part
    #pragma omp parallel for
    for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++) {
        std::vector<real> fu(nVertLevels + 2);
        A1 flux_upwind_arr{fu.data()};
        int k = 1;
        scalar_new(k, iCell) = scalar_old(k, iCell) * rho_zz_old(k, iCell);
        for (k = 2; k <= nVertLevels; k++) {
            scalar_new(k, iCell) = scalar_old(k, iCell) * rho_zz_old(k, iCell);
            flux_upwind_arr(k) = dt * (std::max(0.0, wwAvg(k, iCell)) * scalar_old(k - 1, iCell) + std::min(0.0, wwAvg(k, iCell)) * scalar_old(k, iCell));
        }
        for (k = 1; k <= nVertLevels - 1; k++) scalar_new(k, iCell) = scalar_new(k, iCell) - flux_upwind_arr(k + 1) * rdnw(k);
        for (k = 2; k <= nVertLevels; k++) {
            scalar_new(k, iCell) = scalar_new(k, iCell) + flux_upwind_arr(k) * rdnw(k);
            wdtn(k, iCell) = dt * wdtn(k, iCell) - flux_upwind_arr(k);
        }
        for (k = 1; k <= nVertLevels; k++) {
            scale_arr(k, SCALE_IN, iCell) = -rdnw(k) * (std::min(0.0, wdtn(k + 1, iCell)) - std::max(0.0, wdtn(k, iCell)));
            scale_arr(k, SCALE_OUT, iCell) = -rdnw(k) * (std::max(0.0, wdtn(k + 1, iCell)) - std::min(0.0, wdtn(k, iCell)));
        }
    }
    #pragma omp parallel for
    for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++) {
        int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
        for (int k = 1; k <= nVertLevels; k++) {
            flux_upwind_tmp(k, iEdge) = dvEdge(iEdge) * dt *
                (std::max(0.0, uhAvg(k, iEdge)) * scalar_old(k, cell1) + std::min(0.0, uhAvg(k, iEdge)) * scalar_old(k, cell2));
            flux_tmp(k, iEdge) = dt * flux_arr(k, iEdge) - flux_upwind_tmp(k, iEdge);
        }
        // TI:4479 (config_apply_lbcs .and. A) .or. B with bdyMaskEdge == 0, nRelaxZone-1 == 4: never true on global meshes
    }
    // !$OMP BARRIER TI:4491
    #pragma omp parallel for
    for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++) {
        for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
            int iEdge = edgesOnCell(i, iCell);
            for (int k = 1; k <= nVertLevels; k++) {
                scalar_new(k, iCell) = scalar_new(k, iCell) - edgesOnCell_sign(i, iCell) * flux_upwind_tmp(k, iEdge) * invAreaCell(iCell);
                scale_arr(k, SCALE_OUT, iCell) = scale_arr(k, SCALE_OUT, iCell)
                    - std::max(0.0, edgesOnCell_sign(i, iCell) * flux_tmp(k, iEdge)) * invAreaCell(iCell);
                scale_arr(k, SCALE_IN, iCell) = scale_arr(k, SCALE_IN, iCell)
                    - std::min(0.0, edgesOnCell_sign(i, iCell) * flux_tmp(k, iEdge)) * invAreaCell(iCell);
            }
        }
    }
    // limiter TI:4523-4553
    A2 rho_lim = local_advance_density ? rho_zz_int : rho_zz_new;
    #pragma omp parallel for
    for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++)
        for (int k = 1; k <= nVertLevels; k++) {
            real scale_factor = (s_max(k, iCell) * rho_lim(k, iCell) - scalar_new(k, iCell)) / (scale_arr(k, SCALE_IN, iCell) + eps);
            scale_arr(k, SCALE_IN, iCell) = std::min(1.0, std::max(0.0, scale_factor));
            scale_factor = (s_min(k, iCell) * rho_lim(k, iCell) - scalar_new(k, iCell)) / (scale_arr(k, SCALE_OUT, iCell) - eps);
            scale_arr(k, SCALE_OUT, iCell) = std::min(1.0, std::max(0.0, scale_factor));
        }
}
// part D: per scalar, TI:4579-4715
static void mono_scalar_phase2(Block& b, real dt, int iScalar) {
    const int nVertLevels = b.d.nVertLevels, nCellsSolve = b.d.nCellsSolve;
    const int cellStart = 1, cellEnd = b.d.nCells, edgeStart = 1, edgeEnd = b.d.nEdges, cellSolveStart = 1, cellSolveEnd = nCellsSolve;
    const bool local_advance_density = b.c.config_split_dynamics_transport != 0;
    const int SCALE_IN = 1, SCALE_OUT = 2;
    A3 scalars_new = b.r3("scalars", 2), scale_arr = b.r3("scale_arr");
    A2 scalar_old = b.r2("scalar_old"), scalar_new = b.r2("scalar_new"), wdtn = b.r2("wdtn"), flux_arr = b.r2("flux_arr"),
       rho_zz_int = b.r2("rho_zz_int"), rho_zz_new = b.r2("rho_zz", 2), uhAvg = b.r2("ruAvg");
    A1 invAreaCell = b.r1("invAreaCell"), dvEdge = b.r1("dvEdge"), rdnw = b.r1("rdzw");
    I2 cellsOnEdge = b.i2("cellsOnEdge"), edgesOnCell = b.i2("edgesOnCell");
    I1 nEdgesOnCell = b.i1("nEdgesOnCell");
    A2 edgesOnCell_sign = b.r2("edgesOnCell_sign");
    #pragma omp parallel for
    for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++) {
        int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
        if (cell1 <= nCellsSolve || cell2 <= nCellsSolve) {
            for (int k = 1; k <= nVertLevels; k++) {
                real flux_upwind = dvEdge(iEdge) * dt *
                    (std::max(0.0, uhAvg(k, iEdge)) * scalar_old(k, cell1) + std::min(0.0, uhAvg(k, iEdge)) * scalar_old(k, cell2));
                flux_arr(k, iEdge) = dt * flux_arr(k, iEdge) - flux_upwind;
            }
            for (int k = 1; k <= nVertLevels; k++) {
                real flux = flux_arr(k, iEdge);
                flux = std::max(0.0, flux) * std::min(scale_arr(k, SCALE_OUT, cell1), scale_arr(k, SCALE_IN, cell2))
                     + std::min(0.0, flux) * std::min(scale_arr(k, SCALE_IN, cell1), scale_arr(k, SCALE_OUT, cell2));
                flux_arr(k, iEdge) = flux;
            }
        }
    }
    // !$OMP BARRIER TI:4631
    #pragma omp parallel for
    for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++) {
        for (int k = 2; k <= nVertLevels; k++) {
            real flux = wdtn(k, iCell);
            flux = std::max(0.0, flux) * std::min(scale_arr(k - 1, SCALE_OUT, iCell), scale_arr(k, SCALE_IN, iCell))
                 + std::min(0.0, flux) * std::min(scale_arr(k, SCALE_OUT, iCell), scale_arr(k - 1, SCALE_IN, iCell));
            wdtn(k, iCell) = flux;
        }
        for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
            int iEdge = edgesOnCell(i, iCell);
            for (int k = 1; k <= nVertLevels; k++)
                scalar_new(k, iCell) = scalar_new(k, iCell) - edgesOnCell_sign(i, iCell) * flux_arr(k, iEdge) * invAreaCell(iCell);
        }
        A2 rho_div = local_advance_density ? rho_zz_int : rho_zz_new;
        for (int k = 1; k <= nVertLevels; k++)
            scalar_new(k, iCell) = (scalar_new(k, iCell) + (-rdnw(k) * (wdtn(k + 1, iCell) - wdtn(k, iCell)))) / rho_div(k, iCell);
    }
    // !$OMP BARRIER TI:4703
    #pragma omp parallel for
    for (int iCell = cellStart; iCell <= cellEnd; iCell++)
        for (int k = 1; k <= nVertLevels; k++) scalars_new(iScalar, k, iCell) = std::max(0.0, scalar_new(k, iCell));
}

// ============================================================ TI:4982-6240
static void atm_compute_dyn_tend(Block& b, int rk_step, real dt) {
    const int nVertLevels = b.d.nVertLevels, vertexDegree = b.d.vertexDegree;
    const int cellStart = 1, cellEnd = b.d.nCells, edgeStart = 1, edgeEnd = b.d.nEdges, vertexStart = 1, vertexEnd = b.d.nVertices;
    const int cellSolveStart = 1, cellSolveEnd = b.d.nCellsSolve, edgeSolveStart = 1, edgeSolveEnd = b.d.nEdgesSolve;
    const OCfg& c = b.c;
    A1 dvEdge = b.r1("dvEdge"), dcEdge = b.r1("dcEdge"), invDcEdge = b.r1("invDcEdge"), invDvEdge = b.r1("invDvEdge"),
       invAreaCell = b.r1("invAreaCell"), invAreaTriangle = b.r1("invAreaTriangle"),
       meshScalingDel2 = b.r1("meshScalingDel2"), meshScalingDel4 = b.r1("meshScalingDel4"), angleEdge = b.r1("angleEdge");
    A2 weightsOnEdge = b.r2("weightsOnEdge"), zgrid = b.r2("zgrid"), rho_edge = b.r2("rho_edge"), rho_zz = b.r2("rho_zz", 2),
       ru = b.r2("ru"), u = b.r2("u", 2), v = b.r2("v"), tend_u = b.r2("tend_u"), divergence = b.r2("divergence"),
       vorticity = b.r2("vorticity"), ke = b.r2("ke"), pv_edge = b.r2("pv_edge"), theta_m = b.r2("theta_m", 2), rw = b.r2("rw"),
       tend_rho = b.r2("tend_rho"), rt_diabatic_tend = b.r2("rt_diabatic_tend"), tend_theta = b.r2("tend_theta"),
       tend_w = b.r2("tend_w"), w = b.r2("w", 2), cqw = b.r2("cqw"), rb = b.r2("rho_base"), rr_save = b.r2("rho_p_save"),
       pp = b.r2("pressure_p"), zz = b.r2("zz"), zxu = b.r2("zxu"), cqu = b.r2("cqu"), h_divergence = b.r2("h_divergence"),
       kdiff = b.r2("kdiff"), edgesOnCell_sign = b.r2("edgesOnCell_sign"), edgesOnVertex_sign = b.r2("edgesOnVertex_sign"),
       rw_save = b.r2("rw_save"), ru_save = b.r2("ru_save"), theta_m_save = b.r2("theta_m", 1),
       tend_u_euler = b.r2("tend_u_euler"), tend_w_euler = b.r2("tend_w_euler"), tend_theta_euler = b.r2("tend_theta_euler"),
       adv_coefs = b.r2("adv_coefs"), adv_coefs_3rd = b.r2("adv_coefs_3rd"), t_init = b.r2("t_init"),
       defc_a = b.r2("defc_a"), defc_b = b.r2("defc_b"), rthdynten = b.r2("rthdynten");
    A2 tend_ru_physics = b.r2("tend_ru_physics"), tend_rtheta_physics = b.r2("tend_rtheta_physics"), tend_rho_physics = b.r2("tend_rho_physics");
    A2 qtot = b.r2("qtot"), delsq_theta = b.r2("delsq_theta"), delsq_w = b.r2("delsq_w"), delsq_divergence = b.r2("delsq_divergence"),
       delsq_u = b.r2("delsq_u"), delsq_vorticity = b.r2("delsq_vorticity"), dpdz = b.r2("dpdz");
    A1 rdzu = b.r1("rdzu"), rdzw = b.r1("rdzw"), fzm = b.r1("fzm"), fzp = b.r1("fzp"), u_init = b.r1("u_init"), v_init = b.r1("v_init");
    I2 cellsOnEdge = b.i2("cellsOnEdge"), verticesOnEdge = b.i2("verticesOnEdge"), edgesOnCell = b.i2("edgesOnCell"),
       edgesOnEdge = b.i2("edgesOnEdge"), edgesOnVertex = b.i2("edgesOnVertex"), advCellsForEdge = b.i2("advCellsForEdge");
    I1 nEdgesOnCell = b.i1("nEdgesOnCell"), nEdgesOnEdge = b.i1("nEdgesOnEdge"), nAdvCellsForEdge = b.i1("nAdvCellsForEdge");
    const real coef_3rd_order = c.config_coef_3rd_order, c_s = c.config_smagorinsky_coef, config_len_disp = c.config_len_disp;

    // scratch garbage slots are zeroed by the caller every stage (TI:1160-1170)
    for (int k = 1; k <= nVertLevels; k++) {
        delsq_theta(k, b.d.nCells + 1) = 0.0; delsq_w(k, b.d.nCells + 1) = 0.0; delsq_divergence(k, b.d.nCells + 1) = 0.0;
        delsq_u(k, b.d.nEdges + 1) = 0.0; delsq_vorticity(k, b.d.nVertices + 1) = 0.0; dpdz(k, b.d.nCells + 1) = 0.0;
    }

    const real prandtl_inv = 1.0 / prandtl;
    const real invDt = 1.0 / dt;
    const real v_mom_eddy_visc2 = c.config_v_mom_eddy_visc2;
    const real v_theta_eddy_visc2 = c.config_v_theta_eddy_visc2;
    real h_mom_eddy_visc4 = 0.0, h_theta_eddy_visc4 = 0.0;

    if (rk_step == 1) {
        #pragma omp parallel for
        for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++)
            for (int k = 1; k <= nVertLevels; k++) tend_u_euler(k, iEdge) = 0.0;
        if (c.config_horiz_mixing == 0) {    // "2d_smagorinsky" TI:5226-5259
            #pragma omp parallel for
            for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
                std::vector<real> dd(nVertLevels + 1, 0.0), dod(nVertLevels + 1, 0.0);
                A1 d_diag{dd.data()}, d_off_diag{dod.data()};
                for (int iEdge = 1; iEdge <= nEdgesOnCell(iCell); iEdge++)
                    for (int k = 1; k <= nVertLevels; k++) {
                        d_diag(k) = d_diag(k) + defc_a(iEdge, iCell) * u(k, edgesOnCell(iEdge, iCell))
                                              - defc_b(iEdge, iCell) * v(k, edgesOnCell(iEdge, iCell));
                        d_off_diag(k) = d_off_diag(k) + defc_b(iEdge, iCell) * u(k, edgesOnCell(iEdge, iCell))
                                                      + defc_a(iEdge, iCell) * v(k, edgesOnCell(iEdge, iCell));
                    }
                for (int k = 1; k <= nVertLevels; k++)
                    kdiff(k, iCell) = std::min((c_s * config_len_disp) * (c_s * config_len_disp) * std::sqrt(d_diag(k) * d_diag(k) + d_off_diag(k) * d_off_diag(k)),
                                               (0.01 * (config_len_disp * config_len_disp)) * invDt);
            }
            h_mom_eddy_visc4 = c.config_visc4_2dsmag * (config_len_disp * config_len_disp * config_len_disp);
            h_theta_eddy_visc4 = h_mom_eddy_visc4;
        } else {                              // "2d_fixed" TI:5261-5276
            for (int iCell = cellStart; iCell <= cellEnd; iCell++)
                for (int k = 1; k <= nVertLevels; k++) kdiff(k, iCell) = c.config_h_theta_eddy_visc2;
            h_mom_eddy_visc4 = c.config_h_mom_eddy_visc4;
            h_theta_eddy_visc4 = c.config_h_theta_eddy_visc4;
        }
        if (c.config_mpas_cam_coef > 0.0) {   // TI:5278-5296
            for (int iCell = cellStart; iCell <= cellEnd; iCell++)
                for (int k = nVertLevels - c.config_number_cam_damping_levels + 1; k <= nVertLevels; k++) {
                    real visc2cam = 4.0 * 2.0833 * config_len_disp * c.config_mpas_cam_coef;
                    visc2cam = visc2cam * (1.0 - (real)(nVertLevels - k) / (real)(c.config_number_cam_damping_levels));
                    kdiff(k, iCell) = std::max(kdiff(k, iCell), visc2cam);
                }
        }
    }

    #pragma omp parallel for
    for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
        for (int k = 1; k <= nVertLevels; k++) h_divergence(k, iCell) = 0.0;
        for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
            int iEdge = edgesOnCell(i, iCell);
            real edge_sign = edgesOnCell_sign(i, iCell) * dvEdge(iEdge);
            for (int k = 1; k <= nVertLevels; k++) h_divergence(k, iCell) = h_divergence(k, iCell) + edge_sign * ru(k, iEdge);
        }
        real r = invAreaCell(iCell);
        for (int k = 1; k <= nVertLevels; k++) h_divergence(k, iCell) = h_divergence(k, iCell) * r;
    }
    if (rk_step == 1) {
        #pragma omp parallel for
        for (int iCell = cellStart; iCell <= cellEnd; iCell++)
            for (int k = 1; k <= nVertLevels; k++) {
                tend_rho(k, iCell) = -h_divergence(k, iCell) - rdzw(k) * (rw(k + 1, iCell) - rw(k, iCell)) + tend_rho_physics(k, iCell);
                dpdz(k, iCell) = -gravity * (rb(k, iCell) * (qtot(k, iCell)) + rr_save(k, iCell) * (1. + qtot(k, iCell)));
            }
    }
    // !$OMP BARRIER TI:5364
    #pragma omp parallel for
    for (int iEdge = edgeSolveStart; iEdge <= edgeSolveEnd; iEdge++) {
        std::vector<real> wduz_v(nVertLevels + 2), q_v(nVertLevels + 1);
        A1 wduz{wduz_v.data()}, q{q_v.data()};
        int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
        if (rk_step == 1)
            for (int k = 1; k <= nVertLevels; k++)
                tend_u_euler(k, iEdge) = -cqu(k, iEdge) * ((pp(k, cell2) - pp(k, cell1)) * invDcEdge(iEdge) / (.5 * (zz(k, cell2) + zz(k, cell1)))
                                                           - 0.5 * zxu(k, iEdge) * (dpdz(k, cell1) + dpdz(k, cell2)));
        wduz(1) = 0.;
        int k = 2;
        wduz(k) = 0.5 * (rw(k, cell1) + rw(k, cell2)) * (fzm(k) * u(k, iEdge) + fzp(k) * u(k - 1, iEdge));
        for (k = 3; k <= nVertLevels - 1; k++)
            wduz(k) = flux3(u(k - 2, iEdge), u(k - 1, iEdge), u(k, iEdge), u(k + 1, iEdge), 0.5 * (rw(k, cell1) + rw(k, cell2)), 1.0);
        k = nVertLevels;
        wduz(k) = 0.5 * (rw(k, cell1) + rw(k, cell2)) * (fzm(k) * u(k, iEdge) + fzp(k) * u(k - 1, iEdge));
        wduz(nVertLevels + 1) = 0.;
        for (k = 1; k <= nVertLevels; k++) tend_u(k, iEdge) = -rdzw(k) * (wduz(k + 1) - wduz(k));
        for (k = 1; k <= nVertLevels; k++) q(k) = 0.0;
        for (int j = 1; j <= nEdgesOnEdge(iEdge); j++) {
            int eoe = edgesOnEdge(j, iEdge);
            for (k = 1; k <= nVertLevels; k++) {
                real workpv = 0.5 * (pv_edge(k, iEdge) + pv_edge(k, eoe));
                q(k) = q(k) + weightsOnEdge(j, iEdge) * u(k, eoe) * workpv;
            }
        }
        for (k = 1; k <= nVertLevels; k++)
            tend_u(k, iEdge) = tend_u(k, iEdge) + rho_edge(k, iEdge) * (q(k) - (ke(k, cell2) - ke(k, cell1))
                                                                         * invDcEdge(iEdge))
                               - u(k, iEdge) * 0.5 * (h_divergence(k, cell1) + h_divergence(k, cell2));
    }

    if (rk_step == 1) {
        // !$OMP BARRIER TI:5460
        #pragma omp parallel for
        for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++) {
            for (int k = 1; k <= nVertLevels; k++) delsq_u(k, iEdge) = 0.0;
            int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
            int vertex1 = verticesOnEdge(1, iEdge), vertex2 = verticesOnEdge(2, iEdge);
            real r_dc = invDcEdge(iEdge);
            real r_dv = std::min(invDvEdge(iEdge), 4 * invDcEdge(iEdge));
            for (int k = 1; k <= nVertLevels; k++) {
                real u_diffusion = (divergence(k, cell2) - divergence(k, cell1)) * r_dc
                                   - (vorticity(k, vertex2) - vorticity(k, vertex1)) * r_dv;
                delsq_u(k, iEdge) = delsq_u(k, iEdge) + u_diffusion;
                real kdiffu = 0.5 * (kdiff(k, cell1) + kdiff(k, cell2));
                tend_u_euler(k, iEdge) = tend_u_euler(k, iEdge)
                                         + rho_edge(k, iEdge) * kdiffu * u_diffusion * meshScalingDel2(iEdge);
            }
        }
        if (h_mom_eddy_visc4 > 0.0) {
            // !$OMP BARRIER TI:5508
            #pragma omp parallel for
            for (int iVertex = vertexStart; iVertex <= vertexEnd; iVertex++) {
                for (int k = 1; k <= nVertLevels; k++) delsq_vorticity(k, iVertex) = 0.0;
                for (int i = 1; i <= vertexDegree; i++) {
                    int iEdge = edgesOnVertex(i, iVertex);
                    real edge_sign = invAreaTriangle(iVertex) * dcEdge(iEdge) * edgesOnVertex_sign(i, iVertex);
                    for (int k = 1; k <= nVertLevels; k++)
                        delsq_vorticity(k, iVertex) = delsq_vorticity(k, iVertex) + edge_sign * delsq_u(k, iEdge);
                }
            }
            #pragma omp parallel for
            for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
                for (int k = 1; k <= nVertLevels; k++) delsq_divergence(k, iCell) = 0.0;
                real r = invAreaCell(iCell);
                for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
                    int iEdge = edgesOnCell(i, iCell);
                    real edge_sign = r * dvEdge(iEdge) * edgesOnCell_sign(i, iCell);
                    for (int k = 1; k <= nVertLevels; k++)
                        delsq_divergence(k, iCell) = delsq_divergence(k, iCell) + edge_sign * delsq_u(k, iEdge);
                }
            }
            // !$OMP BARRIER TI:5554
            #pragma omp parallel for
            for (int iEdge = edgeSolveStart; iEdge <= edgeSolveEnd; iEdge++) {
                int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
                int vertex1 = verticesOnEdge(1, iEdge), vertex2 = verticesOnEdge(2, iEdge);
                real u_mix_scale = meshScalingDel4(iEdge) * h_mom_eddy_visc4;
                real r_dc = u_mix_scale * c.config_del4u_div_factor * invDcEdge(iEdge);
                real r_dv = u_mix_scale * std::min(invDvEdge(iEdge), 4 * invDcEdge(iEdge));
                for (int k = 1; k <= nVertLevels; k++) {
                    real u_diffusion = rho_edge(k, iEdge) * ((delsq_divergence(k, cell2) - delsq_divergence(k, cell1)) * r_dc
                                                             - (delsq_vorticity(k, vertex2) - delsq_vorticity(k, vertex1)) * r_dv);
                    tend_u_euler(k, iEdge) = tend_u_euler(k, iEdge) - u_diffusion;
                }
            }
        }
        if (v_mom_eddy_visc2 > 0.0) {        // TI:5592-5658
            #pragma omp parallel for
            for (int iEdge = edgeSolveStart; iEdge <= edgeSolveEnd; iEdge++) {
                int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
                std::vector<real> um(nVertLevels + 1);
                A1 u_mix{um.data()};
                for (int k = 1; k <= nVertLevels; k++)
                    u_mix(k) = c.config_mix_full ? u(k, iEdge)
                                                 : u(k, iEdge) - u_init(k) * std::cos(angleEdge(iEdge)) - v_init(k) * std::sin(angleEdge(iEdge));
                for (int k = 2; k <= nVertLevels - 1; k++) {
                    real z1 = 0.5 * (zgrid(k - 1, cell1) + zgrid(k - 1, cell2));
                    real z2 = 0.5 * (zgrid(k, cell1) + zgrid(k, cell2));
                    real z3 = 0.5 * (zgrid(k + 1, cell1) + zgrid(k + 1, cell2));
                    real z4 = 0.5 * (zgrid(k + 2, cell1) + zgrid(k + 2, cell2));
                    real zm = 0.5 * (z1 + z2), z0 = 0.5 * (z2 + z3), zp = 0.5 * (z3 + z4);
                    tend_u_euler(k, iEdge) = tend_u_euler(k, iEdge) + rho_edge(k, iEdge) * v_mom_eddy_visc2 * (
                                                 (u_mix(k + 1) - u_mix(k)) / (zp - z0)
                                                 - (u_mix(k) - u_mix(k - 1)) / (z0 - zm)) / (0.5 * (zp - zm));
                }
            }
        }
    }
    // !$OMP BARRIER TI:5662
    if (c.config_rayleigh_damp_u) {          // TI:5667-5690
        const int nl = c.config_number_rayleigh_damp_u_levels;
        real rayleigh_coef_inverse = 1.0 / ((real)nl * (c.config_rayleigh_damp_u_timescale_days * seconds_per_day));
        std::vector<real> rdc(nVertLevels + 1, 0.0);
        A1 rayleigh_damp_coef{rdc.data()};
        for (int k = nVertLevels - nl + 1; k <= nVertLevels; k++) rayleigh_damp_coef(k) = (real)(k - (nVertLevels - nl)) * rayleigh_coef_inverse;
        #pragma omp parallel for
        for (int iEdge = edgeSolveStart; iEdge <= edgeSolveEnd; iEdge++)
            for (int k = nVertLevels - nl + 1; k <= nVertLevels; k++)
                tend_u(k, iEdge) = tend_u(k, iEdge) - rho_edge(k, iEdge) * u(k, iEdge) * rayleigh_damp_coef(k);
    }
    #pragma omp parallel for
    for (int iEdge = edgeSolveStart; iEdge <= edgeSolveEnd; iEdge++)
        for (int k = 1; k <= nVertLevels; k++)
            tend_u(k, iEdge) = tend_u(k, iEdge) + tend_u_euler(k, iEdge) + tend_ru_physics(k, iEdge);

    // ----------- rhs for w  TI:5705-5945
    #pragma omp parallel for
    for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++) {
        std::vector<real> re(nVertLevels + 2), fa(nVertLevels + 2);
        A1 ru_edge_w{re.data()}, flux_arr{fa.data()};
        for (int k = 1; k <= nVertLevels + 1; k++) tend_w(k, iCell) = 0.0;
        for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
            int iEdge = edgesOnCell(i, iCell);
            for (int k = 2; k <= nVertLevels; k++) ru_edge_w(k) = fzm(k) * ru(k, iEdge) + fzp(k) * ru(k - 1, iEdge);
            for (int k = 1; k <= nVertLevels; k++) flux_arr(k) = 0.0;
            for (int j = 1; j <= nAdvCellsForEdge(iEdge); j++) {
                int iAdvCell = advCellsForEdge(j, iEdge);
                for (int k = 2; k <= nVertLevels; k++) {
                    real scalar_weight = adv_coefs(j, iEdge) + sign1(ru_edge_w(k)) * adv_coefs_3rd(j, iEdge);
                    flux_arr(k) = flux_arr(k) + scalar_weight * w(k, iAdvCell);
                }
            }
            for (int k = 2; k <= nVertLevels; k++)
                tend_w(k, iCell) = tend_w(k, iCell) - edgesOnCell_sign(i, iCell) * ru_edge_w(k) * flux_arr(k);
        }
    }
    if (rk_step == 1) {
        #pragma omp parallel for
        for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
            for (int k = 1; k <= nVertLevels; k++) delsq_w(k, iCell) = 0.0;
            for (int k = 1; k <= nVertLevels + 1; k++) tend_w_euler(k, iCell) = 0.0;
            real r_areaCell = invAreaCell(iCell);
            for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
                int iEdge = edgesOnCell(i, iCell);
                real edge_sign = 0.5 * r_areaCell * edgesOnCell_sign(i, iCell) * dvEdge(iEdge) * invDcEdge(iEdge);
                int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
                for (int k = 2; k <= nVertLevels; k++) {
                    real w_turb_flux = edge_sign * (rho_edge(k, iEdge) + rho_edge(k - 1, iEdge)) * (w(k, cell2) - w(k, cell1));
                    delsq_w(k, iCell) = delsq_w(k, iCell) + w_turb_flux;
                    w_turb_flux = w_turb_flux * meshScalingDel2(iEdge) * 0.25 *
                                  (kdiff(k, cell1) + kdiff(k, cell2) + kdiff(k - 1, cell1) + kdiff(k - 1, cell2));
                    tend_w_euler(k, iCell) = tend_w_euler(k, iCell) + w_turb_flux;
                }
            }
        }
        // !$OMP BARRIER TI:5832
        if (h_mom_eddy_visc4 > 0.0) {
            #pragma omp parallel for
            for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++) {
                real r_areaCell = h_mom_eddy_visc4 * invAreaCell(iCell);
                for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
                    int iEdge = edgesOnCell(i, iCell);
                    int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
                    real edge_sign = meshScalingDel4(iEdge) * r_areaCell * dvEdge(iEdge) * edgesOnCell_sign(i, iCell) * invDcEdge(iEdge);
                    for (int k = 2; k <= nVertLevels; k++)
                        tend_w_euler(k, iCell) = tend_w_euler(k, iCell) - edge_sign * (delsq_w(k, cell2) - delsq_w(k, cell1));
                }
            }
        }
    }
    #pragma omp parallel for
    for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++) {
        std::vector<real> wd(nVertLevels + 2);
        A1 wdwz{wd.data()};
        wdwz(1) = 0.0;
        int k = 2;
        wdwz(k) = 0.25 * (rw(k, iCell) + rw(k - 1, iCell)) * (w(k, iCell) + w(k - 1, iCell));
        for (k = 3; k <= nVertLevels - 1; k++)
            wdwz(k) = flux3(w(k - 2, iCell), w(k - 1, iCell), w(k, iCell), w(k + 1, iCell), 0.5 * (rw(k, iCell) + rw(k - 1, iCell)), 1.0);
        k = nVertLevels;
        wdwz(k) = 0.25 * (rw(k, iCell) + rw(k - 1, iCell)) * (w(k, iCell) + w(k - 1, iCell));
        wdwz(nVertLevels + 1) = 0.0;
        for (k = 2; k <= nVertLevels; k++)
            tend_w(k, iCell) = tend_w(k, iCell) * invAreaCell(iCell) - rdzu(k) * (wdwz(k + 1) - wdwz(k));
        if (rk_step == 1)
            for (k = 2; k <= nVertLevels; k++)
                tend_w_euler(k, iCell) = tend_w_euler(k, iCell) - cqw(k, iCell) * (
                                             rdzu(k) * (pp(k, iCell) - pp(k - 1, iCell))
                                             - (fzm(k) * dpdz(k, iCell) + fzp(k) * dpdz(k - 1, iCell)));
    }
    if (rk_step == 1 && v_mom_eddy_visc2 > 0.0) {
        #pragma omp parallel for
        for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++)
            for (int k = 2; k <= nVertLevels; k++)
                tend_w_euler(k, iCell) = tend_w_euler(k, iCell) + v_mom_eddy_visc2 * 0.5 * (rho_zz(k, iCell) + rho_zz(k - 1, iCell)) * (
                                             (w(k + 1, iCell) - w(k, iCell)) * rdzw(k)
                                             - (w(k, iCell) - w(k - 1, iCell)) * rdzw(k - 1)) * rdzu(k);
    }
    #pragma omp parallel for
    for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++)
        for (int k = 2; k <= nVertLevels; k++) tend_w(k, iCell) = tend_w(k, iCell) + tend_w_euler(k, iCell);

    // ----------- rhs for theta  TI:5948-6197
    #pragma omp parallel for
    for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++) {
        std::vector<real> fa(nVertLevels + 2);
        A1 flux_arr{fa.data()};
        for (int k = 1; k <= nVertLevels; k++) tend_theta(k, iCell) = 0.0;
        for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
            int iEdge = edgesOnCell(i, iCell);
            for (int k = 1; k <= nVertLevels; k++) flux_arr(k) = 0.0;
            for (int j = 1; j <= nAdvCellsForEdge(iEdge); j++) {
                int iAdvCell = advCellsForEdge(j, iEdge);
                for (int k = 1; k <= nVertLevels; k++) {
                    real scalar_weight = adv_coefs(j, iEdge) + sign1(ru(k, iEdge)) * adv_coefs_3rd(j, iEdge);
                    flux_arr(k) = flux_arr(k) + scalar_weight * theta_m(k, iAdvCell);
                }
            }
            for (int k = 1; k <= nVertLevels; k++)
                tend_theta(k, iCell) = tend_theta(k, iCell) - edgesOnCell_sign(i, iCell) * ru(k, iEdge) * flux_arr(k);
        }
    }
    if (rk_step > 1) {
        #pragma omp parallel for
        for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++)
            for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
                int iEdge = edgesOnCell(i, iCell);
                int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
                for (int k = 1; k <= nVertLevels; k++) {
                    real flux = edgesOnCell_sign(i, iCell) * dvEdge(iEdge) * (ru_save(k, iEdge) - ru(k, iEdge)) * 0.5 * (theta_m_save(k, cell2) + theta_m_save(k, cell1));
                    tend_theta(k, iCell) = tend_theta(k, iCell) - flux;
                }
            }
    }
    if (rk_step == 1) {
        #pragma omp parallel for
        for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
            for (int k = 1; k <= nVertLevels; k++) { delsq_theta(k, iCell) = 0.0; tend_theta_euler(k, iCell) = 0.0; }
            real r_areaCell = invAreaCell(iCell);
            for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
                int iEdge = edgesOnCell(i, iCell);
                real edge_sign = r_areaCell * edgesOnCell_sign(i, iCell) * dvEdge(iEdge) * invDcEdge(iEdge);
                real pr_scale = prandtl_inv * meshScalingDel2(iEdge);
                int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
                for (int k = 1; k <= nVertLevels; k++) {
                    real theta_turb_flux = edge_sign * (theta_m(k, cell2) - theta_m(k, cell1)) * rho_edge(k, iEdge);
                    delsq_theta(k, iCell) = delsq_theta(k, iCell) + theta_turb_flux;
                    theta_turb_flux = theta_turb_flux * 0.5 * (kdiff(k, cell1) + kdiff(k, cell2)) * pr_scale;
                    tend_theta_euler(k, iCell) = tend_theta_euler(k, iCell) + theta_turb_flux;
                }
            }
        }
        // !$OMP BARRIER TI:6060
        if (h_theta_eddy_visc4 > 0.0) {
            #pragma omp parallel for
            for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++) {
                real r_areaCell = h_theta_eddy_visc4 * prandtl_inv * invAreaCell(iCell);
                for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
                    int iEdge = edgesOnCell(i, iCell);
                    real edge_sign = meshScalingDel4(iEdge) * r_areaCell * dvEdge(iEdge) * edgesOnCell_sign(i, iCell) * invDcEdge(iEdge);
                    int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
                    for (int k = 1; k <= nVertLevels; k++)
                        tend_theta_euler(k, iCell) = tend_theta_euler(k, iCell) - edge_sign * (delsq_theta(k, cell2) - delsq_theta(k, cell1));
                }
            }
        }
    }
    #pragma omp parallel for
    for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++) {
        std::vector<real> wd(nVertLevels + 2);
        A1 wdtz{wd.data()};
        wdtz(1) = 0.0;
        int k = 2;
        wdtz(k) = rw(k, iCell) * (fzm(k) * theta_m(k, iCell) + fzp(k) * theta_m(k - 1, iCell));
        wdtz(k) = wdtz(k) + (rw_save(k, iCell) - rw(k, iCell)) * (fzm(k) * theta_m_save(k, iCell) + fzp(k) * theta_m_save(k - 1, iCell));
        for (k = 3; k <= nVertLevels - 1; k++) {
            wdtz(k) = flux3(theta_m(k - 2, iCell), theta_m(k - 1, iCell), theta_m(k, iCell), theta_m(k + 1, iCell), rw(k, iCell), coef_3rd_order);
            wdtz(k) = wdtz(k) + (rw_save(k, iCell) - rw(k, iCell)) * (fzm(k) * theta_m_save(k, iCell) + fzp(k) * theta_m_save(k - 1, iCell));
        }
        k = nVertLevels;
        wdtz(k) = rw_save(k, iCell) * (fzm(k) * theta_m(k, iCell) + fzp(k) * theta_m(k - 1, iCell));
        wdtz(nVertLevels + 1) = 0.0;
        for (k = 1; k <= nVertLevels; k++) {
            tend_theta(k, iCell) = tend_theta(k, iCell) * invAreaCell(iCell) - rdzw(k) * (wdtz(k + 1) - wdtz(k));
            rthdynten(k, iCell) = (tend_theta(k, iCell) - tend_rho(k, iCell) * theta_m(k, iCell)) / rho_zz(k, iCell);
            tend_theta(k, iCell) = tend_theta(k, iCell) + rho_zz(k, iCell) * rt_diabatic_tend(k, iCell);
        }
    }
    if (rk_step == 1 && v_theta_eddy_visc2 > 0.0) {   // TI:6134-6184
        #pragma omp parallel for
        for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++)
            for (int k = 2; k <= nVertLevels - 1; k++) {
                real z1 = zgrid(k - 1, iCell), z2 = zgrid(k, iCell), z3 = zgrid(k + 1, iCell), z4 = zgrid(k + 2, iCell);
                real zm = 0.5 * (z1 + z2), z0 = 0.5 * (z2 + z3), zp = 0.5 * (z3 + z4);
                if (c.config_mix_full)
                    tend_theta_euler(k, iCell) = tend_theta_euler(k, iCell) + v_theta_eddy_visc2 * prandtl_inv * rho_zz(k, iCell) * (
                                                     (theta_m(k + 1, iCell) - theta_m(k, iCell)) / (zp - z0)
                                                     - (theta_m(k, iCell) - theta_m(k - 1, iCell)) / (z0 - zm)) / (0.5 * (zp - zm));
                else
                    tend_theta_euler(k, iCell) = tend_theta_euler(k, iCell) + v_theta_eddy_visc2 * prandtl_inv * rho_zz(k, iCell) * (
                                                     ((theta_m(k + 1, iCell) - t_init(k + 1, iCell)) - (theta_m(k, iCell) - t_init(k, iCell))) / (zp - z0)
                                                     - ((theta_m(k, iCell) - t_init(k, iCell)) - (theta_m(k - 1, iCell) - t_init(k - 1, iCell))) / (z0 - zm)) / (0.5 * (zp - zm));
            }
    }
    #pragma omp parallel for
    for (int iCell = cellSolveStart; iCell <= cellSolveEnd; iCell++)
        for (int k = 1; k <= nVertLevels; k++)
            tend_theta(k, iCell) = tend_theta(k, iCell) + tend_theta_euler(k, iCell) + tend_rtheta_physics(k, iCell);
}

// ============================================================ TI:6337-6773
static void atm_compute_solve_diagnostics(Block& b, real dt, int time_lev, int rk_step /* 0 = absent */) {
    const int nVertLevels = b.d.nVertLevels, vertexDegree = b.d.vertexDegree;
    const int cellStart = 1, cellEnd = b.d.nCells, edgeStart = 1, edgeEnd = b.d.nEdges, vertexStart = 1, vertexEnd = b.d.nVertices;
    const real config_apvm_upwinding = b.c.config_apvm_upwinding;
    A1 fVertex = b.r1("fVertex"), invAreaTriangle = b.r1("invAreaTriangle"), invAreaCell = b.r1("invAreaCell"),
       dvEdge = b.r1("dvEdge"), dcEdge = b.r1("dcEdge"), invDvEdge = b.r1("invDvEdge"), invDcEdge = b.r1("invDcEdge");
    A2 weightsOnEdge = b.r2("weightsOnEdge"), kiteAreasOnVertex = b.r2("kiteAreasOnVertex"), h_edge = b.r2("rho_edge"),
       h = b.r2("rho_zz", time_lev), u = b.r2("u", time_lev), v = b.r2("v"), vorticity = b.r2("vorticity"), ke = b.r2("ke"),
       pv_edge = b.r2("pv_edge"), pv_vertex = b.r2("pv_vertex"), pv_cell = b.r2("pv_cell"), gradPVn = b.r2("gradPVn"),
       gradPVt = b.r2("gradPVt"), divergence = b.r2("divergence"), ke_vertex = b.r2("ke_vertex"), ke_edge = b.r2("ke_edge");
    A2 edgesOnVertex_sign = b.r2("edgesOnVertex_sign"), edgesOnCell_sign = b.r2("edgesOnCell_sign");
    I2 cellsOnEdge = b.i2("cellsOnEdge"), verticesOnEdge = b.i2("verticesOnEdge"), edgesOnCell = b.i2("edgesOnCell"),
       edgesOnEdge = b.i2("edgesOnEdge"), edgesOnVertex = b.i2("edgesOnVertex"), kiteForCell = b.i2("kiteForCell"),
       verticesOnCell = b.i2("verticesOnCell");
    I1 nEdgesOnCell = b.i1("nEdgesOnCell"), nEdgesOnEdge = b.i1("nEdgesOnEdge");
    for (int k = 1; k <= nVertLevels; k++) { ke_vertex(k, b.d.nVertices + 1) = 0.0; ke_edge(k, b.d.nEdges + 1) = 0.0; }  // TI:1441-1447

    #pragma omp parallel for
    for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++) {
        int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
        for (int k = 1; k <= nVertLevels; k++) h_edge(k, iEdge) = 0.5 * (h(k, cell1) + h(k, cell2));
        real efac = dcEdge(iEdge) * dvEdge(iEdge);
        for (int k = 1; k <= nVertLevels; k++) ke_edge(k, iEdge) = efac * (u(k, iEdge) * u(k, iEdge));
    }
    #pragma omp parallel for
    for (int iVertex = vertexStart; iVertex <= vertexEnd; iVertex++) {
        for (int k = 1; k <= nVertLevels; k++) vorticity(k, iVertex) = 0.0;
        for (int i = 1; i <= vertexDegree; i++) {
            int iEdge = edgesOnVertex(i, iVertex);
            real s = edgesOnVertex_sign(i, iVertex) * dcEdge(iEdge);
            for (int k = 1; k <= nVertLevels; k++) vorticity(k, iVertex) = vorticity(k, iVertex) + s * u(k, iEdge);
        }
        for (int k = 1; k <= nVertLevels; k++) vorticity(k, iVertex) = vorticity(k, iVertex) * invAreaTriangle(iVertex);
    }
    #pragma omp parallel for
    for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
        for (int k = 1; k <= nVertLevels; k++) divergence(k, iCell) = 0.0;
        for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
            int iEdge = edgesOnCell(i, iCell);
            real s = edgesOnCell_sign(i, iCell) * dvEdge(iEdge);
            for (int k = 1; k <= nVertLevels; k++) divergence(k, iCell) = divergence(k, iCell) + s * u(k, iEdge);
        }
        real r = invAreaCell(iCell);
        for (int k = 1; k <= nVertLevels; k++) divergence(k, iCell) = divergence(k, iCell) * r;
    }
    // !$OMP BARRIER TI:6503
    #pragma omp parallel for
    for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
        for (int k = 1; k <= nVertLevels; k++) ke(k, iCell) = 0.0;
        for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
            int iEdge = edgesOnCell(i, iCell);
            for (int k = 1; k <= nVertLevels; k++) ke(k, iCell) = ke(k, iCell) + 0.25 * ke_edge(k, iEdge);
        }
        for (int k = 1; k <= nVertLevels; k++) ke(k, iCell) = ke(k, iCell) * invAreaCell(iCell);
    }
    // hollingsworth = .true.  TI:6538-6596
    #pragma omp parallel for
    for (int iVertex = vertexStart; iVertex <= vertexEnd; iVertex++) {
        real r = 0.25 * invAreaTriangle(iVertex);
        for (int k = 1; k <= nVertLevels; k++)
            ke_vertex(k, iVertex) = (ke_edge(k, edgesOnVertex(1, iVertex)) + ke_edge(k, edgesOnVertex(2, iVertex)) + ke_edge(k, edgesOnVertex(3, iVertex))) * r;
    }
    // !$OMP BARRIER TI:6564
    const real ke_fact = 1.0 - .375;
    #pragma omp parallel for
    for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
        for (int k = 1; k <= nVertLevels; k++) ke(k, iCell) = ke_fact * ke(k, iCell);
        real r = invAreaCell(iCell);
        for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
            int iVertex = verticesOnCell(i, iCell);
            int j = kiteForCell(i, iCell);
            for (int k = 1; k <= nVertLevels; k++)
                ke(k, iCell) = ke(k, iCell) + (1. - ke_fact) * kiteAreasOnVertex(j, iVertex) * ke_vertex(k, iVertex) * r;
        }
    }
    bool reconstruct_v = true;
    if (rk_step != 0 && rk_step != 3) reconstruct_v = false;
    if (reconstruct_v) {
        #pragma omp parallel for
        for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++) {
            for (int k = 1; k <= nVertLevels; k++) v(k, iEdge) = 0.0;
            for (int i = 1; i <= nEdgesOnEdge(iEdge); i++) {
                int eoe = edgesOnEdge(i, iEdge);
                for (int k = 1; k <= nVertLevels; k++) v(k, iEdge) = v(k, iEdge) + weightsOnEdge(i, iEdge) * u(k, eoe);
            }
        }
    }
    #pragma omp parallel for
    for (int iVertex = vertexStart; iVertex <= vertexEnd; iVertex++)
        for (int k = 1; k <= nVertLevels; k++) pv_vertex(k, iVertex) = (fVertex(iVertex) + vorticity(k, iVertex));
    // !$OMP BARRIER TI:6662
    #pragma omp parallel for
    for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++)
        for (int k = 1; k <= nVertLevels; k++)
            pv_edge(k, iEdge) = 0.5 * (pv_vertex(k, verticesOnEdge(1, iEdge)) + pv_vertex(k, verticesOnEdge(2, iEdge)));
    if (config_apvm_upwinding > 0.0) {
        #pragma omp parallel for
        for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
            for (int k = 1; k <= nVertLevels; k++) pv_cell(k, iCell) = 0.0;
            real r = invAreaCell(iCell);
            for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
                int iVertex = verticesOnCell(i, iCell);
                int j = kiteForCell(i, iCell);
                for (int k = 1; k <= nVertLevels; k++)
                    pv_cell(k, iCell) = pv_cell(k, iCell) + kiteAreasOnVertex(j, iVertex) * pv_vertex(k, iVertex) * r;
            }
        }
        // !$OMP BARRIER TI:6713
        const real r = config_apvm_upwinding * dt;
        #pragma omp parallel for
        for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++) {
            real r1 = 1.0 * invDvEdge(iEdge);
            real r2 = 1.0 * invDcEdge(iEdge);
            for (int k = 1; k <= nVertLevels; k++) {
                gradPVt(k, iEdge) = (pv_vertex(k, verticesOnEdge(2, iEdge)) - pv_vertex(k, verticesOnEdge(1, iEdge))) * r1;
                gradPVn(k, iEdge) = (pv_cell(k, cellsOnEdge(2, iEdge)) - pv_cell(k, cellsOnEdge(1, iEdge))) * r2;
                pv_edge(k, iEdge) = pv_edge(k, iEdge) - r * (v(k, iEdge) * gradPVt(k, iEdge) + u(k, iEdge) * gradPVn(k, iEdge));
            }
        }
    }
}

// ============================================================ TI:6776-7010
static void atm_init_coupled_diagnostics(Block& b, int time_lev) {
    const int nVertLevels = b.d.nVertLevels, index_qv = b.d.index_qv;
    const int cellStart = 1, cellEnd = b.d.nCells, edgeStart = 1, edgeEnd = b.d.nEdges;
    A2 theta_m = b.r2("theta_m", time_lev), rho_zz = b.r2("rho_zz", time_lev), u = b.r2("u", time_lev), w = b.r2("w", time_lev);
    A3 scalars = b.r3("scalars", time_lev);
    A2 theta = b.r2("theta"), rho = b.r2("rho"), zz = b.r2("zz"), ru = b.r2("ru"), rw = b.r2("rw"), rho_p = b.r2("rho_p"),
       rho_base = b.r2("rho_base"), rtheta_base = b.r2("rtheta_base"), theta_base = b.r2("theta_base"), rtheta_p = b.r2("rtheta_p"),
       exner = b.r2("exner"), exner_base = b.r2("exner_base"), pressure_p = b.r2("pressure_p"), pressure_base = b.r2("pressure_base");
    A1 fzm = b.r1("fzm"), fzp = b.r1("fzp");
    A3 zb_cell = b.r3("zb_cell"), zb3_cell = b.r3("zb3_cell");
    A2 edgesOnCell_sign = b.r2("edgesOnCell_sign");
    I2 cellsOnEdge = b.i2("cellsOnEdge"), edgesOnCell = b.i2("edgesOnCell");
    I1 nEdgesOnCell = b.i1("nEdgesOnCell");
    const real rcv = rgas / (cp - rgas);
    const real p0 = 1.e5;
    for (int iCell = cellStart; iCell <= cellEnd; iCell++)
        for (int k = 1; k <= nVertLevels; k++) {
            theta_m(k, iCell) = theta(k, iCell) * (1. + rvord * scalars(index_qv, k, iCell));
            rho_zz(k, iCell) = rho(k, iCell) / zz(k, iCell);
        }
    for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++) {
        int cell1 = cellsOnEdge(1, iEdge), cell2 = cellsOnEdge(2, iEdge);
        for (int k = 1; k <= nVertLevels; k++) ru(k, iEdge) = 0.5 * u(k, iEdge) * (rho_zz(k, cell1) + rho_zz(k, cell2));
    }
    for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
        rw(1, iCell) = 0.0;
        rw(nVertLevels + 1, iCell) = 0.0;
        for (int k = 2; k <= nVertLevels; k++)
            rw(k, iCell) = w(k, iCell)
                           * (fzp(k) * rho_zz(k - 1, iCell) + fzm(k) * rho_zz(k, iCell))
                           * (fzp(k) * zz(k - 1, iCell) + fzm(k) * zz(k, iCell));
    }
    for (int iCell = cellStart; iCell <= cellEnd; iCell++)
        for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
            int iEdge = edgesOnCell(i, iCell);
            for (int k = 2; k <= nVertLevels; k++) {
                real flux = (fzm(k) * ru(k, iEdge) + fzp(k) * ru(k - 1, iEdge));
                rw(k, iCell) = rw(k, iCell)
                               - edgesOnCell_sign(i, iCell) * (zb_cell(k, i, iCell) + sign1(flux) * zb3_cell(k, i, iCell)) * flux
                               * (fzp(k) * zz(k - 1, iCell) + fzm(k) * zz(k, iCell));
            }
        }
    for (int iCell = cellStart; iCell <= cellEnd; iCell++)
        for (int k = 1; k <= nVertLevels; k++) {
            rho_p(k, iCell) = rho_zz(k, iCell) - rho_base(k, iCell);
            rtheta_base(k, iCell) = theta_base(k, iCell) * rho_base(k, iCell);
            rtheta_p(k, iCell) = theta_m(k, iCell) * rho_p(k, iCell)
                                 + rho_base(k, iCell) * (theta_m(k, iCell) - theta_base(k, iCell));
            exner(k, iCell) = std::pow(zz(k, iCell) * (rgas / p0) * (rtheta_p(k, iCell) + rtheta_base(k, iCell)), rcv);
            exner_base(k, iCell) = std::pow(zz(k, iCell) * (rgas / p0) * (rtheta_base(k, iCell)), rcv);
            pressure_p(k, iCell) = zz(k, iCell) * rgas
                                   * (exner(k, iCell) * rtheta_p(k, iCell)
                                      + rtheta_base(k, iCell) * (exner(k, iCell) - exner_base(k, iCell)));
            pressure_base(k, iCell) = zz(k, iCell) * rgas * exner_base(k, iCell) * rtheta_base(k, iCell);
        }
}

// ============================================================ mpas_vector_reconstruction.F:205-330 (mpas_reconstruct_2d)
static void mpas_reconstruct(Block& b, int time_lev, bool includeHalos) {
    const int nVertLevels = b.d.nVertLevels;
    const int nCells = includeHalos ? b.d.nCells : b.d.nCellsSolve;
    I1 nEdgesOnCell = b.i1("nEdgesOnCell"); I2 edgesOnCell = b.i2("edgesOnCell");
    A3 coeffs_reconstruct = b.r3("coeffs_reconstruct");
    A1 latCell = b.r1("latCell"), lonCell = b.r1("lonCell");
    A2 u = b.r2("u", time_lev);
    A2 uReconstructX = b.r2("uReconstructX"), uReconstructY = b.r2("uReconstructY"), uReconstructZ = b.r2("uReconstructZ");
    A2 uReconstructZonal = b.r2("uReconstructZonal"), uReconstructMeridional = b.r2("uReconstructMeridional");
    #pragma omp parallel for
    for (int iCell = 1; iCell <= nCells; iCell++) {
        for (int k = 1; k <= nVertLevels; k++) {
            uReconstructX(k, iCell) = 0.0;
            uReconstructY(k, iCell) = 0.0;
            uReconstructZ(k, iCell) = 0.0;
        }
        for (int i = 1; i <= nEdgesOnCell(iCell); i++) {
            const int iEdge = edgesOnCell(i, iCell);
            for (int k = 1; k <= nVertLevels; k++) {
                uReconstructX(k, iCell) = uReconstructX(k, iCell) + coeffs_reconstruct(1, i, iCell) * u(k, iEdge);
                uReconstructY(k, iCell) = uReconstructY(k, iCell) + coeffs_reconstruct(2, i, iCell) * u(k, iEdge);
                uReconstructZ(k, iCell) = uReconstructZ(k, iCell) + coeffs_reconstruct(3, i, iCell) * u(k, iEdge);
            }
        }
    }
    if (b.c.on_a_sphere) {
        #pragma omp parallel for
        for (int iCell = 1; iCell <= nCells; iCell++) {
            const real clat = std::cos(latCell(iCell)), slat = std::sin(latCell(iCell));
            const real clon = std::cos(lonCell(iCell)), slon = std::sin(lonCell(iCell));
            for (int k = 1; k <= nVertLevels; k++) {
                uReconstructZonal(k, iCell) = -uReconstructX(k, iCell) * slon + uReconstructY(k, iCell) * clon;
                uReconstructMeridional(k, iCell) = -(uReconstructX(k, iCell) * clon + uReconstructY(k, iCell) * slon) * slat
                                                   + uReconstructZ(k, iCell) * clat;
            }
        }
    } else {
        for (int iCell = 1; iCell <= nCells; iCell++)
            for (int k = 1; k <= nVertLevels; k++) {
                uReconstructZonal(k, iCell) = uReconstructX(k, iCell);
                uReconstructMeridional(k, iCell) = uReconstructY(k, iCell);
            }
    }
}

// ============================================================ mpas_atm_core.F:901-950
static void atm_compute_output_diagnostics(Block& b, int time_lev) {
    const int nVertLevels = b.d.nVertLevels, nCells = b.d.nCells, index_qv = b.d.index_qv;
    A2 theta_m = b.r2("theta_m", time_lev), rho_zz = b.r2("rho_zz", time_lev), zz = b.r2("zz");
    A3 scalars = b.r3("scalars", time_lev);
    A2 theta = b.r2("theta"), rho = b.r2("rho"), pressure_p = b.r2("pressure_p"), pressure_base = b.r2("pressure_base"),
       pressure = b.r2("pressure");
    #pragma omp parallel for
    for (int iCell = 1; iCell <= nCells; iCell++)
        for (int k = 1; k <= nVertLevels; k++) {
            theta(k, iCell) = theta_m(k, iCell) / (1. + rvord * scalars(index_qv, k, iCell));
            rho(k, iCell) = rho_zz(k, iCell) * zz(k, iCell);
            pressure(k, iCell) = pressure_base(k, iCell) + pressure_p(k, iCell);
        }
}

// ============================================================ TI:7013-7191
static void atm_rk_dynamics_substep_finish(Block& b, int dynamics_substep, int dynamics_split) {
    const int nVertLevels = b.d.nVertLevels;
    const int cellStart = 1, cellEnd = b.d.nCells, edgeStart = 1, edgeEnd = b.d.nEdges;
    A2 ru = b.r2("ru"), ru_save = b.r2("ru_save"), rw = b.r2("rw"), rw_save = b.r2("rw_save"), rtheta_p = b.r2("rtheta_p"),
       rtheta_p_save = b.r2("rtheta_p_save"), rho_p = b.r2("rho_p"), rho_p_save = b.r2("rho_p_save"),
       rho_zz_old_split = b.r2("rho_zz_old_split"), ruAvg = b.r2("ruAvg"), wwAvg = b.r2("wwAvg"),
       ruAvg_split = b.r2("ruAvg_split"), wwAvg_split = b.r2("wwAvg_split");
    A2 u_1 = b.r2("u", 1), u_2 = b.r2("u", 2), w_1 = b.r2("w", 1), w_2 = b.r2("w", 2), theta_m_1 = b.r2("theta_m", 1),
       theta_m_2 = b.r2("theta_m", 2), rho_zz_1 = b.r2("rho_zz", 1), rho_zz_2 = b.r2("rho_zz", 2);
    for (int k = 1; k <= nVertLevels; k++) theta_m_1(k, cellEnd + 1) = 0.0;                 // TI:7082
    const real inv_dynamics_split = 1.0 / (real)dynamics_split;
    if (dynamics_substep < dynamics_split) {
        for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++)
            for (int k = 1; k <= nVertLevels; k++) { ru_save(k, iEdge) = ru(k, iEdge); u_1(k, iEdge) = u_2(k, iEdge); }
        for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
            for (int k = 1; k <= nVertLevels; k++) {
                rtheta_p_save(k, iCell) = rtheta_p(k, iCell);
                rho_p_save(k, iCell) = rho_p(k, iCell);
                theta_m_1(k, iCell) = theta_m_2(k, iCell);
                rho_zz_1(k, iCell) = rho_zz_2(k, iCell);
            }
            for (int k = 1; k <= nVertLevels + 1; k++) { rw_save(k, iCell) = rw(k, iCell); w_1(k, iCell) = w_2(k, iCell); }
        }
    }
    if (dynamics_substep == 1) {
        for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++)
            for (int k = 1; k <= nVertLevels; k++) ruAvg_split(k, iEdge) = ruAvg(k, iEdge);
        for (int iCell = cellStart; iCell <= cellEnd; iCell++)
            for (int k = 1; k <= nVertLevels + 1; k++) wwAvg_split(k, iCell) = wwAvg(k, iCell);
    } else {
        for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++)
            for (int k = 1; k <= nVertLevels; k++) ruAvg_split(k, iEdge) = ruAvg(k, iEdge) + ruAvg_split(k, iEdge);
        for (int iCell = cellStart; iCell <= cellEnd; iCell++)
            for (int k = 1; k <= nVertLevels + 1; k++) wwAvg_split(k, iCell) = wwAvg(k, iCell) + wwAvg_split(k, iCell);
    }
    if (dynamics_substep == dynamics_split) {
        for (int iEdge = edgeStart; iEdge <= edgeEnd; iEdge++)
            for (int k = 1; k <= nVertLevels; k++) ruAvg(k, iEdge) = ruAvg_split(k, iEdge) * inv_dynamics_split;
        for (int iCell = cellStart; iCell <= cellEnd; iCell++) {
            for (int k = 1; k <= nVertLevels + 1; k++) wwAvg(k, iCell) = wwAvg_split(k, iCell) * inv_dynamics_split;
            for (int k = 1; k <= nVertLevels; k++) rho_zz_1(k, iCell) = rho_zz_old_split(k, iCell);
        }
    }
}

// ============================================================ halo exchange
// Group table: src/core_atmosphere/mpas_atm_halos.F:211-298 (mpas_halo back-end).
struct GroupField { const char* name; int lev; int kind; int layers; /* bit mask of halo layers 1..3 */ };
struct Group { const char* name; std::vector<GroupField> fields; };
static const std::vector<Group>& halo_groups() {
    static const std::vector<Group> g = {
        {"dynamics:theta_m,scalars,pressure_p,rtheta_p", {{"theta_m", 1, 0, 3}, {"scalars", 1, 0, 3}, {"pressure_p", 1, 0, 3}, {"rtheta_p", 1, 0, 3}}},
        {"dynamics:exner", {{"exner", 1, 0, 3}}},
        {"dynamics:tend_u", {{"tend_u", 1, 1, 1}}},
        {"dynamics:rho_pp", {{"rho_pp", 1, 0, 1}}},
        {"dynamics:rtheta_pp", {{"rtheta_pp", 1, 0, 1}}},
        {"dynamics:rw_p,ru_p,rho_pp,rtheta_pp", {{"rw_p", 1, 0, 1}, {"ru_p", 1, 1, 2}, {"rho_pp", 1, 0, 3}, {"rtheta_pp", 1, 0, 2}}},
        {"dynamics:u_3", {{"u", 2, 1, 4}}},
        {"dynamics:w,pv_edge,rho_edge", {{"w", 2, 0, 3}, {"pv_edge", 1, 1, 3}, {"rho_edge", 1, 1, 3}}},
        {"dynamics:w,pv_edge,rho_edge,scalars", {{"w", 2, 0, 3}, {"pv_edge", 1, 1, 3}, {"rho_edge", 1, 1, 3}, {"scalars", 2, 0, 3}}},
        {"dynamics:theta_m,pressure_p,rtheta_p", {{"theta_m", 2, 0, 3}, {"pressure_p", 1, 0, 3}, {"rtheta_p", 1, 0, 3}}},
        {"dynamics:scalars_old", {{"scalars", 1, 0, 3}}},
        {"dynamics:scale", {{"scale_arr", 1, 0, 3}}},
        {"dynamics:scalars", {{"scalars", 2, 0, 3}}},
        {"dynamics:w", {{"w", 2, 0, 3}}},
        {"initialization:u", {{"u", 1, 1, 7}}},
        {"initialization:pv_edge,ru,rw", {{"pv_edge", 1, 1, 7}, {"ru", 1, 1, 7}, {"rw", 1, 0, 3}}},
    };
    return g;
}

struct Domain { std::vector<Block*> blocks; };

static void exchange_halo_group(Domain& dom, const char* name) {
    if (dom.blocks.size() == 1 && dom.blocks[0]->halo[0].empty()) return;
    const Group* grp = nullptr;
    for (const Group& g : halo_groups()) if (!strcmp(g.name, name)) grp = &g;
    if (!grp) { fprintf(stderr, "oracle: unknown halo group %s\n", name); abort(); }
    for (const GroupField& gf : grp->fields) {
        for (Block* dst : dom.blocks) {
            const FieldDef* fd = dst->defs.at(gf.name);
            const int inner = dst->inner1(fd->inner) * dst->inner2(fd->inner);
            for (const HaloList& rl : dst->halo[gf.kind]) {
                if (!(gf.layers & (1 << (rl.layer - 1)))) continue;
                Block* src = dom.blocks[rl.nbr];
                const HaloList* sl = nullptr;
                for (const HaloList& s : src->halo[gf.kind]) if (s.nbr == dst->rank && s.layer == rl.layer) sl = &s;
                if (!sl || sl->send_src.size() != rl.recv_dst.size()) {
                    if (rl.recv_dst.empty()) continue;
                    fprintf(stderr, "oracle: halo list mismatch\n"); abort();
                }
                real* d = dst->rp(gf.name, gf.lev);
                const real* s = src->rp(gf.name, gf.lev);
                for (size_t n = 0; n < rl.recv_dst.size(); n++)
                    memcpy(d + (size_t)(rl.recv_dst[n] - 1) * inner, s + (size_t)(sl->send_src[n] - 1) * inner, sizeof(real) * inner);
            }
        }
    }
}

// ============================================================ TI:803-1725
#define FOR_BLOCKS for (Block* bp : dom.blocks)
// advance_scalars (TI:1730-1927): plain RK stage, or the monotonic routine with its two exchange points on stage 3
static void advance_scalars(Domain& dom, int rk_step, real dt_rk) {
    const OCfg& c = dom.blocks[0]->c;
    if (rk_step < 3 || (!c.config_monotonic && !c.config_positive_definite)) {
        FOR_BLOCKS atm_advance_scalars(*bp, dt_rk, rk_step);
    } else {
        FOR_BLOCKS mono_pre_update(*bp, dt_rk);
        exchange_halo_group(dom, "dynamics:scalars_old");
        FOR_BLOCKS mono_rho_zz_int(*bp, dt_rk);
        for (int iScalar = 1; iScalar <= dom.blocks[0]->d.num_scalars; iScalar++) {
            FOR_BLOCKS mono_scalar_phase1(*bp, dt_rk, iScalar);
            exchange_halo_group(dom, "dynamics:scale");
            FOR_BLOCKS mono_scalar_phase2(*bp, dt_rk, iScalar);
        }
    }
}
static void atm_srk3(Domain& dom, real dt) {
    const OCfg& c = dom.blocks[0]->c;
    FOR_BLOCKS {   // TI:967-991, 1091-1093
        Block& b = *bp;
        std::fill(b.rf["qtot"].begin(), b.rf["qtot"].end(), 0.0);
        std::fill(b.rf["tend_ru_physics"].begin(), b.rf["tend_ru_physics"].end(), 0.0);
        std::fill(b.rf["tend_rtheta_physics"].begin(), b.rf["tend_rtheta_physics"].end(), 0.0);
        std::fill(b.rf["tend_rho_physics"].begin(), b.rf["tend_rho_physics"].end(), 0.0);
    }
    int dynamics_split = c.config_dynamics_split_steps;
    real dt_dynamics;
    if (c.config_split_dynamics_transport) dt_dynamics = dt / (real)dynamics_split;
    else { dynamics_split = 1; dt_dynamics = dt; }
    const int number_of_sub_steps = c.config_number_of_sub_steps;
    real rk_timestep[4], rk_sub_timestep[4];
    int number_sub_steps[4];
    if (c.config_time_integration_order == 3) {
        rk_timestep[1] = dt_dynamics / 3.; rk_timestep[2] = dt_dynamics / 2.; rk_timestep[3] = dt_dynamics;
        rk_sub_timestep[1] = dt_dynamics / 3.; rk_sub_timestep[2] = dt_dynamics / (real)number_of_sub_steps; rk_sub_timestep[3] = dt_dynamics / (real)number_of_sub_steps;
        number_sub_steps[1] = 1; number_sub_steps[2] = std::max(1, number_of_sub_steps / 2); number_sub_steps[3] = number_of_sub_steps;
    } else {
        rk_timestep[1] = dt_dynamics / 2.; rk_timestep[2] = dt_dynamics / 2.; rk_timestep[3] = dt_dynamics;
        rk_sub_timestep[1] = rk_sub_timestep[2] = rk_sub_timestep[3] = dt_dynamics / (real)number_of_sub_steps;
        number_sub_steps[1] = std::max(1, number_of_sub_steps / 2); number_sub_steps[2] = std::max(1, number_of_sub_steps / 2); number_sub_steps[3] = number_of_sub_steps;
    }
    exchange_halo_group(dom, "dynamics:theta_m,scalars,pressure_p,rtheta_p");
    FOR_BLOCKS atm_rk_integration_setup(*bp);
    FOR_BLOCKS atm_compute_moist_coefficients(*bp);
    for (int dynamics_substep = 1; dynamics_substep <= dynamics_split; dynamics_substep++) {
        FOR_BLOCKS atm_compute_vert_imp_coefs(*bp, rk_sub_timestep[1]);
        exchange_halo_group(dom, "dynamics:exner");
        for (int rk_step = 1; rk_step <= 3; rk_step++) {
            if (c.config_time_integration_order == 3 && rk_step == 2)
                FOR_BLOCKS atm_compute_vert_imp_coefs(*bp, rk_sub_timestep[rk_step]);
            FOR_BLOCKS atm_compute_dyn_tend(*bp, rk_step, dt);
            exchange_halo_group(dom, "dynamics:tend_u");
            FOR_BLOCKS atm_set_smlstep_pert_variables(*bp);
            for (int small_step = 1; small_step <= number_sub_steps[rk_step]; small_step++) {
                exchange_halo_group(dom, "dynamics:rho_pp");
                FOR_BLOCKS atm_advance_acoustic_step(*bp, rk_sub_timestep[rk_step], small_step);
                exchange_halo_group(dom, "dynamics:rtheta_pp");
                FOR_BLOCKS atm_divergence_damping_3d(*bp, rk_sub_timestep[rk_step]);
            }
            exchange_halo_group(dom, "dynamics:rw_p,ru_p,rho_pp,rtheta_pp");
            FOR_BLOCKS atm_recover_large_step_variables(*bp, rk_timestep[rk_step], number_sub_steps[rk_step], rk_step);
            exchange_halo_group(dom, "dynamics:u_3");
            const bool coupled_transport = c.config_scalar_advection && !c.config_split_dynamics_transport;
            if (coupled_transport) advance_scalars(dom, rk_step, rk_timestep[rk_step]);                  // TI:1404-1407
            FOR_BLOCKS atm_compute_solve_diagnostics(*bp, dt, 2, rk_step);
            exchange_halo_group(dom, coupled_transport ? "dynamics:w,pv_edge,rho_edge,scalars" : "dynamics:w,pv_edge,rho_edge");   // TI:1463-1473
        }
        if (dynamics_substep < dynamics_split) exchange_halo_group(dom, "dynamics:theta_m,pressure_p,rtheta_p");
        FOR_BLOCKS atm_rk_dynamics_substep_finish(*bp, dynamics_substep, dynamics_split);
    }
    if (c.config_scalar_advection && c.config_split_dynamics_transport) {
        rk_timestep[1] = dt / 3.; rk_timestep[2] = dt / 2.; rk_timestep[3] = dt;
        if (c.config_time_integration_order == 2) rk_timestep[1] = dt / 2.;
        for (int rk_step = 1; rk_step <= 3; rk_step++) {
            advance_scalars(dom, rk_step, rk_timestep[rk_step]);
            if (rk_step < 3) exchange_halo_group(dom, "dynamics:scalars");
        }
    }
    FOR_BLOCKS mpas_reconstruct(*bp, 2, false);                                   // TI:1596-1611
}

// ============================================================ C API (ctypes)
extern "C" {
void* oracle_create(const mpasb_dims* d, const mpasb_config* c, int rank) {
    Block* b = new Block();
    b->d = *d; b->c = OCfg(*c); b->rank = rank;
    b->init();
    return b;
}
void oracle_destroy(void* h) { delete (Block*)h; }
int oracle_real_bytes(void) { return (int)sizeof(real); }
long oracle_field_count(void* h, const char* name) {
    Block* b = (Block*)h;
    auto it = b->defs.find(name);
    return it == b->defs.end() ? -1 : b->count(it->second);
}
int oracle_set_field(void* h, const char* name, int lev, const real* src, long n) {
    Block* b = (Block*)h;
    auto it = b->rf.find(Block::key(name, lev));
    if (it == b->rf.end() || (long)it->second.size() != n) return 1;
    memcpy(it->second.data(), src, n * sizeof(real));
    return 0;
}
int oracle_get_field(void* h, const char* name, int lev, real* dst, long n) {
    Block* b = (Block*)h;
    auto it = b->rf.find(Block::key(name, lev));
    if (it == b->rf.end() || (long)it->second.size() != n) return 1;
    memcpy(dst, it->second.data(), n * sizeof(real));
    return 0;
}
int oracle_set_field_int(void* h, const char* name, const int* src, long n) {   // 1-based, like the ABI
    Block* b = (Block*)h;
    auto it = b->nf.find(name);
    if (it == b->nf.end() || (long)it->second.size() != n) return 1;
    memcpy(it->second.data(), src, n * sizeof(int));
    return 0;
}
int oracle_set_halo_lists(void* h, int kind, int n_neighbors, const int* neighbor_rank, int n_layers,
                          const int* n_send, const int* send_src, const int* n_recv, const int* recv_dst) {
    Block* b = (Block*)h;
    b->halo[kind].clear();
    long so = 0, ro = 0;
    for (int n = 0; n < n_neighbors; n++)
        for (int l = 0; l < n_layers; l++) {
            HaloList hl; hl.nbr = neighbor_rank[n]; hl.layer = l + 1;
            int ns = n_send[n * n_layers + l], nr = n_recv[n * n_layers + l];
            hl.send_src.assign(send_src + so, send_src + so + ns); so += ns;
            hl.recv_dst.assign(recv_dst + ro, recv_dst + ro + nr); ro += nr;
            b->halo[kind].push_back(hl);
        }
    return 0;
}
void oracle_shift_time_levels(void* h) {
    Block* b = (Block*)h;
    for (const char* n : {"u", "w", "rho_zz", "theta_m", "scalars"}) std::swap(b->rf[Block::key(n, 1)], b->rf[Block::key(n, 2)]);
}
static Domain make_domain(void** hs, int n) { Domain d; for (int i = 0; i < n; i++) d.blocks.push_back((Block*)hs[i]); return d; }
void oracle_step(void** hs, int n, double dt) { Domain d = make_domain(hs, n); atm_srk3(d, dt); }
void oracle_exchange(void** hs, int n, const char* group) { Domain d = make_domain(hs, n); exchange_halo_group(d, group); }
void oracle_minmax(void* h, double out[4]) {      // TI:8286-8319: reductions start from 0, owned elements only
    Block* b = (Block*)h;
    A2 w = b->r2("w", 2), u = b->r2("u", 2);
    real wmin = 0, wmax = 0, umin = 0, umax = 0;
    for (int i = 1; i <= b->d.nCellsSolve; i++) for (int k = 1; k <= b->d.nVertLevels; k++) { wmin = std::min(wmin, w(k, i)); wmax = std::max(wmax, w(k, i)); }
    for (int i = 1; i <= b->d.nEdgesSolve; i++) for (int k = 1; k <= b->d.nVertLevels; k++) { umin = std::min(umin, u(k, i)); umax = std::max(umax, u(k, i)); }
    out[0] = wmin; out[1] = wmax; out[2] = umin; out[3] = umax;
}
// one routine at a time (kernel-level parity)
void oracle_init_coupled_diagnostics(void* h) { atm_init_coupled_diagnostics(*(Block*)h, 1); }
void oracle_init_solve_diagnostics(void* h, double dt) { atm_compute_solve_diagnostics(*(Block*)h, dt, 1, 0); }
void oracle_rk_integration_setup(void* h) { atm_rk_integration_setup(*(Block*)h); }
void oracle_compute_moist_coefficients(void* h) { atm_compute_moist_coefficients(*(Block*)h); }
void oracle_compute_vert_imp_coefs(void* h, double dts) { atm_compute_vert_imp_coefs(*(Block*)h, dts); }
void oracle_compute_dyn_tend(void* h, int rk_step, double dt) { atm_compute_dyn_tend(*(Block*)h, rk_step, dt); }
void oracle_set_smlstep_pert_variables(void* h) { atm_set_smlstep_pert_variables(*(Block*)h); }
void oracle_advance_acoustic_step(void* h, double dts, int small_step) { atm_advance_acoustic_step(*(Block*)h, dts, small_step); }
void oracle_divergence_damping_3d(void* h, double dts) { atm_divergence_damping_3d(*(Block*)h, dts); }
void oracle_recover_large_step_variables(void* h, double dt, int ns, int rk_step) { atm_recover_large_step_variables(*(Block*)h, dt, ns, rk_step); }
void oracle_compute_solve_diagnostics(void* h, double dt, int rk_step) { atm_compute_solve_diagnostics(*(Block*)h, dt, 2, rk_step); }
void oracle_rk_dynamics_substep_finish(void* h, int s, int n) { atm_rk_dynamics_substep_finish(*(Block*)h, s, n); }
void oracle_advance_scalars(void* h, double dt, int rk_step) { atm_advance_scalars(*(Block*)h, dt, rk_step); }
// the monotonic transport split at its two exchange points (TI:4155, TI:4568), for drivers that
// perform the halo exchanges themselves (mpas_model_b200/multigpu.py: srk3_host_exchange)
void oracle_reconstruct(void* h, int time_lev, int include_halos) { mpas_reconstruct(*(Block*)h, time_lev, include_halos != 0); }
void oracle_compute_output_diagnostics(void* h, int time_lev) { atm_compute_output_diagnostics(*(Block*)h, time_lev); }
void oracle_advance_scalars_mono_pre(void* h, double dt) { mono_pre_update(*(Block*)h, dt); }
void oracle_advance_scalars_mono_a(void* h, double dt, int s) {
    Block& b = *(Block*)h;
    if (s == 0) mono_rho_zz_int(b, dt);
    mono_scalar_phase1(b, dt, s + 1);
}
void oracle_advance_scalars_mono_b(void* h, double dt, int s) { mono_scalar_phase2(*(Block*)h, dt, s + 1); }
void oracle_advance_scalars_mono(void* h, double dt) {
    Block& b = *(Block*)h;
    mono_pre_update(b, dt); mono_rho_zz_int(b, dt);
    for (int s = 1; s <= b.d.num_scalars; s++) { mono_scalar_phase1(b, dt, s); mono_scalar_phase2(b, dt, s); }
}
}
