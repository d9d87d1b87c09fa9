// mpasb_dev.cuh -- device-side view of one mesh block and small device helpers.
//
// Layout in HBM: every level-dimensioned field is [n+1][LDK] with the
// vertical index fastest ("level-contiguous"), LDK = nVertLevels+1 rounded up to an
// even count, shared by nVertLevels and nVertLevels+1 arrays so that one (k, column)
// thread index addresses all of them.  Columns are 16-byte aligned.  The trailing
// column n is the reference's garbage slot n+1 (mpas_block_creator.F:1050-1121).
// Connectivity is 0-based on the device.
#pragma once
#include <cuda_runtime.h>
#include <cstring>
#include "pow_cr.cuh"

// RKIND: PRECISION=double (default) or PRECISION=single (-DMPASB_SINGLE, reference Makefile: PRECISION=single)
#ifdef MPASB_SINGLE
typedef float real;
#else
typedef double real;
#endif

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// become resident while its predecessor in the stream is still draining.  pdl_trigger() lets the successor of THIS kernel
// do that (its blocks take the SM slots this kernel's last wave vacates); pdl_wait() returns once the predecessor grid has
// completed and its writes are visible.  Everything before pdl_wait() may only touch data no kernel of the step writes (mesh
// connectivity, weights); launched without the attribute both are no-ops.
#ifdef MPASB_EXPERIMENT_NOPDL
__device__ __forceinline__ void pdl_trigger() {}
__device__ __forceinline__ void pdl_wait() {}
#else
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
#define PDL_ENTER pdl_trigger(); pdl_wait();

enum { LOC_CELL = 0, LOC_EDGE = 1, LOC_VERTEX = 2, LOC_LEVS = 3 };
enum { IN_ONE, IN_NL, IN_NL1, IN_ME, IN_ME2, IN_VD, IN_TWO, IN_F15, IN_NL1_ME, IN_S_NL, IN_NL_TWO, IN_THREE_ME };

#define FIELD_REAL(name) real* name; real* name##_2;
#define FIELD_INT(name) int* name;
#define F(name, loc, inner, lev, type, tgt) FIELD_##type(name)

struct Dev {
#include "../../include/mpasb_fields.def"
    // dimensions (0-based counts; garbage column index == n)
    int nCells, nEdges, nVertices, nCellsSolve, nEdgesSolve, nVerticesSolve;
    int nl;            // nVertLevels
    int LDK;           // column pitch in reals (>= LDKA; MPASB_LDK_ALIGN rounds it up, e.g. to whole 128-byte lines)
    int LDKA;          // active width: nVertLevels+1 rounded up to an even count; rows [LDKA, LDK) are never touched
    int maxEdges, maxEdges2, num_scalars;
    int index_qv, moist_start, moist_end;      // 0-based
    int apply_lbcs;    // config_apply_lbcs: regional run (kernels_lbc.cuh and the bdyMask branches of the work routines)
    size_t cellPlane, edgePlane;               // (n+1)*LDK, stride between scalar planes
    // derived, library-internal: 1 where any zb_cell/zb3_cell entry of the cell is non-zero (terrain slope);
    // cells with 0 skip the 2 x maxEdges x nVertLevels metric reads of TI:2480-2500 and TI:3379-3414
    int* zb_any;
    int pf_next;       // the persistent kernels (k6_acoustic_cell, k7_dt_cell_f) ask L2 for the own-column operands of their next cell (MPASB_PF_NEXT=0: off)
    // per-edge 3rd/4th-order advective fluxes of w and theta_m (kernels_col.cuh: k2_dt_edge_flux -> k2_dt_cell_f)
    real* adv_flux_w; real* adv_flux_theta;
    // cell-centred flux sweep of the relaxed-arithmetic path (kernels_col.cuh: k5_flux_cell -> k2_dt_cell_f): per owned cell
    // the 18 cells of its two rings in canonical order + a regularity flag, the 2 x 60 stencil weights of its 6 edges in
    // canonical slot order, and the horizontal flux divergences of w and theta_m it produces
    int* fx_ring; real* fx_w; real* hdiv_w; real* hdiv_theta;
    // cell-centred partial sums of the nonlinear Coriolis term (k8_coriolis_cell -> k2_dt_edge_b<true>): per cell the weights
    // W[i][j] of edge j of the cell in the edgesOnEdge list of its edge i (8 x 8 reals), per edge its slot in its two cells,
    // and the partial sums [cell][slot][LDK]
    real* cor_w; int* cor_slot; real* cor_part;
    // monotonic transport batched over scalars (advance_scalars_mono): the per-scalar work arrays of TI:4220-4719 exist in
    // mb_planes copies, one per scalar, so that each of its five kernels runs ONCE for all scalars (gridDim.y = scalar) and
    // `scale_arr` of every scalar travels in ONE exchange; the arrays of the field table are the LAST plane of each (what the
    // reference's arrays hold when the routine returns).  mb_planes == 1: not batched, the bases are the table arrays.
    int mb_planes;
    real *mb_wdtn, *mb_s_max, *mb_s_min, *mb_scalar_new, *mb_scale, *mb_flux_tmp, *mb_flux_upwind_tmp, *mb_flux_arr;
};
#undef F
#undef FIELD_REAL
#undef FIELD_INT

// src/framework/mpas_constants.F:43-56
#define GRAVITY 9.80616
#define RGAS 287.0
#define CP_ (7.0 * 287.0 / 2.0)
#define RV_ 461.6
#define PRANDTL 1.0

// min/max in RKIND whatever the literal types of the arguments (fmax(0.0, x) would not resolve in the single build)
#ifdef MPASB_SINGLE
__device__ __forceinline__ real rmax(real a, real b) { return fmaxf(a, b); }
__device__ __forceinline__ real rmin(real a, real b) { return fminf(a, b); }
#else
__device__ __forceinline__ real rmax(real a, real b) { return fmax(a, b); }
__device__ __forceinline__ real rmin(real a, real b) { return fmin(a, b); }
#endif
__device__ __forceinline__ real sign1(real x) { return copysign(1.0, x); }      // Fortran sign(1.0, x)

// statement functions, mpas_atm_time_integration.F:5156-5161
__device__ __forceinline__ real flux4(real q_im2, real q_im1, real q_i, real q_ip1, real ua) {
    return ua * (7. * (q_i + q_im1) - (q_ip1 + q_im2)) / 12.0;
}
__device__ __forceinline__ real flux3(real q_im2, real q_im1, real q_i, real q_ip1, real ua, real coef3) {
    return flux4(q_im2, q_im1, q_i, q_ip1, ua) + coef3 * fabs(ua) * ((q_ip1 - q_im2) - 3. * (q_i - q_im1)) / 12.0;
}
