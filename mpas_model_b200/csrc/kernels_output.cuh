// kernels_output.cuh -- the step's trailing velocity reconstruction and the host-visible output
// diagnostics.
//
//   mpas_reconstruct_2d              src/operators/mpas_vector_reconstruction.F:205-330 (called at TI:1606)
//   atm_compute_output_diagnostics   src/core_atmosphere/mpas_atm_core.F:901-950
//
// Both are pure streams: 1 E read + 5 C written, and 5 C read + 3 C written.  The reconstruction keeps the
// reference's accumulation order (edge 1..nEdgesOnCell, starting from 0) so that the X/Y/Z components equal
// the CPU arithmetic bit for bit; zonal/meridional go through cos/sin of the cell's latitude/longitude.
#pragma once
#include "kernels_col.cuh"

// column-warp version: one warp per cell, lane l owns the level pair (2l, 2l+1)
__global__ void __launch_bounds__(CW_THREADS) k2_reconstruct(const Dev D, const real* u, int ncells, int on_a_sphere) {
    CW_SETUP_R(ncells)
    const int ne = D.nEdgesOnCell[i];
    const int le = min(lane, ne - 1);
    const unsigned slot = (unsigned)i * D.maxEdges + le;
    const int my_e = D.edgesOnCell[slot];
    const real my_cx = D.coeffs_reconstruct[3 * slot], my_cy = D.coeffs_reconstruct[3 * slot + 1], my_cz = D.coeffs_reconstruct[3 * slot + 2];
    r2 ux = mk2(0.0, 0.0), uy = mk2(0.0, 0.0), uz = mk2(0.0, 0.0);
#define REC_U(E)                                                                                            \
    {                                                                                                       \
        const r2 uu = LD(u, BC(my_e, (E)));                                                                 \
        ux = selb((E) < ne, ux + BC(my_cx, (E)) * uu, ux);                                                  \
        uy = selb((E) < ne, uy + BC(my_cy, (E)) * uu, uy);                                                  \
        uz = selb((E) < ne, uz + BC(my_cz, (E)) * uu, uz);                                                  \
    }
#pragma unroll
    for (int e = 0; e < CW_NE; e++) REC_U(e)
    for (int e = CW_NE; e < ne; e++) REC_U(e)
#undef REC_U
    r2 uzon = ux, umer = uy;
    if (on_a_sphere) {
        const real lat = D.latCell[i], lon = D.lonCell[i];
        const real clat = cos(lat), slat = sin(lat), clon = cos(lon), slon = sin(lon);
        uzon = -ux * slon + uy * clon;
        umer = -(ux * clon + uy * slon) * slat + uz * clat;
    }
    const b2 k_lt_nl = lv.lt(nl);
    ST(D.uReconstructX, i, sel(k_lt_nl, ux, 0.0));
    ST(D.uReconstructY, i, sel(k_lt_nl, uy, 0.0));
    ST(D.uReconstructZ, i, sel(k_lt_nl, uz, 0.0));
    ST(D.uReconstructZonal, i, sel(k_lt_nl, uzon, 0.0));
    ST(D.uReconstructMeridional, i, sel(k_lt_nl, umer, 0.0));
}

// generic version: one thread per (level, cell), for columns taller than 64 levels
__global__ void k_reconstruct(const Dev D, const real* __restrict__ u, int ncells, int on_a_sphere) {
    KI;
    if (i >= ncells || k >= nl) return;
    const int ne = D.nEdgesOnCell[i];
    real ux = 0.0, uy = 0.0, uz = 0.0;
    for (int e = 0; e < ne; e++) {
        const size_t slot = (size_t)i * D.maxEdges + e;
        const real uu = AT(u, D.edgesOnCell[slot], k);
        ux = ux + D.coeffs_reconstruct[3 * slot] * uu;
        uy = uy + D.coeffs_reconstruct[3 * slot + 1] * uu;
        uz = uz + D.coeffs_reconstruct[3 * slot + 2] * uu;
    }
    real uzon = ux, umer = uy;
    if (on_a_sphere) {
        const real lat = D.latCell[i], lon = D.lonCell[i];
        const real clat = cos(lat), slat = sin(lat), clon = cos(lon), slon = sin(lon);
        uzon = -ux * slon + uy * clon;
        umer = -(ux * clon + uy * slon) * slat + uz * clat;
    }
    AT(D.uReconstructX, i, k) = ux;
    AT(D.uReconstructY, i, k) = uy;
    AT(D.uReconstructZ, i, k) = uz;
    AT(D.uReconstructZonal, i, k) = uzon;
    AT(D.uReconstructMeridional, i, k) = umer;
}

// atm_compute_output_diagnostics (mpas_atm_core.F:940-946): theta, rho, pressure over all cells of the block
__global__ void k_output_diagnostics(const Dev D, const real* __restrict__ theta_m, const real* __restrict__ rho_zz,
                                     const real* __restrict__ qv, real rvord) {
    KI;
    if (i >= D.nCells || k >= nl) return;
    AT(D.theta, i, k) = AT(theta_m, i, k) / (1. + rvord * AT(qv, i, k));
    AT(D.rho, i, k) = AT(rho_zz, i, k) * AT(D.zz, i, k);
    AT(D.pressure, i, k) = AT(D.pressure_base, i, k) + AT(D.pressure_p, i, k);
}
