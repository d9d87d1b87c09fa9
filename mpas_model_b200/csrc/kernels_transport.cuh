// kernels_transport.cuh -- scalar transport: RK stages without limiter
// (atm_advance_scalars_work, TI:3575-3855) and the monotonic flux-corrected final
// stage (atm_advance_scalars_mono_work, TI:4012-4734).
// Scalars live on the device as num_scalars separate level-contiguous planes
// [S][nCells+1][LDK]; the reference's scalar-fastest (S, k, cell) order is converted at the ABI.
#pragma once
#include "kernels_dyn.cuh"

// ---- edge value of every scalar ("horiz_flux_arr"), TI:3670-3751
__global__ void k_scalars_edge(const Dev D) {
    KI;
    if (i >= D.nEdges || k >= nl) return;
    const int nadv = D.nAdvCellsForEdge[i];
    IP adv = D.advCellsForEdge + (size_t)i * 15;
    RP ac = D.adv_coefs + (size_t)i * 15;
    RP ac3 = D.adv_coefs_3rd + (size_t)i * 15;
    const real su = sign1(AT(D.ruAvg, i, k));
    if (D.apply_lbcs && D.bdyMaskEdge[i] >= 4) {            // regional run, TI:3732-3750: the two outermost relaxation-zone edges
        if (D.bdyMaskEdge[i] > 5) return;                    // take the upwind value, edges of the specified zone nothing
        const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
        const real u_direction = copysign((real)0.5, AT(D.ruAvg, i, k));
        const real u_positive = D.dvEdge[i] * fabs(u_direction + 0.5);
        const real u_negative = D.dvEdge[i] * fabs(u_direction - 0.5);
        for (int s = 0; s < D.num_scalars; s++) {
            RP q = D.scalars_2 + (size_t)s * D.cellPlane;
            AT(D.horiz_flux_arr + (size_t)s * D.edgePlane, i, k) = u_positive * AT(q, cell1, k) + u_negative * AT(q, cell2, k);
        }
        return;
    }
    for (int s = 0; s < D.num_scalars; s++) {
        RP q = D.scalars_2 + (size_t)s * D.cellPlane;
        real acc; int j0;
        if (nadv == 10) { acc = (ac[0] + su * ac3[0]) * AT(q, adv[0], k); j0 = 1; }   // single expression, TI:3694-3704
        else { acc = 0.0; j0 = 0; }                                                     // accumulation from 0, TI:3711-3729
        for (int j = j0; j < nadv; j++) {
            const real scalar_weight = ac[j] + su * ac3[j];
            acc = acc + scalar_weight * AT(q, adv[j], k);
        }
        AT(D.horiz_flux_arr + (size_t)s * D.edgePlane, i, k) = acc;
    }
}

// vertical scalar flux at interface kk (Fortran kk+1), TI:3817-3833 / 4280-4302
__device__ __forceinline__ real wdtn_raw(const real* __restrict__ q, const real* __restrict__ ww, const Dev& D,
                                         int kk, int iCell, real coef3, int LDK, int nl) {
    if (kk <= 0 || kk >= nl) return 0.0;
    if (kk == 1 || kk == nl - 1)
        return AT(ww, iCell, kk) * (D.fzm[kk] * AT(q, iCell, kk) + D.fzp[kk] * AT(q, iCell, kk - 1));
    return flux3(AT(q, iCell, kk - 2), AT(q, iCell, kk - 1), AT(q, iCell, kk), AT(q, iCell, kk + 1), AT(ww, iCell, kk), coef3);
}

// ---- owned cells: flux divergence + vertical flux + update, TI:3773-3846
__global__ void k_scalars_cell(const Dev D, real dt, real weight_time_old, real weight_time_new, real coef3) {
    KI;
    const bool act = i < D.nCellsSolve && k < nl && D.bdyMaskCell[min(i, D.nCells)] <= 5;       // TI:3775: not the specified zone
    const int ne = act ? D.nEdgesOnCell[i] : 0;
    const real rho_old = act ? AT(D.rho_zz, i, k) : 1.0, rho_new = act ? AT(D.rho_zz_2, i, k) : 1.0;
    const real rho_zz_new_inv = 1.0 / (weight_time_old * rho_old + weight_time_new * rho_new);
    for (int s = 0; s < D.num_scalars; s++) {
        real* qn = D.scalars_2 + (size_t)s * D.cellPlane;
        real val = 0.0;
        if (act) {
            RP hf = D.horiz_flux_arr + (size_t)s * D.edgePlane;
            real tend = 0.0;
            for (int e = 0; e < ne; e++) {
                const int iEdge = D.edgesOnCell[(size_t)i * D.maxEdges + e];
                tend = tend - D.edgesOnCell_sign[(size_t)i * D.maxEdges + e] * AT(D.ruAvg, iEdge, k) * AT(hf, iEdge, k);
            }
            tend = tend * D.invAreaCell[i] + 0.0;        // + scalar_tend_save, zero without physics (TI:3781-3783)
            AT(D.scalars_tend + (size_t)s * D.cellPlane, i, k) = 0.0;
            const real w0 = wdtn_raw(qn, D.wwAvg, D, k, i, coef3, LDK, nl);
            const real w1 = wdtn_raw(qn, D.wwAvg, D, k + 1, i, coef3, LDK, nl);
            val = (AT(D.scalars + (size_t)s * D.cellPlane, i, k) * rho_old
                   + dt * (tend - D.rdzw[k] * (w1 - w0))) * rho_zz_new_inv;
        }
        __syncthreads();        // every read of scalar_new in this column precedes its update
        if (act) AT(qn, i, k) = val;
    }
}

// ================================================================== monotonic transport
// (A) TI:4129-4143: physics-tendency pre-update of scalars_old (zero tendency without physics)
__global__ void k_mono_pre(const Dev D, real dt) {
    KI;
    if (i >= D.nCellsSolve || k >= nl) return;
    const real rho_old = AT(D.rho_zz, i, k);
    for (int s = 0; s < D.num_scalars; s++) {
        real* q = D.scalars + (size_t)s * D.cellPlane;
        real* t = D.scalars_tend + (size_t)s * D.cellPlane;
        const real st = 0.0;
        AT(q, i, k) = AT(q, i, k) + dt * st / rho_old;
        AT(t, i, k) = 0.0;
    }
}
// (B) TI:4177-4204: re-integrated density
__global__ void k_mono_rho_int(const Dev D, real dt) {
    KI;
    if (i >= D.nCellsSolve || k >= nl) return;
    const int ne = D.nEdgesOnCell[i];
    const real invArea = D.invAreaCell[i];
    real r = 0.0;
    for (int e = 0; e < ne; e++) {
        const int iEdge = D.edgesOnCell[(size_t)i * D.maxEdges + e];
        r = r - D.edgesOnCell_sign[(size_t)i * D.maxEdges + e]
                * AT(D.ruAvg, iEdge, k) * D.dvEdge[iEdge] * invArea;
    }
    AT(D.rho_zz_int, i, k) = AT(D.rho_zz, i, k) + dt * (r - D.rdzw[k] * (AT(D.wwAvg, i, k + 1) - AT(D.wwAvg, i, k)));
}

__device__ __forceinline__ real mono_wdtn(const real* __restrict__ so, const real* __restrict__ sn, const Dev& D,
                                          int kk, int iCell, real dt, real coef3, int LDK, int nl) {
    if (kk <= 0 || kk >= nl) return 0.0;
    const real ww = AT(D.wwAvg, iCell, kk);
    const real fu = dt * (rmax(0.0, ww) * AT(so, iCell, kk - 1) + rmin(0.0, ww) * AT(so, iCell, kk));
    return dt * wdtn_raw(sn, D.wwAvg, D, kk, iCell, coef3, LDK, nl) - fu;
}

// (C1) owned cells: vertical fluxes, bounds, vertical part of the upwind update and of scale_arr  TI:4277-4344, 4426-4459
__global__ void k_mono_cell1(const Dev D, int s, real dt, real coef3) {
    KI;
    if (i >= D.nCellsSolve || k > nl) return;
    RP so = D.scalars + (size_t)s * D.cellPlane;
    RP sn = D.scalars_2 + (size_t)s * D.cellPlane;
    const real wd0 = mono_wdtn(so, sn, D, k, i, dt, coef3, LDK, nl);
    AT(D.wdtn, i, k) = wd0;
    if (k == nl) return;
    const real wd1 = mono_wdtn(so, sn, D, k + 1, i, dt, coef3, LDK, nl);
    const real sok = AT(so, i, k);
    real smax, smin;
    if (k == 0) { smax = rmax(sok, AT(so, i, 1)); smin = rmin(sok, AT(so, i, 1)); }
    else if (k == nl - 1) { smax = rmax(sok, AT(so, i, k - 1)); smin = rmin(sok, AT(so, i, k - 1)); }
    else { smax = rmax(rmax(AT(so, i, k - 1), sok), AT(so, i, k + 1)); smin = rmin(rmin(AT(so, i, k - 1), sok), AT(so, i, k + 1)); }
    const int ne = D.nEdgesOnCell[i];
    for (int e = 0; e < ne; e++) {
        const real v = AT(so, D.cellsOnCell[(size_t)i * D.maxEdges + e], k);
        smax = rmax(smax, v); smin = rmin(smin, v);
    }
    AT(D.s_max, i, k) = smax; AT(D.s_min, i, k) = smin;
    // upwind vertical update, TI:4428-4446
    real snew = sok * AT(D.rho_zz, i, k);
    const real rdnw = D.rdzw[k];
    if (k <= nl - 2) {
        const real ww = AT(D.wwAvg, i, k + 1);
        const real fu1 = dt * (rmax(0.0, ww) * sok + rmin(0.0, ww) * AT(so, i, k + 1));
        snew = snew - fu1 * rdnw;
    }
    if (k >= 1) {
        const real ww = AT(D.wwAvg, i, k);
        const real fu0 = dt * (rmax(0.0, ww) * AT(so, i, k - 1) + rmin(0.0, ww) * sok);
        snew = snew + fu0 * rdnw;
    }
    AT(D.scalar_new, i, k) = snew;
    AT(D.scale_arr, i, k) = -rdnw * (rmin(0.0, wd1) - rmax(0.0, wd0));                      // SCALE_IN
    AT(D.scale_arr + D.cellPlane, i, k) = -rdnw * (rmax(0.0, wd1) - rmin(0.0, wd0));        // SCALE_OUT
}
// (C2) edges: high-order flux (4356-4413), upwind flux and their difference (4467-4487)
__global__ void k_mono_edge2(const Dev D, int s, real dt) {
    KI;
    if (i >= D.nEdges || k >= nl) return;
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    RP so = D.scalars + (size_t)s * D.cellPlane;
    RP sn = D.scalars_2 + (size_t)s * D.cellPlane;
    const real uh = AT(D.ruAvg, i, k);
    real flux = 0.0;
    if (cell1 < D.nCellsSolve || cell2 < D.nCellsSolve) {
        const int nadv = D.nAdvCellsForEdge[i];
        IP adv = D.advCellsForEdge + (size_t)i * 15;
        RP ac = D.adv_coefs + (size_t)i * 15;
        RP ac3 = D.adv_coefs_3rd + (size_t)i * 15;
        if (nadv == 10) {
            const bool pos = uh > 0;
            real acc = (pos ? ac[0] + ac3[0] : ac[0] - ac3[0]) * AT(sn, adv[0], k);
            for (int j = 1; j < 10; j++) acc = acc + (pos ? ac[j] + ac3[j] : ac[j] - ac3[j]) * AT(sn, adv[j], k);
            flux = uh * (acc);
        } else {
            const real su = sign1(uh);
            for (int j = 0; j < nadv; j++) {
                const real scalar_weight = uh * (ac[j] + su * ac3[j]);
                flux = flux + scalar_weight * AT(sn, adv[j], k);
            }
        }
    }
    const real fup = D.dvEdge[i] * dt * (rmax(0.0, uh) * AT(so, cell1, k) + rmin(0.0, uh) * AT(so, cell2, k));
    AT(D.flux_upwind_tmp, i, k) = fup;
    // TI:4479 (and :4592), as written there: `config_apply_lbcs .and. (m == nRelaxZone) .or. (m == nRelaxZone-1)`
    const int m = D.bdyMaskEdge[i];
    const bool upwind_only = (D.apply_lbcs && m == 5) || m == 4;
    AT(D.flux_tmp, i, k) = upwind_only ? (real)0.0 : dt * flux - fup;
}
// (C3) owned cells: horizontal part of the upwind update and of scale_arr (4496-4513) and the limiter (4523-4553)
__global__ void k_mono_cell3(const Dev D, const real* __restrict__ rho_lim) {
    KI;
    if (i >= D.nCellsSolve || k >= nl) return;
    const int ne = D.nEdgesOnCell[i];
    const real invArea = D.invAreaCell[i];
    real snew = AT(D.scalar_new, i, k);
    real sin_ = AT(D.scale_arr, i, k), sout = AT(D.scale_arr + D.cellPlane, i, k);
    for (int e = 0; e < ne; e++) {
        const int iEdge = D.edgesOnCell[(size_t)i * D.maxEdges + e];
        const real sg = D.edgesOnCell_sign[(size_t)i * D.maxEdges + e];
        const real ft = AT(D.flux_tmp, iEdge, k);
        snew = snew - sg * AT(D.flux_upwind_tmp, iEdge, k) * invArea;
        sout = sout - rmax(0.0, sg * ft) * invArea;
        sin_ = sin_ - rmin(0.0, sg * ft) * invArea;
    }
    AT(D.scalar_new, i, k) = snew;
    const real eps = 1.e-20;
    const real rl = AT(rho_lim, i, k);
    real scale_factor = (AT(D.s_max, i, k) * rl - snew) / (sin_ + eps);
    AT(D.scale_arr, i, k) = rmin(1.0, rmax(0.0, scale_factor));
    scale_factor = (AT(D.s_min, i, k) * rl - snew) / (sout - eps);
    AT(D.scale_arr + D.cellPlane, i, k) = rmin(1.0, rmax(0.0, scale_factor));
}
// (D1) edges of owned cells: rescale the anti-diffusive flux (4579-4623)
__global__ void k_mono_edge4(const Dev D) {
    KI;
    if (i >= D.nEdges || k >= nl) return;
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    if (!(cell1 < D.nCellsSolve || cell2 < D.nCellsSolve)) return;
    RP s_in = D.scale_arr; RP s_out = D.scale_arr + D.cellPlane;
    real flux = AT(D.flux_tmp, i, k);
    flux = rmax(0.0, flux) * rmin(AT(s_out, cell1, k), AT(s_in, cell2, k))
         + rmin(0.0, flux) * rmin(AT(s_in, cell1, k), AT(s_out, cell2, k));
    AT(D.flux_arr, i, k) = flux;
}
__device__ __forceinline__ real mono_wdtn_scaled(const Dev& D, int kk, int iCell, int LDK, int nl) {
    if (kk <= 0 || kk >= nl) return 0.0;
    RP s_in = D.scale_arr; RP s_out = D.scale_arr + D.cellPlane;
    real flux = AT(D.wdtn, iCell, kk);
    flux = rmax(0.0, flux) * rmin(AT(s_out, iCell, kk - 1), AT(s_in, iCell, kk))
         + rmin(0.0, flux) * rmin(AT(s_out, iCell, kk), AT(s_in, iCell, kk - 1));
    return flux;
}
// (D2) all cells: rescaled vertical flux (4636-4645), final update (4651-4674), positive-definite copy-out (4708-4715)
__global__ void k_mono_cell5(const Dev D, int s, const real* __restrict__ rho_div) {
    KI;
    if (i >= D.nCells || k >= nl) return;
    real* out = D.scalars_2 + (size_t)s * D.cellPlane;
    if (D.bdyMaskCell[i] > 2) return;              // TI:4709 `bdyMaskCell <= nSpecZone`: these cells are set after the transport
    if (i >= D.nCellsSolve) { AT(out, i, k) = rmax(0.0, AT(out, i, k)); return; }
    const real w0 = mono_wdtn_scaled(D, k, i, LDK, nl), w1 = mono_wdtn_scaled(D, k + 1, i, LDK, nl);
    const int ne = D.nEdgesOnCell[i];
    const real invArea = D.invAreaCell[i];
    real snew = AT(D.scalar_new, i, k);
    for (int e = 0; e < ne; e++) {
        const int iEdge = D.edgesOnCell[(size_t)i * D.maxEdges + e];
        snew = snew - D.edgesOnCell_sign[(size_t)i * D.maxEdges + e] * AT(D.flux_arr, iEdge, k) * invArea;
    }
    snew = (snew + (-D.rdzw[k] * (w1 - w0))) / AT(rho_div, i, k);
    AT(out, i, k) = rmax(0.0, snew);
}
