// kernels_dyn.cuh -- large-step tendency kernels and small per-step coefficient kernels.
// One thread per (level k, column i); block = (LDK, columns-per-block).  Every
// expression keeps the reference's operation order (TI = mpas_atm_time_integration.F)
// so that results equal the fp64 CPU arithmetic bit for bit when built --fmad=false.
#pragma once
#include "mpasb_dev.cuh"

#define KI const int k = threadIdx.x; const int i = blockIdx.x * blockDim.y + threadIdx.y; \
           const int LDK = D.LDK; const int nl = D.nl; (void)LDK; (void)nl;
// 32-bit element offsets: a block holds < 2^32 reals per field, and unsigned wrap-around keeps k-1, k-2 exact
#define AT(p, i, k) (p)[(unsigned)((unsigned)(i) * (unsigned)LDK + (unsigned)(k))]
#define RP const real* __restrict__
#define IP const int* __restrict__

// ------------------------------------------------------------------ atm_compute_moist_coefficients  TI:2042-2146
__global__ void k_moist_cell(const Dev D) {
    KI;
    if (i >= D.nCells || k >= nl) return;
    real q = 0.0, qm = 0.0;
    for (int iq = D.moist_start; iq <= D.moist_end; iq++) {
        RP s = D.scalars_2 + (size_t)iq * D.cellPlane;
        q = q + AT(s, i, k);
        if (k >= 1) qm = qm + AT(s, i, k - 1);
    }
    AT(D.qtot, i, k) = q;
    if (k >= 1) {
        real qtotal = 0.5 * (q + qm);
        AT(D.cqw, i, k) = 1.0 / (1.0 + qtotal);
    }
}

__global__ void k_moist_edge(const Dev D) {
    KI;
    if (i >= D.nEdges || k >= nl) return;
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    if (cell1 < D.nCellsSolve || cell2 < D.nCellsSolve) {
        real qtotal = 0.0;
        for (int iq = D.moist_start; iq <= D.moist_end; iq++) {
            RP s = D.scalars_2 + (size_t)iq * D.cellPlane;
            qtotal = qtotal + 0.5 * (AT(s, cell1, k) + AT(s, cell2, k));
        }
        AT(D.cqu, i, k) = 1.0 / (1.0 + qtotal);
    }
}

// ------------------------------------------------------------------ atm_compute_vert_imp_coefs_work  TI:2225-2366
// Shared memory per column: cofwz, cofwr, coftz, cofwt, a, b, c, alpha, gamma (9 * LDK reals).
__global__ void k_vert_imp_coefs(const Dev D, real dtseps, real c2, real rcv) {
    KI;
    extern __shared__ real sm[];
    real* s_cofwz = sm + (size_t)threadIdx.y * 9 * LDK;
    real* s_cofwr = s_cofwz + LDK;
    real* s_coftz = s_cofwr + LDK;
    real* s_cofwt = s_coftz + LDK;
    real* s_a = s_cofwt + LDK;
    real* s_b = s_a + LDK;
    real* s_c = s_b + LDK;
    real* s_alpha = s_c + LDK;
    real* s_gamma = s_alpha + LDK;
    const bool col = i < D.nCellsSolve;
    RP fzm = D.fzm; RP fzp = D.fzp; RP rdzw = D.rdzw; RP rdzu = D.rdzu;
    if (i == 0 && k < nl) D.cofrz[k] = dtseps * rdzw[k];
    s_cofwz[k] = 0.0; s_cofwr[k] = 0.0; s_coftz[k] = 0.0; s_cofwt[k] = 0.0;
    if (col) {
        if (k >= 1 && k < nl) {
            const real zzk = AT(D.zz, i, k), zzm = AT(D.zz, i, k - 1);
            const real cofwr = .5 * dtseps * GRAVITY * (fzm[k] * zzk + fzp[k] * zzm);
            const real cofwz = dtseps * c2 * (fzm[k] * zzk + fzp[k] * zzm)
                               * rdzu[k] * AT(D.cqw, i, k) * (fzm[k] * AT(D.exner, i, k) + fzp[k] * AT(D.exner, i, k - 1));
            const real coftz = dtseps * (fzm[k] * AT(D.theta_m_2, i, k) + fzp[k] * AT(D.theta_m_2, i, k - 1));
            s_cofwr[k] = cofwr; s_cofwz[k] = cofwz; s_coftz[k] = coftz;
            AT(D.cofwr, i, k) = cofwr; AT(D.cofwz, i, k) = cofwz; AT(D.coftz, i, k) = coftz;
        }
        if (k == 0 || k == nl) AT(D.coftz, i, k) = 0.0;
        if (k < nl) {
            const real qtotal = AT(D.qtot, i, k);
            const real cofwt = .5 * dtseps * rcv * AT(D.zz, i, k) * GRAVITY * AT(D.rho_base, i, k) / (1. + qtotal)
                               * AT(D.exner, i, k) / ((AT(D.rtheta_base, i, k) + AT(D.rtheta_p, i, k)) * AT(D.exner_base, i, k));
            s_cofwt[k] = cofwt;
            AT(D.cofwt, i, k) = cofwt;
        }
    }
    __syncthreads();
    if (col && k >= 1 && k < nl) {
        const real cofrz_k = dtseps * rdzw[k], cofrz_m = dtseps * rdzw[k - 1];
        const real zzk = AT(D.zz, i, k), zzm = AT(D.zz, i, k - 1);
        const real a = -s_cofwz[k] * s_coftz[k - 1] * rdzw[k - 1] * zzm
                       + s_cofwr[k] * cofrz_m
                       - s_cofwt[k - 1] * s_coftz[k - 1] * rdzw[k - 1];
        const real b = 1.
                       + s_cofwz[k] * (s_coftz[k] * rdzw[k] * zzk
                                       + s_coftz[k] * rdzw[k - 1] * zzm)
                       - s_coftz[k] * (s_cofwt[k] * rdzw[k]
                                       - s_cofwt[k - 1] * rdzw[k - 1])
                       + s_cofwr[k] * (cofrz_k - cofrz_m);
        const real c = -s_cofwz[k] * s_coftz[k + 1] * rdzw[k] * zzk
                       - s_cofwr[k] * cofrz_k
                       + s_cofwt[k] * s_coftz[k + 1] * rdzw[k];
        s_a[k] = a; s_b[k] = b; s_c[k] = c;
        AT(D.a_tri, i, k) = a;
    }
    __syncthreads();
    if (col && k == 0) {
        AT(D.a_tri, i, 0) = 0.;
        real g = 0.;
        s_gamma[0] = 0.; s_alpha[0] = 0.;
        for (int kk = 1; kk < nl; kk++) {
            const real al = 1. / (s_b[kk] - s_a[kk] * g);
            g = s_c[kk] * al;
            s_alpha[kk] = al; s_gamma[kk] = g;
        }
    }
    __syncthreads();
    if (col && k < nl) { AT(D.alpha_tri, i, k) = s_alpha[k]; AT(D.gamma_tri, i, k) = s_gamma[k]; }
}

// ------------------------------------------------------------------ atm_compute_dyn_tend_work  TI:4982-6240
struct DynTendArgs {
    int rk_step;
    int smag;                 // config_horiz_mixing == "2d_smagorinsky"
    real invDt, cs_len2, kdiff_cap, fixed_visc2;      // (c_s*len_disp)**2, (0.01*len_disp**2)*invDt
    real h_mom_eddy_visc4, h_theta_eddy_visc4, del4u_div_factor;
    real v_mom_eddy_visc2, v_theta_eddy_visc2;
    real coef_3rd_order, prandtl_inv;
    int mix_full;
    real cam_coef, len_disp; int n_cam_levels;
    int rayleigh_damp_u, n_rayleigh_levels; real rayleigh_coef_inverse;
};

// (a) cell-all: Smagorinsky kdiff (rk 1, TI:5226-5296), h_divergence (5307-5338), tend_rho + dpdz (rk 1, 5345-5362)
__global__ void k_dt_cell_a(const Dev D, const DynTendArgs A) {
    KI;
    if (i >= D.nCells || k >= nl) return;
    const int ne = D.nEdgesOnCell[i];
    IP eoc = D.edgesOnCell + (size_t)i * D.maxEdges;
    RP sgn = D.edgesOnCell_sign + (size_t)i * D.maxEdges;
    if (A.rk_step == 1) {
        real kd;
        if (A.smag) {
            real d_diag = 0.0, d_off_diag = 0.0;
            RP da = D.defc_a + (size_t)i * D.maxEdges;
            RP db = D.defc_b + (size_t)i * D.maxEdges;
            for (int e = 0; e < ne; e++) {
                const real uu = AT(D.u_2, eoc[e], k), vv = AT(D.v, eoc[e], k);
                d_diag = d_diag + da[e] * uu - db[e] * vv;
                d_off_diag = d_off_diag + db[e] * uu + da[e] * vv;
            }
            kd = rmin(A.cs_len2 * sqrt(d_diag * d_diag + d_off_diag * d_off_diag), A.kdiff_cap);
        } else {
            kd = A.fixed_visc2;
        }
        if (A.cam_coef > 0.0 && k >= nl - A.n_cam_levels) {       // TI:5278-5296
            real visc2cam = 4.0 * 2.0833 * A.len_disp * A.cam_coef;
            visc2cam = visc2cam * (1.0 - (real)(nl - (k + 1)) / (real)(A.n_cam_levels));
            kd = rmax(kd, visc2cam);
        }
        AT(D.kdiff, i, k) = kd;
    }
    real hd = 0.0;
    for (int e = 0; e < ne; e++) {
        const int iEdge = eoc[e];
        const real edge_sign = sgn[e] * D.dvEdge[iEdge];
        hd = hd + edge_sign * AT(D.ru, iEdge, k);
    }
    hd = hd * D.invAreaCell[i];
    AT(D.h_divergence, i, k) = hd;
    if (A.rk_step == 1) {
        AT(D.tend_rho, i, k) = -hd - D.rdzw[k] * (AT(D.rw, i, k + 1) - AT(D.rw, i, k)) + AT(D.tend_rho_physics, i, k);
        const real qt = AT(D.qtot, i, k);
        AT(D.dpdz, i, k) = -GRAVITY * (AT(D.rho_base, i, k) * (qt) + AT(D.rho_p_save, i, k) * (1. + qt));
    }
}

// vertical flux of u at interface kk (0-based kk = Fortran k-1), TI:5391-5402
__device__ __forceinline__ real wduz_at(const Dev& D, int kk, int iEdge, int cell1, int cell2, int LDK, int nl) {
    if (kk <= 0 || kk >= nl) return 0.;
    const real rwf = AT(D.rw, cell1, kk) + AT(D.rw, cell2, kk);
    RP u = D.u_2;
    if (kk == 1 || kk == nl - 1)
        return 0.5 * (rwf) * (D.fzm[kk] * AT(u, iEdge, kk) + D.fzp[kk] * AT(u, iEdge, kk - 1));
    return flux3(AT(u, iEdge, kk - 2), AT(u, iEdge, kk - 1), AT(u, iEdge, kk), AT(u, iEdge, kk + 1), 0.5 * (rwf), 1.0);
}

// (b) edge-all: [rk 1] PGF (5379-5387), delsq_u + del2 mixing (5467-5503); [owned edges] vertical transport,
//     nonlinear Coriolis, KE gradient (5391-5447); [rk > 1] final sum with tend_u_euler (5694-5701)
__global__ void k_dt_edge_b(const Dev D, const DynTendArgs A) {
    KI;
    if (i >= D.nEdges || k >= nl) return;
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    const bool solve = i < D.nEdgesSolve;
    const real invDc = D.invDcEdge[i];
    const real rho_e = AT(D.rho_edge, i, k);
    if (A.rk_step == 1) {
        real tue = 0.0;
        if (solve)
            tue = -AT(D.cqu, i, k) * ((AT(D.pressure_p, cell2, k) - AT(D.pressure_p, cell1, k)) * invDc / (.5 * (AT(D.zz, cell2, k) + AT(D.zz, cell1, k)))
                                      - 0.5 * AT(D.zxu, i, k) * (AT(D.dpdz, cell1, k) + AT(D.dpdz, cell2, k)));
        const int vertex1 = D.verticesOnEdge[2 * i], vertex2 = D.verticesOnEdge[2 * i + 1];
        const real r_dc = invDc;
        const real r_dv = rmin(D.invDvEdge[i], 4 * invDc);
        const real u_diffusion = (AT(D.divergence, cell2, k) - AT(D.divergence, cell1, k)) * r_dc
                                 - (AT(D.vorticity, vertex2, k) - AT(D.vorticity, vertex1, k)) * r_dv;
        AT(D.delsq_u, i, k) = 0.0 + u_diffusion;
        const real kdiffu = 0.5 * (AT(D.kdiff, cell1, k) + AT(D.kdiff, cell2, k));
        tue = tue + rho_e * kdiffu * u_diffusion * D.meshScalingDel2[i];
        AT(D.tend_u_euler, i, k) = tue;
    }
    if (!solve) return;
    const real w0 = wduz_at(D, k, i, cell1, cell2, LDK, nl);
    const real w1 = wduz_at(D, k + 1, i, cell1, cell2, LDK, nl);
    real tu = -D.rdzw[k] * (w1 - w0);
    real q = 0.0;
    const int neoe = D.nEdgesOnEdge[i];
    IP eoe_l = D.edgesOnEdge + (size_t)i * D.maxEdges2;
    RP woe = D.weightsOnEdge + (size_t)i * D.maxEdges2;
    const real pv_e = AT(D.pv_edge, i, k);
    for (int j = 0; j < neoe; j++) {
        const int eoe = eoe_l[j];
        const real workpv = 0.5 * (pv_e + AT(D.pv_edge, eoe, k));
        q = q + woe[j] * AT(D.u_2, eoe, k) * workpv;
    }
    const real uk = AT(D.u_2, i, k);
    tu = tu + rho_e * (q - (AT(D.ke, cell2, k) - AT(D.ke, cell1, k))
                           * invDc)
         - uk * 0.5 * (AT(D.h_divergence, cell1, k) + AT(D.h_divergence, cell2, k));
    if (A.rk_step != 1) {
        if (A.rayleigh_damp_u && k >= nl - A.n_rayleigh_levels)
            tu = tu - rho_e * uk * ((real)((k + 1) - (nl - A.n_rayleigh_levels)) * A.rayleigh_coef_inverse);
        tu = tu + AT(D.tend_u_euler, i, k) + AT(D.tend_ru_physics, i, k);
    }
    AT(D.tend_u, i, k) = tu;
}

// (c) rk 1: del^2 of delsq_u on vertices (5512-5529) and cells (5532-5551)
__global__ void k_dt_delsq_vertex(const Dev D) {
    KI;
    if (i >= D.nVertices || k >= nl) return;
    real acc = 0.0;
    const real iat = D.invAreaTriangle[i];
    for (int j = 0; j < 3; j++) {
        const int iEdge = D.edgesOnVertex[3 * i + j];
        const real edge_sign = iat * D.dcEdge[iEdge] * D.edgesOnVertex_sign[3 * i + j];
        acc = acc + edge_sign * AT(D.delsq_u, iEdge, k);
    }
    AT(D.delsq_vorticity, i, k) = acc;
}
__global__ void k_dt_delsq_cell(const Dev D) {
    KI;
    if (i >= D.nCells || k >= nl) return;
    real acc = 0.0;
    const real r = D.invAreaCell[i];
    const int ne = D.nEdgesOnCell[i];
    for (int e = 0; e < ne; e++) {
        const int iEdge = D.edgesOnCell[(size_t)i * D.maxEdges + e];
        const real edge_sign = r * D.dvEdge[iEdge] * D.edgesOnCell_sign[(size_t)i * D.maxEdges + e];
        acc = acc + edge_sign * AT(D.delsq_u, iEdge, k);
    }
    AT(D.delsq_divergence, i, k) = acc;
}

// (d) rk 1, owned edges: del^4 (5558-5584), vertical mixing (5592-5658), Rayleigh damping (5667-5690), final sum (5694-5701)
__global__ void k_dt_edge_d(const Dev D, const DynTendArgs A) {
    KI;
    if (i >= D.nEdgesSolve || k >= nl) return;
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    const real rho_e = AT(D.rho_edge, i, k);
    real tue = AT(D.tend_u_euler, i, k);
    if (A.h_mom_eddy_visc4 > 0.0) {
        const int vertex1 = D.verticesOnEdge[2 * i], vertex2 = D.verticesOnEdge[2 * i + 1];
        const real u_mix_scale = D.meshScalingDel4[i] * A.h_mom_eddy_visc4;
        const real r_dc = u_mix_scale * A.del4u_div_factor * D.invDcEdge[i];
        const real r_dv = u_mix_scale * rmin(D.invDvEdge[i], 4 * D.invDcEdge[i]);
        const real u_diffusion = rho_e * ((AT(D.delsq_divergence, cell2, k) - AT(D.delsq_divergence, cell1, k)) * r_dc
                                          - (AT(D.delsq_vorticity, vertex2, k) - AT(D.delsq_vorticity, vertex1, k)) * r_dv);
        tue = tue - u_diffusion;
    }
    if (A.v_mom_eddy_visc2 > 0.0 && k >= 1 && k < nl - 1) {
        real um[3];
        for (int dk = -1; dk <= 1; dk++) {
            real uu = AT(D.u_2, i, k + dk);
            if (!A.mix_full) uu = uu - D.u_init[k + dk] * cos(D.angleEdge[i]) - D.v_init[k + dk] * sin(D.angleEdge[i]);
            um[dk + 1] = uu;
        }
        const real z1 = 0.5 * (AT(D.zgrid, cell1, k - 1) + AT(D.zgrid, cell2, k - 1));
        const real z2 = 0.5 * (AT(D.zgrid, cell1, k) + AT(D.zgrid, cell2, k));
        const real z3 = 0.5 * (AT(D.zgrid, cell1, k + 1) + AT(D.zgrid, cell2, k + 1));
        const real z4 = 0.5 * (AT(D.zgrid, cell1, k + 2) + AT(D.zgrid, cell2, k + 2));
        const real zm = 0.5 * (z1 + z2), z0 = 0.5 * (z2 + z3), zp = 0.5 * (z3 + z4);
        tue = tue + rho_e * A.v_mom_eddy_visc2 * (
                        (um[2] - um[1]) / (zp - z0)
                        - (um[1] - um[0]) / (z0 - zm)) / (0.5 * (zp - zm));
    }
    AT(D.tend_u_euler, i, k) = tue;
    real tu = AT(D.tend_u, i, k);
    if (A.rayleigh_damp_u && k >= nl - A.n_rayleigh_levels)
        tu = tu - rho_e * AT(D.u_2, i, k) * ((real)((k + 1) - (nl - A.n_rayleigh_levels)) * A.rayleigh_coef_inverse);
    AT(D.tend_u, i, k) = tu + tue + AT(D.tend_ru_physics, i, k);
}

// (e) rk 1, cell-all: first del^2 of w (5795-5829) and of theta_m (6027-6057) with their 2nd-order mixing tendencies
__global__ void k_dt_cell_e(const Dev D, const DynTendArgs A) {
    KI;
    if (i >= D.nCells || k > nl) return;
    if (k == nl) { AT(D.tend_w_euler, i, k) = 0.0; return; }
    const int ne = D.nEdgesOnCell[i];
    IP eoc = D.edgesOnCell + (size_t)i * D.maxEdges;
    RP sgn = D.edgesOnCell_sign + (size_t)i * D.maxEdges;
    const real r_areaCell = D.invAreaCell[i];
    real dsw = 0.0, twe = 0.0, dst = 0.0, tte = 0.0;
    for (int e = 0; e < ne; e++) {
        const int iEdge = eoc[e];
        const int cell1 = D.cellsOnEdge[2 * iEdge], cell2 = D.cellsOnEdge[2 * iEdge + 1];
        const real dv = D.dvEdge[iEdge], idc = D.invDcEdge[iEdge], msd2 = D.meshScalingDel2[iEdge];
        const real rho_e = AT(D.rho_edge, iEdge, k);
        const real kd1 = AT(D.kdiff, cell1, k), kd2 = AT(D.kdiff, cell2, k);
        if (k >= 1) {
            const real edge_sign = 0.5 * r_areaCell * sgn[e] * dv * idc;
            real w_turb_flux = edge_sign * (rho_e + AT(D.rho_edge, iEdge, k - 1)) * (AT(D.w_2, cell2, k) - AT(D.w_2, cell1, k));
            dsw = dsw + w_turb_flux;
            w_turb_flux = w_turb_flux * msd2 * 0.25 *
                          (kd1 + kd2 + AT(D.kdiff, cell1, k - 1) + AT(D.kdiff, cell2, k - 1));
            twe = twe + w_turb_flux;
        }
        {
            const real edge_sign = r_areaCell * sgn[e] * dv * idc;
            const real pr_scale = A.prandtl_inv * msd2;
            real theta_turb_flux = edge_sign * (AT(D.theta_m_2, cell2, k) - AT(D.theta_m_2, cell1, k)) * rho_e;
            dst = dst + theta_turb_flux;
            theta_turb_flux = theta_turb_flux * 0.5 * (kd1 + kd2) * pr_scale;
            tte = tte + theta_turb_flux;
        }
    }
    AT(D.delsq_w, i, k) = dsw;
    AT(D.tend_w_euler, i, k) = twe;
    AT(D.delsq_theta, i, k) = dst;
    AT(D.tend_theta_euler, i, k) = tte;
}

// vertical flux of w (TI:5878-5891) and of theta_m (TI:6101-6116) at Fortran index kk+1
__device__ __forceinline__ real wdwz_at(const Dev& D, int kk, int iCell, int LDK, int nl) {
    if (kk <= 0 || kk >= nl) return 0.0;
    RP w = D.w_2; RP rw = D.rw;
    if (kk == 1 || kk == nl - 1)
        return 0.25 * (AT(rw, iCell, kk) + AT(rw, iCell, kk - 1)) * (AT(w, iCell, kk) + AT(w, iCell, kk - 1));
    return flux3(AT(w, iCell, kk - 2), AT(w, iCell, kk - 1), AT(w, iCell, kk), AT(w, iCell, kk + 1),
                 0.5 * (AT(rw, iCell, kk) + AT(rw, iCell, kk - 1)), 1.0);
}
__device__ __forceinline__ real wdtz_at(const Dev& D, int kk, int iCell, real coef3, int LDK, int nl) {
    if (kk <= 0 || kk >= nl) return 0.0;
    RP t = D.theta_m_2; RP ts = D.theta_m; RP rw = D.rw; RP rws = D.rw_save;
    const real fm = D.fzm[kk], fp = D.fzp[kk];
    if (kk == nl - 1)
        return AT(rws, iCell, kk) * (fm * AT(t, iCell, kk) + fp * AT(t, iCell, kk - 1));
    real f;
    if (kk == 1) f = AT(rw, iCell, kk) * (fm * AT(t, iCell, kk) + fp * AT(t, iCell, kk - 1));
    else f = flux3(AT(t, iCell, kk - 2), AT(t, iCell, kk - 1), AT(t, iCell, kk), AT(t, iCell, kk + 1), AT(rw, iCell, kk), coef3);
    return f + (AT(rws, iCell, kk) - AT(rw, iCell, kk)) * (fm * AT(ts, iCell, kk) + fp * AT(ts, iCell, kk - 1));
}

// (f) owned cells: tend_w (5713-5757, 5838-5945) and tend_theta (5956-6016, 6066-6126, 6134-6197)
__global__ void k_dt_cell_f(const Dev D, const DynTendArgs A) {
    KI;
    if (i >= D.nCellsSolve || k > nl) return;
    if (k == nl) { AT(D.tend_w, i, k) = 0.0; return; }
    const int ne = D.nEdgesOnCell[i];
    IP eoc = D.edgesOnCell + (size_t)i * D.maxEdges;
    RP sgn = D.edgesOnCell_sign + (size_t)i * D.maxEdges;
    const real invArea = D.invAreaCell[i];
    const real fm = D.fzm[k], fp = D.fzp[k];
    real tw = 0.0, tt = 0.0;
    for (int e = 0; e < ne; e++) {
        const int iEdge = eoc[e];
        const int nadv = D.nAdvCellsForEdge[iEdge];
        IP adv = D.advCellsForEdge + (size_t)iEdge * 15;
        RP ac = D.adv_coefs + (size_t)iEdge * 15;
        RP ac3 = D.adv_coefs_3rd + (size_t)iEdge * 15;
        const real ruk = AT(D.ru, iEdge, k);
        real ru_edge_w = 0.0, sw = 0.0;
        if (k >= 1) { ru_edge_w = fm * ruk + fp * AT(D.ru, iEdge, k - 1); sw = sign1(ru_edge_w); }
        const real st = sign1(ruk);
        real fw = 0.0, ft = 0.0;
        for (int j = 0; j < nadv; j++) {
            const int c = adv[j];
            if (k >= 1) { const real scalar_weight = ac[j] + sw * ac3[j]; fw = fw + scalar_weight * AT(D.w_2, c, k); }
            const real scalar_weight = ac[j] + st * ac3[j];
            ft = ft + scalar_weight * AT(D.theta_m_2, c, k);
        }
        if (k >= 1) tw = tw - sgn[e] * ru_edge_w * fw;
        tt = tt - sgn[e] * ruk * ft;
    }
    if (A.rk_step > 1) {          // perturbation flux for the rtheta_pp equation, TI:5995-6016
        for (int e = 0; e < ne; e++) {
            const int iEdge = eoc[e];
            const int cell1 = D.cellsOnEdge[2 * iEdge], cell2 = D.cellsOnEdge[2 * iEdge + 1];
            const real flux = sgn[e] * D.dvEdge[iEdge] * (AT(D.ru_save, iEdge, k) - AT(D.ru, iEdge, k)) * 0.5 * (AT(D.theta_m, cell2, k) + AT(D.theta_m, cell1, k));
            tt = tt - flux;
        }
    }
    real twe = 0.0, tte = 0.0;
    if (A.rk_step == 1) {
        twe = AT(D.tend_w_euler, i, k);
        tte = AT(D.tend_theta_euler, i, k);
        if (A.h_mom_eddy_visc4 > 0.0 && k >= 1) {
            const real r_areaCell = A.h_mom_eddy_visc4 * invArea;
            for (int e = 0; e < ne; e++) {
                const int iEdge = eoc[e];
                const int cell1 = D.cellsOnEdge[2 * iEdge], cell2 = D.cellsOnEdge[2 * iEdge + 1];
                const real edge_sign = D.meshScalingDel4[iEdge] * r_areaCell * D.dvEdge[iEdge] * sgn[e] * D.invDcEdge[iEdge];
                twe = twe - edge_sign * (AT(D.delsq_w, cell2, k) - AT(D.delsq_w, cell1, k));
            }
        }
        if (A.h_theta_eddy_visc4 > 0.0) {
            const real r_areaCell = A.h_theta_eddy_visc4 * A.prandtl_inv * invArea;
            for (int e = 0; e < ne; e++) {
                const int iEdge = eoc[e];
                const int cell1 = D.cellsOnEdge[2 * iEdge], cell2 = D.cellsOnEdge[2 * iEdge + 1];
                const real edge_sign = D.meshScalingDel4[iEdge] * r_areaCell * D.dvEdge[iEdge] * sgn[e] * D.invDcEdge[iEdge];
                tte = tte - edge_sign * (AT(D.delsq_theta, cell2, k) - AT(D.delsq_theta, cell1, k));
            }
        }
    }
    // ---- w: vertical advection, pressure gradient, buoyancy
    if (k >= 1) {
        const real f0 = wdwz_at(D, k, i, LDK, nl), f1 = wdwz_at(D, k + 1, i, LDK, nl);
        tw = tw * invArea - D.rdzu[k] * (f1 - f0);
        if (A.rk_step == 1) {
            twe = twe - AT(D.cqw, i, k) * (
                      D.rdzu[k] * (AT(D.pressure_p, i, k) - AT(D.pressure_p, i, k - 1))
                      - (fm * AT(D.dpdz, i, k) + fp * AT(D.dpdz, i, k - 1)));
            if (A.v_mom_eddy_visc2 > 0.0)
                twe = twe + A.v_mom_eddy_visc2 * 0.5 * (AT(D.rho_zz_2, i, k) + AT(D.rho_zz_2, i, k - 1)) * (
                          (AT(D.w_2, i, k + 1) - AT(D.w_2, i, k)) * D.rdzw[k]
                          - (AT(D.w_2, i, k) - AT(D.w_2, i, k - 1)) * D.rdzw[k - 1]) * D.rdzu[k];
            AT(D.tend_w_euler, i, k) = twe;
        } else {
            twe = AT(D.tend_w_euler, i, k);
        }
        AT(D.tend_w, i, k) = tw + twe;
    } else {
        AT(D.tend_w, i, k) = 0.0;
    }
    // ---- theta_m: vertical advection, mixing
    {
        const real f0 = wdtz_at(D, k, i, A.coef_3rd_order, LDK, nl), f1 = wdtz_at(D, k + 1, i, A.coef_3rd_order, LDK, nl);
        const real rho = AT(D.rho_zz_2, i, k);
        tt = tt * invArea - D.rdzw[k] * (f1 - f0);
        AT(D.rthdynten, i, k) = (tt - AT(D.tend_rho, i, k) * AT(D.theta_m_2, i, k)) / rho;
        tt = tt + rho * AT(D.rt_diabatic_tend, i, k);
        if (A.rk_step == 1) {
            if (A.v_theta_eddy_visc2 > 0.0 && k >= 1 && k < nl - 1) {
                const real z1 = AT(D.zgrid, i, k - 1), z2 = AT(D.zgrid, i, k), z3 = AT(D.zgrid, i, k + 1), z4 = AT(D.zgrid, i, k + 2);
                const real zm = 0.5 * (z1 + z2), z0 = 0.5 * (z2 + z3), zp = 0.5 * (z3 + z4);
                real tp = AT(D.theta_m_2, i, k + 1), t0 = AT(D.theta_m_2, i, k), tmm = AT(D.theta_m_2, i, k - 1);
                if (!A.mix_full) { tp = tp - AT(D.t_init, i, k + 1); t0 = t0 - AT(D.t_init, i, k); tmm = tmm - AT(D.t_init, i, k - 1); }
                tte = tte + A.v_theta_eddy_visc2 * A.prandtl_inv * rho * (
                          (tp - t0) / (zp - z0)
                          - (t0 - tmm) / (z0 - zm)) / (0.5 * (zp - zm));
            }
            AT(D.tend_theta_euler, i, k) = tte;
        } else {
            tte = AT(D.tend_theta_euler, i, k);
        }
        AT(D.tend_theta, i, k) = tt + tte + AT(D.tend_rtheta_physics, i, k);
    }
}

// ------------------------------------------------------------------ atm_set_smlstep_pert_variables_work  TI:2427-2508
__global__ void k_smlstep_pert(const Dev D) {
    KI;
    if (i >= D.nCellsSolve || k < 1 || k >= nl) return;
    if (D.bdyMaskCell[i] > 5) return;              // no conversion in the specified zone of a regional run, TI:2482
    const int ne = D.nEdgesOnCell[i];
    const real fm = D.fzm[k], fp = D.fzp[k];
    real wt = AT(D.tend_w, i, k);
    if (D.zb_any[i]) for (int e = 0; e < ne; e++) {
        const int iEdge = D.edgesOnCell[(size_t)i * D.maxEdges + e];
        const real tuk = AT(D.tend_u, iEdge, k);
        const real flux = D.edgesOnCell_sign[(size_t)i * D.maxEdges + e] * (fm * tuk + fp * AT(D.tend_u, iEdge, k - 1));
        const size_t zi = ((size_t)i * D.maxEdges + e) * LDK + k;
        wt = wt - (D.zb_cell[zi] + sign1(tuk) * D.zb3_cell[zi]) * flux;
    }
    AT(D.tend_w, i, k) = (fm * AT(D.zz, i, k) + fp * AT(D.zz, i, k - 1)) * wt;
}
