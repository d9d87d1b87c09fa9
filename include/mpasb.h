/* mpasb.h -- C ABI of the B200-native MPAS-Atmosphere dycore step.
 *
 * Drop-in boundary: the reference's time-integration driver
 *     call atm_timestep(domain, dt, currTime, itimestep, exchange_halo_group)
 *         src/core_atmosphere/mpas_atm_core.F:1022
 *         src/core_atmosphere/dynamics/mpas_atm_time_integration.F:739-800 (atm_timestep)
 *         ... :803-1725 (atm_srk3)
 * keeps its name and signature in the Fortran host; its body becomes
 * mpasb_set_field (first call / after a restart read) -> mpasb_step ->
 * mpasb_get_field (on output/restart alarms).  INTEGRATION.md shows the
 * ISO_C_BINDING interface block that binds these symbols.
 *
 * Conventions (identical to what a Fortran caller holds in its pools):
 *   - real arrays are RKIND (double in this build), dense, Fortran order,
 *     outer extent n+1 (the trailing "garbage" element, mpas_block_creator.F:1050-1121)
 *   - integer arrays are default 4-byte integers, connectivity is 1-BASED with
 *     out-of-block references equal to n+1 (mpas_block_creator.F:1464-1531)
 *   - every function returns 0 on success, nonzero on error (the host maps it to
 *     MPAS_LOG_CRIT, src/framework/mpas_log.F:627-629); mpasb_last_error() gives text.
 *   - one host thread per handle; one handle per mesh block per GPU.
 */
#ifndef MPASB_H
#define MPASB_H

#ifdef __cplusplus
extern "C" {
#endif

#ifdef MPASB_SINGLE
typedef float mpasb_real;       /* RKIND, PRECISION=single build (libmpasb_sp.so) */
#else
typedef double mpasb_real;      /* RKIND, PRECISION=double build (libmpasb.so) */
#endif

/* Block dimensions: mesh pool dims + block%dimensions (mpas_atm_time_integration.F:921-962). */
typedef struct mpasb_dims {
    int nCells, nEdges, nVertices;                 /* owned + halo, without the garbage slot */
    int nCellsSolve, nEdgesSolve, nVerticesSolve;  /* owned prefix */
    int nVertLevels, maxEdges, maxEdges2, vertexDegree;
    int num_scalars, index_qv, moist_start, moist_end;   /* 1-based scalar indices */
} mpasb_dims;

/* Namelist options the hot path reads (defaults Registry.xml:63-395)
 * plus the three mesh scalars cf1..cf3 and sphere_radius. */
typedef struct mpasb_config {
    int config_time_integration_order;      /* 2 | 3,          TI:891  */
    int config_number_of_sub_steps;         /*                 TI:890  */
    int config_dynamics_split_steps;        /*                 TI:898  */
    int config_split_dynamics_transport;    /* logical         TI:897  */
    int config_scalar_advection;            /* logical         TI:892  */
    int config_monotonic;                   /* logical         TI:894  */
    int config_positive_definite;           /* logical         TI:893  */
    int config_horiz_mixing;                /* 0 = "2d_smagorinsky", 1 = "2d_fixed"  TI:5226,5261 */
    int config_mix_full;                    /* logical         TI:5594 */
    int config_rayleigh_damp_u;             /* logical         TI:5667 */
    int config_number_rayleigh_damp_u_levels;
    int config_number_cam_damping_levels;
    int config_apply_lbcs;                  /* logical: regional run with lateral boundary conditions, TI:773 */
    int config_print_global_minmax_vel;     /* logical         TI:7960 */
    double config_epssm, config_smdiv, config_len_disp, config_coef_3rd_order;
    double config_visc4_2dsmag, config_smagorinsky_coef, config_del4u_div_factor;
    double config_h_mom_eddy_visc2, config_h_mom_eddy_visc4, config_v_mom_eddy_visc2;
    double config_h_theta_eddy_visc2, config_h_theta_eddy_visc4, config_v_theta_eddy_visc2;
    double config_apvm_upwinding, config_mpas_cam_coef, config_rayleigh_damp_u_timescale_days;
    double config_relax_zone_divdamp_coef;  /* Registry.xml:254, TI:7346-7580 */
    double cf1, cf2, cf3, sphere_radius;
    int on_a_sphere;                        /* mesh attribute (logical), mpas_vector_reconstruction.F:250 */
} mpasb_config;

typedef struct mpasb_handle_s* mpasb_handle;

/* mpas_atm_dynamics_init (TI:205) / mpas_atm_dynamics_finalize (TI:479) */
int  mpasb_create(const mpasb_dims* dims, const mpasb_config* cfg, int device, mpasb_handle* out);
int  mpasb_destroy(mpasb_handle h);
const char* mpasb_last_error(mpasb_handle h);

/* Field import/export by pool key (include/mpasb_fields.def lists every key).
 * time_level is 1 or 2 for state fields, 1 otherwise.  count = number of
 * elements of the dense host array (checked).  mpas_pool_get_array equivalents. */
int  mpasb_set_field(mpasb_handle h, const char* name, int time_level, const mpasb_real* src, long count);
int  mpasb_get_field(mpasb_handle h, const char* name, int time_level, mpasb_real* dst, long count);
int  mpasb_set_field_int(mpasb_handle h, const char* name, const int* src, long count);
int  mpasb_field_count(mpasb_handle h, const char* name, long* count);   /* dense host element count */
/* The same for a whole batch of real fields without stalling the host: all host->device (device->host) copies of a batch go
 * back to back on a copy stream of their own, ordered against the compute stream by events, so that the upload of the next
 * request, the step, and the download of the previous request overlap (PCIe is full duplex).  Host arrays should be pinned.
 * Upload: the host arrays may be reused after mpasb_wait_uploads; download: they are valid after mpasb_wait_downloads.
 * Device-side order is call order: set_fields_async; step; get_fields_async moves one request through. */
int  mpasb_set_fields_async(mpasb_handle h, int n, const char* const* names, const int* time_levels, const mpasb_real* const* src, const long* counts);
int  mpasb_get_fields_async(mpasb_handle h, int n, const char* const* names, const int* time_levels, mpasb_real* const* dst, const long* counts);
int  mpasb_wait_uploads(mpasb_handle h);
int  mpasb_wait_downloads(mpasb_handle h);
/* ... all but the last `lag` batches (0 <= lag < 4): lets a host keep `lag` requests in flight */
int  mpasb_wait_uploads_lag(mpasb_handle h, int lag);
int  mpasb_wait_downloads_lag(mpasb_handle h, int lag);

/* atm_srk3 (TI:803-1725): advance state level 1 -> level 2 by dt.  Followed by
 * mpasb_shift_time_levels == mpas_pool_shift_time_levels(state) (mpas_atm_core.F:808). */
int  mpasb_step(mpasb_handle h, mpasb_real dt, int itimestep);
/* The physics tendencies tend_ru_physics, tend_rtheta_physics, tend_rho_physics (filled by physics_get_tend inside atm_srk3 in
 * the reference, TI:1091-1093) are zero after mpasb_create and keep what mpasb_set_field uploads until this call. */
int  mpasb_zero_physics_tendencies(mpasb_handle h);
int  mpasb_shift_time_levels(mpasb_handle h);
/* Regional runs (config_apply_lbcs): seconds from the START of the next step to the end of the current LBC interval, i.e. what
 * mpas_atm_get_bdy_state derives from the clock (LBC_intv_end - currTime, mpas_atm_boundaries.F:497-503).  Call before every
 * mpasb_step; the lbc_* fields are uploaded with mpasb_set_field whenever the host reads a new LBC time. */
int  mpasb_set_lbc_time(mpasb_handle h, mpasb_real seconds_to_interval_end);
/* summarize_timestep (TI:7914-8357): out = {min w, max w, min u, max u} of level 2,
 * reductions start from 0 as in TI:8291-8292. */
int  mpasb_minmax(mpasb_handle h, mpasb_real out[4]);
/* The same without stalling the host: _async enqueues, behind the step, the reductions of w and u,
 * of every scalar (config_print_global_minmax_sca, TI:8322-8342) and the NaN tests of w and u (TI:8258-8281); _fetch
 * waits for their copy to the host.  minmax = {min w, max w, min u, max u, min s1, max s1, ...}, 4 <= n_minmax <=
 * 2*(2+num_scalars); nan_count = {NaNs in w, NaNs in u} (may be NULL).  A nonzero count is what the reference turns
 * into MPAS_LOG_CRIT "NaN detected in 'w' field." (TI:8268, 8280). */
int  mpasb_summarize_timestep_async(mpasb_handle h);
int  mpasb_summarize_timestep_fetch(mpasb_handle h, mpasb_real* minmax, long n_minmax, long nan_count[2]);
int  mpasb_synchronize(mpasb_handle h);

/* Init-time routines of the path: atm_init_coupled_diagnostics (TI:6776) and
 * atm_compute_solve_diagnostics without rk_step (mpas_atm_core.F:515-527). */
int  mpasb_init_coupled_diagnostics(mpasb_handle h);
int  mpasb_init_solve_diagnostics(mpasb_handle h, mpasb_real dt);
int  mpasb_init_solve_diagnostics_async(mpasb_handle h, mpasb_real dt);   /* enqueued only: no host synchronisation */

/* The mesh part of atm_mpas_init_block (mpas_atm_core.F:368-602): atm_compute_signs (:1151-1238), the inverses (:456-470),
 * atm_adv_coef_compression (:1285-1430), atm_couple_coef_3rd_order (:1433-1452), atm_compute_mesh_scaling (:1091-1148) and
 * atm_compute_damping_coefs (:1241-1282), computed by the library (C++, host) from the raw mesh fields of an init file, so
 * that a run can start with nothing derived on the caller's side.
 * Inputs, by name, in the layout a Fortran caller holds (dense, 1-based connectivity, garbage slot n+1):
 *   int : nEdgesOnCell, edgesOnCell, cellsOnCell, verticesOnCell, cellsOnEdge, verticesOnEdge, edgesOnVertex, cellsOnVertex
 *   real: zb, zb3 (nVertLevels+1, 2, nEdges+1), deriv_two (15, 2, nEdges+1), dcEdge, dvEdge, areaCell, areaTriangle,
 *         meshDensity, zgrid
 * Derived: edgesOnVertex_sign, edgesOnCell_sign, zb_cell, zb3_cell, kiteForCell, invAreaCell, invDvEdge, invDcEdge,
 * invAreaTriangle, nAdvCellsForEdge, advCellsForEdge, adv_coefs, adv_coefs_3rd, meshScalingDel2, meshScalingDel4,
 * meshScalingRegionalCell, meshScalingRegionalEdge, dss (shapes: include/mpasb_fields.def).
 * Optional inputs xCell, yCell, zCell, xEdge, yEdge, zEdge (all six, mesh on a sphere): coeffs_reconstruct is derived as well --
 * mpas_initialize_vectors (src/operators/mpas_vector_operations.F:652-771) and mpas_init_reconstruct
 * (src/operators/mpas_vector_reconstruction.F:60-177; radial-basis-function fit of mpas_rbf_interpolation.F:1079-1145 solved by
 * elgs / mpas_legs :1670-1846), which mpas_atm_core.F:534-535 call at start-up.
 * mpasb_init_block stores them in the handle (as mpasb_set_field / mpasb_set_field_int would); mpasb_init_block_host is the
 * same computation without a handle or a device, written to the caller's arrays.  The three namelist values are the ones
 * only this routine reads (Registry.xml:177, 261, 266).  Returns 1 when a named input / output is missing or unknown. */
int  mpasb_init_block(mpasb_handle h, int config_h_ScaleWithMesh, double config_zd, double config_xnutr,
                      int n_in, const char* const* in_names, const void* const* in_arrays);
int  mpasb_init_block_host(const mpasb_dims* dims, const mpasb_config* cfg, int config_h_ScaleWithMesh, double config_zd, double config_xnutr,
                           int n_in, const char* const* in_names, const void* const* in_arrays,
                           int n_out, const char* const* out_names, void* const* out_arrays);

/* mpas_reconstruct (src/operators/mpas_vector_reconstruction.F:205-330): uReconstructX/Y/Z/Zonal/Meridional from u of
 * the given time level through the init-time field coeffs_reconstruct; mpasb_step already ends with the call of
 * TI:1606 (time level 2, owned cells), mpas_atm_core.F:543 is (1, 0) at start-up.
 * atm_compute_output_diagnostics (mpas_atm_core.F:901-950): theta, rho, pressure of the given time level. */
int  mpasb_reconstruct(mpasb_handle h, int time_level, int include_halos);
int  mpasb_compute_output_diagnostics(mpasb_handle h, int time_level);

/* Kernel-level entry points, one per *_work routine, operating on the
 * device-resident fields of the handle (parity tests drive one at a time). */
int  mpasb_k_rk_integration_setup(mpasb_handle h);                                   /* TI:1930 */
/* regional path (TI:7198-7910, 1343-1388): routine = "speczone_tend" | "relaxzone_tend" (a = time_dyn_step, b = dt) |
 * "reset_u_ru" (a = time_dyn_step) | "adjust_scalars" (a = dt, b = rk_timestep) | "zero_gradient_w" |
 * "reset_speczone_values" (a = dt) | "set_scalars" (a = dt) */
int  mpasb_k_lbc(mpasb_handle h, const char* routine, mpasb_real a, mpasb_real b);
int  mpasb_k_compute_moist_coefficients(mpasb_handle h);                             /* TI:2042 */
int  mpasb_k_compute_vert_imp_coefs(mpasb_handle h, mpasb_real dts);                 /* TI:2225 */
int  mpasb_k_compute_dyn_tend(mpasb_handle h, int rk_step, mpasb_real dt);           /* TI:4982 */
int  mpasb_k_set_smlstep_pert_variables(mpasb_handle h);                             /* TI:2427 */
int  mpasb_k_advance_acoustic_step(mpasb_handle h, mpasb_real dts, int small_step);  /* TI:2646 */
int  mpasb_k_divergence_damping_3d(mpasb_handle h, mpasb_real dts);                  /* TI:2987 */
int  mpasb_k_recover_large_step_variables(mpasb_handle h, mpasb_real dt, int ns, int rk_step); /* TI:3189 */
int  mpasb_k_compute_solve_diagnostics(mpasb_handle h, mpasb_real dt, int rk_step);  /* TI:6337 */
int  mpasb_k_rk_dynamics_substep_finish(mpasb_handle h, int dynamics_substep, int dynamics_split); /* TI:7013 */
int  mpasb_k_advance_scalars(mpasb_handle h, mpasb_real dt, int rk_step);            /* TI:3575 */
int  mpasb_k_advance_scalars_mono(mpasb_handle h, mpasb_real dt);                    /* TI:4012 */
/* ... and split at its two exchange points (TI:4155 scalars_old, TI:4568 scale) for hosts that
 * keep calling the reference's own exchange_halo_group between the pieces; s is 0-based */
int  mpasb_k_advance_scalars_mono_pre(mpasb_handle h, mpasb_real dt);                /* TI:4129-4143 */
int  mpasb_k_advance_scalars_mono_a(mpasb_handle h, mpasb_real dt, int s);           /* TI:4177-4553 */
int  mpasb_k_advance_scalars_mono_b(mpasb_handle h, mpasb_real dt, int s);           /* TI:4579-4715 */

/* Halo exchange (mpas_halo_exch_group_full_halo_exch, src/framework/mpas_halo.F:498-846).
 * Lists are the reference's per-field sendListSrc/recvListDst (1-based local
 * indices), grouped per neighbour rank and halo layer (mpas_halo_types.inc:22-32). */
int  mpasb_set_halo_lists(mpasb_handle h, int kind /*0 cell,1 edge,2 vertex*/, int n_neighbors,
                          const int* neighbor_rank, int n_layers,
                          const int* n_send /*[n_neighbors*n_layers]*/, const int* send_src,
                          const int* n_recv /*[n_neighbors*n_layers]*/, const int* recv_dst);
int  mpasb_comm_init(mpasb_handle h, int rank, int world_size, const void* nccl_unique_id /*128 bytes*/);
int  mpasb_get_nccl_unique_id(void* out128);
int  mpasb_exchange_halo_group(mpasb_handle h, const char* group_name);   /* HALOS:90-167 names */
/* the same, only enqueued: for requests queued back to back (mpasb_set_fields_async ... mpasb_get_fields_async) */
int  mpasb_exchange_halo_group_async(mpasb_handle h, const char* group_name);
/* Optional: exchanges by direct stores into the neighbours' mailboxes over NVLink (CUDA IPC) instead of NCCL send/recv.
 * After mpasb_set_halo_lists and mpasb_comm_init, on every rank of ONE node (2..9 ranks):
 *   n = mpasb_p2p_max_message(h)                      largest message of this rank, in reals
 *   mpasb_p2p_prepare(h, max over ranks of n, out)    allocates mailbox + flags, writes two 64-byte IPC handles
 *   mpasb_p2p_open(h, handles of all ranks)           world x 128 bytes in rank order (host all-gather)
 *   mpasb_p2p_enable(h, 1)                            once EVERY rank succeeded so far (otherwise all stay on NCCL)
 * From then on every exchange of the handle is one put kernel and one get kernel. */
long mpasb_p2p_max_message(mpasb_handle h);
int  mpasb_p2p_prepare(mpasb_handle h, long slot_elems, void* out_handles128);
int  mpasb_p2p_open(mpasb_handle h, const void* all_handles);
int  mpasb_p2p_enable(mpasb_handle h, int on);

/* 1 (MPASB_STRICT=1 in the environment): handles created from now on keep the reference's operation order everywhere, nothing
 * contracted into FMAs: results bit-identical to the fp64 CPU arithmetic.  0 (default): the relaxed path (re-associated
 * stencil sums, explicit fma, scan-based column solve) within the parity bars (rel-L2 <= 1e-11 after one step). */
int  mpasb_strict_arithmetic(void);
int  mpasb_real_bytes(void);     /* sizeof(mpasb_real) of this build: 8 (libmpasb.so) or 4 (libmpasb_sp.so) */

/* Instrumentation */
long mpasb_kernel_launch_count(mpasb_handle h);      /* kernels launched by this handle so far */
int  mpasb_timer_start(mpasb_handle h);              /* CUDA event on the compute stream */
int  mpasb_timer_stop(mpasb_handle h, double* ms);   /* second event + elapsed time between the two */
int  mpasb_set_profile(mpasb_handle h, int on);      /* per-routine and per-kernel CUDA-event timing */
int  mpasb_get_profile(mpasb_handle h, char* buf, long buflen);  /* "name ms count\n" lines */

#ifdef __cplusplus
}
#endif
#endif /* MPASB_H */
