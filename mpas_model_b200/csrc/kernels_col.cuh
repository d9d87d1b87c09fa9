// kernels_col.cuh -- "column-warp" versions of the hot kernels of the dycore step.
//
// Mapping: one warp per mesh column (cell or edge), lane l owns the level pair
// (2l, 2l+1), so every field access is a 16-byte vector load of a level-contiguous
// column (LDK <= 64 levels; taller columns use the generic (k, column) kernels of
// kernels_dyn.cuh / kernels_acoustic.cuh).  Vertical neighbours (k-2 .. k+1) come from
// warp shuffles, the per-column connectivity and weights are loaded once per warp
// (one entry per lane) and broadcast, advection lists are staged in shared memory, and
// the column tridiagonal solve runs out of shared memory.
//
// Every expression keeps the operand order of the generic kernels (and therefore of
// mpas_atm_time_integration.F, "TI"), so results stay bit-identical to the fp64 CPU
// arithmetic when built --fmad=false.
//
// Latency structure: these kernels are chains of gathers whose addresses come from other loads, so
// what bounds them is the number of EXPOSED memory latencies per warp, not bytes.  Each kernel
// therefore (1) loads its connectivity once, with lanes beyond the list length duplicating the last
// valid entry so that every later gather is unconditional, (2) runs its gather loops fully unrolled
// over a fixed trip count (CW_NE edges, CW_NADV stencil cells; longer lists take a tail loop) with
// the accumulation -- not the load -- predicated, so all loads of a loop are in flight together,
// and (3) issues every store at the very end: a store in the middle would pin all later loads
// behind it (the compiler must assume the Dev pointers alias).
#pragma once
#include "kernels_dyn.cuh"

#define CW_FULL 0xffffffffu
#ifndef CW_WARPS
#define CW_WARPS 8                      // warps (= columns) per block
#endif
#define CW_THREADS (CW_WARPS * 32)
#ifndef EB_WARPS
#define EB_WARPS CW_WARPS               // ... of the edge tendency kernel
#endif

// resident blocks per SM each kernel's register allocation aims for (1 = no constraint); tuned on B200, DESIGN.md §4
#ifndef MB_EDGE_B
#define MB_EDGE_B 6
#endif
#ifndef MB_SML
#define MB_SML 5
#endif
#ifndef MB_REC2
#define MB_REC2 6
#endif
#ifndef MB_DIAG_E
#define MB_DIAG_E 0
#endif
#ifndef MB_REC1
#define MB_REC1 4
#endif
#ifndef MB_DIAG_C
#define MB_DIAG_C 5
#endif
#ifndef MB_CELL_A
#define MB_CELL_A 6
#endif
#ifndef MB_CELL_E
#define MB_CELL_E 3
#endif
#ifndef MB_SC_CELL
#define MB_SC_CELL 4
#endif

struct __align__(2 * sizeof(real)) r2 { real x, y; };
struct b2 { bool x, y; };

__device__ __forceinline__ r2 mk2(real a, real b) { r2 r; r.x = a; r.y = b; return r; }
__device__ __forceinline__ r2 splat(real a) { return mk2(a, a); }
__device__ __forceinline__ r2 operator+(r2 a, r2 b) { return mk2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ r2 operator-(r2 a, r2 b) { return mk2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ r2 operator*(r2 a, r2 b) { return mk2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ r2 operator/(r2 a, r2 b) { return mk2(a.x / b.x, a.y / b.y); }
__device__ __forceinline__ r2 operator+(r2 a, real b) { return mk2(a.x + b, a.y + b); }
__device__ __forceinline__ r2 operator-(r2 a, real b) { return mk2(a.x - b, a.y - b); }
__device__ __forceinline__ r2 operator*(r2 a, real b) { return mk2(a.x * b, a.y * b); }
__device__ __forceinline__ r2 operator/(r2 a, real b) { return mk2(a.x / b, a.y / b); }
__device__ __forceinline__ r2 operator+(real a, r2 b) { return mk2(a + b.x, a + b.y); }
__device__ __forceinline__ r2 operator-(real a, r2 b) { return mk2(a - b.x, a - b.y); }
__device__ __forceinline__ r2 operator*(real a, r2 b) { return mk2(a * b.x, a * b.y); }
__device__ __forceinline__ r2 operator/(real a, r2 b) { return mk2(a / b.x, a / b.y); }
__device__ __forceinline__ r2 operator-(r2 a) { return mk2(-a.x, -a.y); }
__device__ __forceinline__ r2 sel(b2 m, r2 a, r2 b) { return mk2(m.x ? a.x : b.x, m.y ? a.y : b.y); }
__device__ __forceinline__ r2 sel(b2 m, r2 a, real b) { return mk2(m.x ? a.x : b, m.y ? a.y : b); }
__device__ __forceinline__ r2 selb(bool on, r2 a, r2 b) { return mk2(on ? a.x : b.x, on ? a.y : b.y); }   // warp-uniform condition
__device__ __forceinline__ r2 abs2(r2 a) { return mk2(fabs(a.x), fabs(a.y)); }
__device__ __forceinline__ b2 operator&&(b2 a, b2 b) { b2 r; r.x = a.x && b.x; r.y = a.y && b.y; return r; }
__device__ __forceinline__ b2 operator||(b2 a, b2 b) { b2 r; r.x = a.x || b.x; r.y = a.y || b.y; return r; }
__device__ __forceinline__ b2 operator!(b2 a) { b2 r; r.x = !a.x; r.y = !a.y; return r; }
// Fortran sign(1.0, x) > 0  <=>  sign bit clear (so +0 -> +1, -0 -> -1, TI:5744, 5978)
__device__ __forceinline__ b2 nonneg_sign(r2 a) { b2 r; r.x = !signbit(a.x); r.y = !signbit(a.y); return r; }

// per-lane level predicates for the pair (k0, k0 + 1)
struct Lv {
    int k0;
    __device__ __forceinline__ b2 lt(int n) const { b2 r; r.x = k0 < n; r.y = k0 + 1 < n; return r; }
    __device__ __forceinline__ b2 ge(int n) const { b2 r; r.x = k0 >= n; r.y = k0 + 1 >= n; return r; }
    __device__ __forceinline__ b2 eq(int n) const { b2 r; r.x = k0 == n; r.y = k0 + 1 == n; return r; }
    __device__ __forceinline__ b2 in(int lo, int hi) const { b2 r; r.x = k0 >= lo && k0 <= hi; r.y = k0 + 1 >= lo && k0 + 1 <= hi; return r; }
};

// column pair load/store.  Element offsets are 32-bit (a block holds < 2^32 reals per field), so an
// address costs one IMAD + one IMAD.WIDE; lanes beyond the padded column re-read its last pair
// (kc = min(k0, LDK-2)) instead of being predicated off -- their results are never stored.
__device__ __forceinline__ r2 ld2(const real* __restrict__ p, unsigned off) {
    return *reinterpret_cast<const r2*>(p + off);
}
__device__ __forceinline__ void st2(real* p, unsigned off, bool act, r2 v) {
    if (act) *reinterpret_cast<r2*>(p + off) = v;
}
// vertical shifts inside the warp: result[k] = v[k-1], v[k+1], v[k-2]
__device__ __forceinline__ r2 up1(r2 v) { return mk2(__shfl_up_sync(CW_FULL, v.y, 1), v.x); }
__device__ __forceinline__ r2 dn1(r2 v) { return mk2(v.y, __shfl_down_sync(CW_FULL, v.x, 1)); }
__device__ __forceinline__ r2 up2(r2 v) { return mk2(__shfl_up_sync(CW_FULL, v.x, 1), __shfl_up_sync(CW_FULL, v.y, 1)); }
__device__ __forceinline__ r2 dn2(r2 v) { return mk2(__shfl_down_sync(CW_FULL, v.x, 1), __shfl_down_sync(CW_FULL, v.y, 1)); }

__device__ __forceinline__ r2 flux4_2(r2 q_im2, r2 q_im1, r2 q_i, r2 q_ip1, r2 ua) {
    return ua * (7. * (q_i + q_im1) - (q_ip1 + q_im2)) / 12.0;
}
__device__ __forceinline__ r2 flux3_2(r2 q_im2, r2 q_im1, r2 q_i, r2 q_ip1, r2 ua, real coef3) {
    return flux4_2(q_im2, q_im1, q_i, q_ip1, ua) + coef3 * abs2(ua) * ((q_ip1 - q_im2) - 3. * (q_i - q_im1)) / 12.0;
}

#define CW_SETUP(ncols) PDL_ENTER CW_SETUP_NW(ncols)
// the hot kernels: CW_ENTER, loads of static mesh data (connectivity, metric weights), then pdl_wait() before the first field access
#define CW_ENTER(ncols) pdl_trigger(); CW_SETUP_NW(ncols)
// Sweep direction.  Consecutive kernels of a step walk their columns in OPPOSITE directions, so a kernel starts where its
// predecessor just finished and finds the tail of that kernel's inputs and outputs still in the 126 MB L2.  The direction is a
// compile-time property of each kernel (the order of the kernels in a stage is fixed and alternates: see srk3): the _R forms
// start at the last column.  (A run-time direction flag cost every kernel a register -- a reversed index cannot be
// rematerialised from blockIdx for free -- and pushed three kernels over an occupancy step: measured)
#define CW_SETUP_R(ncols) PDL_ENTER CW_SETUP_NW_R(ncols)
#define CW_ENTER_R(ncols) pdl_trigger(); CW_SETUP_NW_R(ncols)
#define CW_SETUP_NW_R(ncols)                                                                  \
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;                                \
    const int i = (ncols) - 1 - (int)(blockIdx.x * (blockDim.x >> 5) + wib);                           \
    const int LDK = D.LDK, nl = D.nl;                                                         \
    if (i < 0) return;                                                                        \
    CW_SETUP_REST
// ... without the wait: the kernel places pdl_wait() itself, below its loads of static mesh data
#define CW_SETUP_NW(ncols)                                                                    \
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;                                \
    const int i = blockIdx.x * (blockDim.x >> 5) + wib;                                       \
    const int LDK = D.LDK, nl = D.nl;                                                         \
    if (i >= (ncols)) return;                                                                 \
    CW_SETUP_REST
#define CW_SETUP_REST                                                                         \
    Lv lv; lv.k0 = 2 * lane;                                                                  \
    const int k0 = lv.k0; const bool act = k0 < D.LDKA; (void)nl;                             \
    const unsigned uLDK = (unsigned)LDK, kc = (unsigned)min(k0, D.LDKA - 2);
#define LD(p, col) ld2((p), (unsigned)(col) * uLDK + kc)
#define ST(p, col, v) st2((p), (unsigned)(col) * uLDK + kc, act, (v))
#define BC(v, src) __shfl_sync(CW_FULL, (v), (src))
// request a column pair into L2 without holding registers: used for operands of a later, dependent phase
__device__ __forceinline__ void pf2(const real* p, unsigned off) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p + off)); }
#define PF(p, col) pf2((p), (unsigned)(col) * uLDK + kc)


// Persistent-warp schedules.  Default: warp g walks columns g, g + G, ... (the whole grid sweeps one window of G columns).
// MPASB_PERS_CHUNK: the R blocks b, b + NS, b + 2 NS, ... (NS = gridDim / R; the block scheduler deals consecutive blocks round
// robin over the SMs, so these R are the ones resident together on one SM) share ONE contiguous range of columns and walk it
// side by side, R*W consecutive columns per trip: the columns a SM gathers at any time are neighbours on the mesh (Morton
// order), so what one warp pulled into L1 the next ones find there.  Results do not depend on the schedule.
struct Pers { int j, end, stride; };
template <int W, int R>
__device__ __forceinline__ Pers pers_init(int n, int wib) {
    Pers p;
#ifdef MPASB_PERS_CHUNK
    const int nb = (int)gridDim.x;
    const bool grp = R > 1 && nb % R == 0;
    const int NS = grp ? nb / R : nb, r = grp ? (int)blockIdx.x / NS : 0, s = grp ? (int)blockIdx.x % NS : (int)blockIdx.x;
    const int L = (n + NS - 1) / NS;
    p.stride = (grp ? R : 1) * W;
    p.end = min(n, (s + 1) * L);
    p.j = s * L + r * W + wib;
#else
    p.j = (int)blockIdx.x * W + wib; p.end = n; p.stride = (int)gridDim.x * W;
#endif
    return p;
}

// ------------------------------------------------------------------ atm_compute_dyn_tend_work, part (f)
// The reference computes the 3rd/4th-order horizontal flux of w and theta_m inside the cell loop, i.e. every
// edge flux twice -- once from each adjacent cell (TI:5713-5757, 5956-5991) -- through a 10-cell stencil.  That
// loop is bound by the L1 data pipe (120 gathered columns per cell).  Here the flux is computed ONCE per edge
// (k2_dt_edge_flux: 20 gathered columns per edge = 60 per cell) into two edge scratch arrays, and the cell kernel
// only sums its <= CW_MAXNE edge fluxes.  The result is bit-identical: the reference's cell-side expression
// (sign * ru_edge) * flux and sign * (ru_edge * flux) differ by an exact sign flip only (sign = +/-1).
#define CW_MAXNE 8                      // most edges per cell these kernels handle (host checks nEdgesOnCell)
#define CW_NE 6                         // edges per cell covered by the unrolled loops
__global__ void __launch_bounds__(CW_THREADS) k2_dt_edge_flux(const Dev D) {
    CW_ENTER(D.nEdges)
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    pdl_wait();
    if (!(cell1 < D.nCellsSolve || cell2 < D.nCellsSolve)) return;           // only edges of owned cells are consumed
    const int nadv = D.nAdvCellsForEdge[i];
    // one stencil entry per lane: cell index and the two possible weights adv_coefs +/- adv_coefs_3rd
    // (TI:5744-5750: coef + sign(ru) * coef_3rd with sign = +/-1)
    int my_c = 0; real my_wp = 0.0, my_wm = 0.0;
    if (lane < nadv) {
        my_c = D.advCellsForEdge[(unsigned)i * 15 + lane];
        const real a = D.adv_coefs[(unsigned)i * 15 + lane], b = D.adv_coefs_3rd[(unsigned)i * 15 + lane];
        my_wp = a + b; my_wm = a - b;
    }
    const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0);
    const r2 ruk = LD(D.ru, i);
    const r2 ruw = fm * ruk + fp * up1(ruk);               // ru_edge_w (levels k >= 1)
    const b2 pw = nonneg_sign(ruw), pt = nonneg_sign(ruk);
    r2 fw = mk2(0.0, 0.0), ft = mk2(0.0, 0.0);
#pragma unroll 5
    for (int j = 0; j < nadv; j++) {
        const int c = BC(my_c, j);
        const real wp = BC(my_wp, j), wm = BC(my_wm, j);
        const r2 w2 = LD(D.w_2, c), t2 = LD(D.theta_m_2, c);
        fw.x = fw.x + (pw.x ? wp : wm) * w2.x;
        fw.y = fw.y + (pw.y ? wp : wm) * w2.y;
        ft.x = ft.x + (pt.x ? wp : wm) * t2.x;
        ft.y = ft.y + (pt.y ? wp : wm) * t2.y;
    }
    ST(D.adv_flux_w, i, sel(lv.ge(1) && lv.lt(nl), ruw * fw, 0.0));
    ST(D.adv_flux_theta, i, sel(lv.lt(nl), ruk * ft, 0.0));
}

// ---- the same per-edge flux with the stencil columns staged in shared memory by bulk copies (TMA) ----
// k2_dt_edge_flux gathers 20 columns of 448 B per edge through L1 (the L1 data pipe is what bounds it).  Here a block
// owns EF_EB consecutive edges; the UNION of their stencil cells (47 on average for 32 edges of a Morton-ordered
// icosahedral mesh, against 320 gathers) is precomputed at upload (flux_tiles, mpasb.cu) and each of its w / theta_m
// columns is brought into shared memory ONCE: the union is covered by RUNS of consecutive cell indices (cells are
// Morton-ordered, so 9 runs on average when gaps of <= EF_GAP unused cells are bridged: 14 % more columns than the
// union), one cp.async.bulk (UBLKCP) of ncols x 448 B per run and field, completion on an mbarrier.
// The 10-cell stencil sums then read shared memory (LDS.128, conflict free: lane l reads bytes [16 l, 16 l + 16) of a
// 448-byte row).  Tiles whose union exceeds EF_MAXT columns (about 1 %) gather from global memory as before.
// Accumulation order per edge is unchanged: bit-identical to k2_dt_edge_flux.
#define EF_EB 32                        // edges per block
#define EF_MAXT 80                      // staged columns per block and field
#define EF_MAXR 24                      // runs of consecutive cells per block
#define EF_GAP 2                        // unused cells bridged inside a run
#define EF_EPW (EF_EB / CW_WARPS)       // edges per warp
#ifndef EF_MINB
#define EF_MINB 3
#endif
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__global__ void __launch_bounds__(CW_THREADS, EF_MINB) k4_dt_edge_flux(const Dev D, const int4* __restrict__ tile_hdr,
                                                             const int4* __restrict__ tile_runs, const unsigned char* __restrict__ tile_slot) {
    extern __shared__ __align__(128) unsigned char ef_raw[];
    __shared__ __align__(8) unsigned long long ef_bar;
    PDL_ENTER
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int LDK = D.LDK, nl = D.nl;
    real* s_w = reinterpret_cast<real*>(ef_raw);
    real* s_t = s_w + EF_MAXT * LDK;
    const int tile = blockIdx.x;
    const int4 hdr = tile_hdr[tile];                        // (runs, staged columns, active-edge mask); runs < 0: gather from global memory
    const int nt = hdr.x;
    Lv lv; lv.k0 = 2 * lane;
    const int k0 = lv.k0; const bool act = k0 < D.LDKA;
    const unsigned uLDK = (unsigned)LDK, kc = (unsigned)min(k0, D.LDKA - 2);
    const unsigned bar = smem_u32(&ef_bar);
    if (nt > 0) {
        if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_expect_tx(bar, (unsigned)(hdr.y * 2 * LDK * sizeof(real))); }
        __syncthreads();
        if ((int)threadIdx.x < 2 * nt) {                    // thread 2r: w columns of run r, thread 2r + 1: theta_m columns
            const int4 run = tile_runs[tile * EF_MAXR + (threadIdx.x >> 1)];        // (first cell, columns, first slot)
            const unsigned bytes = (unsigned)(run.y * LDK * sizeof(real));
            if (threadIdx.x & 1) bulk_g2s(smem_u32(s_t + run.z * LDK), D.theta_m_2 + (size_t)run.x * LDK, bytes, bar);
            else bulk_g2s(smem_u32(s_w + run.z * LDK), D.w_2 + (size_t)run.x * LDK, bytes, bar);
        }
    }
    // per-edge metadata and the edge's own column while the copies are in flight: every load is addressed by the edge
    // index alone (the active mask and the stencil length come with the tile tables), so all of them are in flight together
    const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0);
    int e_nadv[EF_EPW], e_c[EF_EPW]; real e_wp[EF_EPW], e_wm[EF_EPW]; r2 e_ru[EF_EPW];
    const int l15 = min(lane, 14);
#pragma unroll
    for (int q = 0; q < EF_EPW; q++) {
        const int i = tile * EF_EB + q * CW_WARPS + wib;
        const unsigned ic = (unsigned)min(i, D.nEdges);    // the garbage row of the per-edge arrays
        const int sl = tile_slot[(unsigned)i * 15 + l15];
        const int cg = D.advCellsForEdge[ic * 15 + l15];
        const real a = D.adv_coefs[ic * 15 + l15], b = D.adv_coefs_3rd[ic * 15 + l15];
        e_ru[q] = LD(D.ru, ic);
        const bool valid = lane < 15 && sl != 0xff;
        e_nadv[q] = __popc(__ballot_sync(CW_FULL, valid));
        e_c[q] = nt > 0 ? sl : cg;
        e_wp[q] = a + b; e_wm[q] = a - b;
    }
    if (nt > 0) mbar_wait(bar, 0);
#pragma unroll
    for (int q = 0; q < EF_EPW; q++) {
        if (!((hdr.z >> (q * CW_WARPS + wib)) & 1)) continue;                      // warp-uniform: edge without an owned cell
        const int i = tile * EF_EB + q * CW_WARPS + wib;
        const r2 ruk = e_ru[q];
        const r2 ruw = fm * ruk + fp * up1(ruk);
        const b2 pw = nonneg_sign(ruw), pt = nonneg_sign(ruk);
        r2 fw = mk2(0.0, 0.0), ft = mk2(0.0, 0.0);
        const int nadv = e_nadv[q];
        if (nt > 0) {
#pragma unroll 5
            for (int j = 0; j < nadv; j++) {
                const int sl = BC(e_c[q], j);
                const real wp = BC(e_wp[q], j), wm = BC(e_wm[q], j);
                const r2 w2 = *reinterpret_cast<const r2*>(s_w + sl * LDK + kc), t2 = *reinterpret_cast<const r2*>(s_t + sl * LDK + kc);
                fw.x = fw.x + (pw.x ? wp : wm) * w2.x;
                fw.y = fw.y + (pw.y ? wp : wm) * w2.y;
                ft.x = ft.x + (pt.x ? wp : wm) * t2.x;
                ft.y = ft.y + (pt.y ? wp : wm) * t2.y;
            }
        } else {
#pragma unroll 5
            for (int j = 0; j < nadv; j++) {
                const int c = BC(e_c[q], j);
                const real wp = BC(e_wp[q], j), wm = BC(e_wm[q], j);
                const r2 w2 = LD(D.w_2, c), t2 = LD(D.theta_m_2, c);
                fw.x = fw.x + (pw.x ? wp : wm) * w2.x;
                fw.y = fw.y + (pw.y ? wp : wm) * w2.y;
                ft.x = ft.x + (pt.x ? wp : wm) * t2.x;
                ft.y = ft.y + (pt.y ? wp : wm) * t2.y;
            }
        }
        ST(D.adv_flux_w, i, sel(lv.ge(1) && lv.lt(nl), ruw * fw, 0.0));
        ST(D.adv_flux_theta, i, sel(lv.lt(nl), ruk * ft, 0.0));
    }
}

// ---- cell-centred flux sweep (relaxed arithmetic): horizontal flux divergence of w and theta_m, TI:5713-5757, 5956-5991 ----
// One warp per owned cell c.  The reference evaluates, for each of the 6 edges of c, flux_e = ru_e * sum_j (adv_coefs_j
// +/- adv_coefs_3rd_j) q(cell_j) over the 10-cell stencil of the edge; the stencils of the 6 edges of one cell cover only 19
// distinct cells (c, its ring n_0..n_5, the second ring m_0..m_11), and on a hexagonal neighbourhood WHICH of them an edge
// uses is known statically: {c, n_0..n_5, m_{2i-1}, m_{2i}, m_{2i+1}} for edge i (tables built and checked on the host,
// build_flux_rings in mpasb.cu).  So the 7 inner columns of w and theta_m are loaded once and stay in registers, the outer
// ring as well, and every stencil sum is register arithmetic: 38 gathered columns per cell
// instead of 60 (per-edge kernel) or 120 (reference loop nest), and no per-edge flux arrays (12 C less HBM traffic per call).
// Each edge flux is computed by both adjacent cells, as in the reference.  Arithmetic differs from the strict path only by
// association: F4 = sum a_j q_j and F3 = sum b_j q_j are accumulated separately with fma() and combined as F4 +/- F3.
// The 2 x 60 weights of the cell are staged in shared memory by the warp and read back as warp-uniform LDS.128 pairs.
// Irregular cells (pentagons, heptagons and their neighbours; ~0.2 % of an icosahedral mesh) walk advCellsForEdge.
#define FX_RING 20                      // ints per cell: n_0..n_5, m_0..m_11, regular flag, pad
#define FX_WTS 128                      // reals per cell: a[6][10] at 0, b[6][10] at 64
#define FX_WARPS 4
#ifndef FX_MINB
#define FX_MINB 2
#endif
__device__ __forceinline__ r2 fma2(real a, r2 q, r2 acc) { return mk2(fma(a, q.x, acc.x), fma(a, q.y, acc.y)); }
#ifndef FX_TMA
#define FX_TMA 1                        // 0: the weight row travels through registers (measured the same: 0.977 vs 0.966-0.982 ms/step)
#endif
// FX_TMA: the cell's weight row (FX_WTS reals, one contiguous 1 KB slab of D.fx_w) is not carried through registers but fetched
// by one bulk-asynchronous copy (cp.async.bulk global -> shared, completion on an mbarrier; SASS UBLKCP / SYNCS) issued by lane 0
// one cell AHEAD into the other half of a double buffer, so it costs the warp neither registers nor LSU issue slots.
__global__ void __launch_bounds__(FX_WARPS * 32, FX_MINB) k5_flux_cell(const Dev D) {
#if FX_TMA
    __shared__ __align__(128) real s_wts2[FX_WARPS][2][FX_WTS];
    __shared__ __align__(8) unsigned long long s_bar[FX_WARPS][2];
#else
    __shared__ __align__(16) real s_wts[FX_WARPS][FX_WTS];
#endif
    pdl_trigger();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int LDK = D.LDK, nl = D.nl;
    Lv lv; lv.k0 = 2 * lane;
    const int k0 = lv.k0; const bool act = k0 < D.LDKA;
    const unsigned uLDK = (unsigned)LDK, kc = (unsigned)min(k0, D.LDKA - 2);
    const b2 k_lt_nl = lv.lt(nl), k_mid = lv.ge(1) && lv.lt(nl);
    const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0);
    // Persistent warps: warp g handles cells g, g + G, g + 2G, ...  The kernel is latency bound (two dependent memory round
    // trips per cell: neighbourhood table -> columns), so the table row, the edge ids and the signs of the NEXT cell are
    // fetched while the current cell's columns are in flight: one exposed round trip per cell instead of two.
    const int le = min(lane, 5), lr = min(lane, FX_RING - 1);
    const Pers ps = pers_init<FX_WARPS, FX_MINB>(D.nCellsSolve, wib);
    const int G = ps.stride, nSolve = ps.end;
#define FX_COL(j) (j)                                         /* forward sweep (see CW_SETUP_R) */
    int j = ps.j;
    int my_ring = 0, my_e = 0; real my_sgn = 0.0;
#if FX_TMA
    const unsigned bar0 = smem_u32(&s_bar[wib][0]), buf0 = smem_u32(&s_wts2[wib][0][0]);
    unsigned trip = 0;                                   // buffer = trip & 1, mbarrier phase = (trip >> 1) & 1
    if (lane == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); }
    __syncwarp();
#endif
    if (j < nSolve) {
        const int i = FX_COL(j);
        my_ring = D.fx_ring[(unsigned)i * FX_RING + lr];
        my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
        my_sgn = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le];
#if FX_TMA
        if (lane == 0) { mbar_expect_tx(bar0, FX_WTS * sizeof(real)); bulk_g2s(buf0, D.fx_w + (size_t)i * FX_WTS, FX_WTS * sizeof(real), bar0); }
#endif
    }
    pdl_wait();                                          // everything above is static mesh data
    for (; j < nSolve; j += G) {
    const int i = FX_COL(j);
    const int regular = BC(my_ring, 18);
    r2 tw = mk2(0.0, 0.0), tt = mk2(0.0, 0.0);
    const int inext = j + G < nSolve ? FX_COL(j + G) : D.nCellsSolve;
    int nx_ring = 0, nx_e = 0; real nx_sgn = 0.0;
#if FX_TMA
    const unsigned cur = trip & 1u, ph = (trip >> 1) & 1u;
    trip++;
    __syncwarp();                                        // every lane is done with the other buffer (read two trips ago ... one trip ago)
    if (inext < D.nCellsSolve && lane == 0) {            // next cell's weights into the other buffer
        const unsigned nb = bar0 + 8u * (cur ^ 1u);
        mbar_expect_tx(nb, FX_WTS * sizeof(real));
        bulk_g2s(buf0 + (cur ^ 1u) * (unsigned)(FX_WTS * sizeof(real)), D.fx_w + (size_t)inext * FX_WTS, FX_WTS * sizeof(real), nb);
    }
#endif
    if (regular) {
#if !FX_TMA
        const real* __restrict__ wsrc = D.fx_w + (size_t)i * FX_WTS + 4 * lane;
        const r2 wts_a = *reinterpret_cast<const r2*>(wsrc), wts_b = *reinterpret_cast<const r2*>(wsrc + 2);
#endif
        // all 19 columns of both fields and the 6 edge columns are requested before any of them is used
        const r2 wc = LD(D.w_2, i), tc = LD(D.theta_m_2, i);
        r2 wn[6], tn[6], wm[12], tm[12], ru6[6];
#pragma unroll
        for (int q = 0; q < 6; q++) { const int c = BC(my_ring, q); wn[q] = LD(D.w_2, c); tn[q] = LD(D.theta_m_2, c); }
#pragma unroll
        for (int q = 0; q < 12; q++) { const int c = BC(my_ring, 6 + q); wm[q] = LD(D.w_2, c); tm[q] = LD(D.theta_m_2, c); }
#pragma unroll
        for (int q = 0; q < 6; q++) ru6[q] = LD(D.ru, BC(my_e, q));
        if (inext < D.nCellsSolve) {                    // next cell's table row, edges, signs: in flight with the columns
            nx_ring = D.fx_ring[(unsigned)inext * FX_RING + lr];
            nx_e = D.edgesOnCell[(unsigned)inext * D.maxEdges + le];
            nx_sgn = D.edgesOnCell_sign[(unsigned)inext * D.maxEdges + le];
        }
#if FX_TMA
        mbar_wait(bar0 + 8u * cur, ph);                  // this cell's weights have landed (requested one trip ago)
        const real* __restrict__ W = s_wts2[wib][cur];
#else
        __syncwarp();                                    // the previous cell's reads of s_wts are done
        *reinterpret_cast<r2*>(&s_wts[wib][4 * lane]) = wts_a; *reinterpret_cast<r2*>(&s_wts[wib][4 * lane + 2]) = wts_b;
        __syncwarp();
        const real* __restrict__ W = s_wts[wib];
#endif
#pragma unroll
        for (int e = 0; e < 6; e++) {
            const r2 wa = wm[(2 * e + 11) % 12], ta = tm[(2 * e + 11) % 12], wb = wm[2 * e], tb = tm[2 * e],
                     wd = wm[(2 * e + 1) % 12], td = tm[(2 * e + 1) % 12];
            const r2 ruk = ru6[e];
            const real sg = BC(my_sgn, e);
            const real* __restrict__ A = W + e * 10;
            const real* __restrict__ B = W + FX_WTS / 2 + e * 10;
            const r2 a01 = *reinterpret_cast<const r2*>(A), a23 = *reinterpret_cast<const r2*>(A + 2), a45 = *reinterpret_cast<const r2*>(A + 4),
                     a67 = *reinterpret_cast<const r2*>(A + 6), a89 = *reinterpret_cast<const r2*>(A + 8);
            const r2 b01 = *reinterpret_cast<const r2*>(B), b23 = *reinterpret_cast<const r2*>(B + 2), b45 = *reinterpret_cast<const r2*>(B + 4),
                     b67 = *reinterpret_cast<const r2*>(B + 6), b89 = *reinterpret_cast<const r2*>(B + 8);
            r2 f4w = wc * a01.x, f3w = wc * b01.x, f4t = tc * a01.x, f3t = tc * b01.x;
#define FX_TERM(AW, BW, QW, QT) { f4w = fma2((AW), (QW), f4w); f3w = fma2((BW), (QW), f3w); f4t = fma2((AW), (QT), f4t); f3t = fma2((BW), (QT), f3t); }
            FX_TERM(a01.y, b01.y, wn[0], tn[0]) FX_TERM(a23.x, b23.x, wn[1], tn[1]) FX_TERM(a23.y, b23.y, wn[2], tn[2])
            FX_TERM(a45.x, b45.x, wn[3], tn[3]) FX_TERM(a45.y, b45.y, wn[4], tn[4]) FX_TERM(a67.x, b67.x, wn[5], tn[5])
            FX_TERM(a67.y, b67.y, wa, ta) FX_TERM(a89.x, b89.x, wb, tb) FX_TERM(a89.y, b89.y, wd, td)
#undef FX_TERM
            const r2 ruw = fm * ruk + fp * up1(ruk);               // ru at w levels
            const b2 pw = nonneg_sign(ruw), pt = nonneg_sign(ruk);
            const r2 fxw = ruw * sel(pw, f4w + f3w, f4w - f3w);
            const r2 fxt = ruk * sel(pt, f4t + f3t, f4t - f3t);
            tw = tw - sg * fxw;
            tt = tt - sg * fxt;
        }
    } else {
#if FX_TMA
        mbar_wait(bar0 + 8u * cur, ph);                  // (unused here, but every fill is waited for before its barrier is re-armed)
#endif
        if (inext < D.nCellsSolve) {
            nx_ring = D.fx_ring[(unsigned)inext * FX_RING + lr];
            nx_e = D.edgesOnCell[(unsigned)inext * D.maxEdges + le];
            nx_sgn = D.edgesOnCell_sign[(unsigned)inext * D.maxEdges + le];
        }
        // the reference's loop nest over the edges of the cell and their advCellsForEdge lists
        const int ne = D.nEdgesOnCell[i];
        for (int e = 0; e < ne; e++) {
            const int iEdge = D.edgesOnCell[(unsigned)i * D.maxEdges + e];
            const real sg = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + e];
            const int nadv = D.nAdvCellsForEdge[iEdge];
            int my_c = 0; real my_a = 0.0, my_b = 0.0;
            if (lane < nadv) {
                my_c = D.advCellsForEdge[(unsigned)iEdge * 15 + lane];
                my_a = D.adv_coefs[(unsigned)iEdge * 15 + lane]; my_b = D.adv_coefs_3rd[(unsigned)iEdge * 15 + lane];
            }
            const r2 ruk = LD(D.ru, iEdge);
            const r2 ruw = fm * ruk + fp * up1(ruk);
            r2 f4w = mk2(0.0, 0.0), f3w = f4w, f4t = f4w, f3t = f4w;
            for (int j = 0; j < nadv; j++) {
                const int c = BC(my_c, j);
                const real a = BC(my_a, j), b = BC(my_b, j);
                const r2 w2 = LD(D.w_2, c), t2 = LD(D.theta_m_2, c);
                f4w = fma2(a, w2, f4w); f3w = fma2(b, w2, f3w); f4t = fma2(a, t2, f4t); f3t = fma2(b, t2, f3t);
            }
            const b2 pw = nonneg_sign(ruw), pt = nonneg_sign(ruk);
            tw = tw - sg * (ruw * sel(pw, f4w + f3w, f4w - f3w));
            tt = tt - sg * (ruk * sel(pt, f4t + f3t, f4t - f3t));
        }
    }
    ST(D.hdiv_w, i, sel(k_mid, tw, 0.0));
    ST(D.hdiv_theta, i, sel(k_lt_nl, tt, 0.0));
    my_ring = nx_ring; my_e = nx_e; my_sgn = nx_sgn;
    }
#undef FX_COL
}

// ---- the same sweep with ONE FIELD PER WARP (warp 2c: w of cell c, warp 2c + 1: theta_m): half the registers per thread, twice
// the resident warps.  k5_flux_cell holds both fields' 19 columns (248 registers, 8 warps per SM) and is bound by the
// latency of its own instruction stream at 1.8 warps per scheduler (ncu: issue slots 32 % busy, L1 data pipe 50 %, DRAM 16 %).
#ifndef FX1_MINB
#define FX1_MINB 2
#endif
__global__ void __launch_bounds__(CW_THREADS, FX1_MINB) k5s_flux_cell(const Dev D) {
    __shared__ __align__(16) real s_wts[CW_WARPS][FX_WTS];
    PDL_ENTER
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = blockIdx.x * CW_WARPS + wib;
    const int field = g & 1;
    const int LDK = D.LDK, nl = D.nl;
    if ((g >> 1) >= D.nCellsSolve) return;
    const int i = g >> 1;
    Lv lv; lv.k0 = 2 * lane;
    const int k0 = lv.k0; const bool act = k0 < D.LDKA;
    const unsigned uLDK = (unsigned)LDK, kc = (unsigned)min(k0, D.LDKA - 2);
    const b2 k_lt_nl = lv.lt(nl), k_mid = lv.ge(1) && lv.lt(nl);
    const real* __restrict__ Q = field ? D.theta_m_2 : D.w_2;
    real* OUT = field ? D.hdiv_theta : D.hdiv_w;
    const int my_ring = D.fx_ring[(unsigned)i * FX_RING + min(lane, FX_RING - 1)];
    const int regular = BC(my_ring, 18);
    r2 acc = mk2(0.0, 0.0);
    if (regular) {
        const int le = min(lane, 5);
        const int my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
        const real my_sgn = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le];
        const real* __restrict__ wsrc = D.fx_w + (size_t)i * FX_WTS + 4 * lane;
        const r2 wts_a = *reinterpret_cast<const r2*>(wsrc), wts_b = *reinterpret_cast<const r2*>(wsrc + 2);
        const r2 qc = LD(Q, i);
        r2 qn[6], qm[12], ru6[6];
#pragma unroll
        for (int q = 0; q < 6; q++) qn[q] = LD(Q, BC(my_ring, q));
#pragma unroll
        for (int q = 0; q < 12; q++) qm[q] = LD(Q, BC(my_ring, 6 + q));
#pragma unroll
        for (int q = 0; q < 6; q++) ru6[q] = LD(D.ru, BC(my_e, q));
        *reinterpret_cast<r2*>(&s_wts[wib][4 * lane]) = wts_a; *reinterpret_cast<r2*>(&s_wts[wib][4 * lane + 2]) = wts_b;
        __syncwarp();
        const real* __restrict__ W = s_wts[wib];
        r2 fm = mk2(1.0, 1.0), fp = mk2(0.0, 0.0);
        if (!field) { fm = LD(D.fzm, 0); fp = LD(D.fzp, 0); }
#pragma unroll
        for (int e = 0; e < 6; e++) {
            const r2 qa = qm[(2 * e + 11) % 12], qb = qm[2 * e], qd = qm[(2 * e + 1) % 12];
            const real sg = BC(my_sgn, e);
            const real* __restrict__ A = W + e * 10;
            const real* __restrict__ B = W + FX_WTS / 2 + e * 10;
            const r2 a01 = *reinterpret_cast<const r2*>(A), a23 = *reinterpret_cast<const r2*>(A + 2), a45 = *reinterpret_cast<const r2*>(A + 4),
                     a67 = *reinterpret_cast<const r2*>(A + 6), a89 = *reinterpret_cast<const r2*>(A + 8);
            const r2 b01 = *reinterpret_cast<const r2*>(B), b23 = *reinterpret_cast<const r2*>(B + 2), b45 = *reinterpret_cast<const r2*>(B + 4),
                     b67 = *reinterpret_cast<const r2*>(B + 6), b89 = *reinterpret_cast<const r2*>(B + 8);
            r2 f4 = qc * a01.x, f3 = qc * b01.x;
#define FX1_TERM(AW, BW, QQ) { f4 = fma2((AW), (QQ), f4); f3 = fma2((BW), (QQ), f3); }
            FX1_TERM(a01.y, b01.y, qn[0]) FX1_TERM(a23.x, b23.x, qn[1]) FX1_TERM(a23.y, b23.y, qn[2])
            FX1_TERM(a45.x, b45.x, qn[3]) FX1_TERM(a45.y, b45.y, qn[4]) FX1_TERM(a67.x, b67.x, qn[5])
            FX1_TERM(a67.y, b67.y, qa) FX1_TERM(a89.x, b89.x, qb) FX1_TERM(a89.y, b89.y, qd)
#undef FX1_TERM
            const r2 ruk = ru6[e];
            const r2 rue = field ? ruk : fm * ruk + fp * up1(ruk);     // theta_m: ru at mass levels; w: ru at w levels
            const b2 pos = nonneg_sign(rue);
            acc = acc - sg * (rue * sel(pos, f4 + f3, f4 - f3));
        }
    } else {
        const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0);
        const int ne = D.nEdgesOnCell[i];
        for (int e = 0; e < ne; e++) {
            const int iEdge = D.edgesOnCell[(unsigned)i * D.maxEdges + e];
            const real sg = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + e];
            const int nadv = D.nAdvCellsForEdge[iEdge];
            int my_c = 0; real my_a = 0.0, my_b = 0.0;
            if (lane < nadv) {
                my_c = D.advCellsForEdge[(unsigned)iEdge * 15 + lane];
                my_a = D.adv_coefs[(unsigned)iEdge * 15 + lane]; my_b = D.adv_coefs_3rd[(unsigned)iEdge * 15 + lane];
            }
            const r2 ruk = LD(D.ru, iEdge);
            const r2 rue = field ? ruk : fm * ruk + fp * up1(ruk);
            r2 f4 = mk2(0.0, 0.0), f3 = f4;
            for (int j = 0; j < nadv; j++) {
                const r2 q2 = LD(Q, BC(my_c, j));
                f4 = fma2(BC(my_a, j), q2, f4); f3 = fma2(BC(my_b, j), q2, f3);
            }
            const b2 pos = nonneg_sign(rue);
            acc = acc - sg * (rue * sel(pos, f4 + f3, f4 - f3));
        }
    }
    ST(OUT, i, field ? sel(k_lt_nl, acc, 0.0) : sel(k_mid, acc, 0.0));
}

// owned cells: tend_w (TI:5713-5757, 5838-5945) and tend_theta (TI:5956-6016, 6066-6126, 6134-6197).
// Restrictions (the host falls back to k_dt_cell_f otherwise): v_mom_eddy_visc2 == v_theta_eddy_visc2 == 0.
#ifndef CELLF_MINB
#define CELLF_MINB 3
#endif
// HDIV: the horizontal flux divergences come ready-made from k5_flux_cell (relaxed arithmetic) instead of being summed
// here from the per-edge fluxes of k2_dt_edge_flux
template <bool HDIV>
__global__ void __launch_bounds__(CW_THREADS, CELLF_MINB) k2_dt_cell_f(const Dev D, const DynTendArgs A) {
    CW_SETUP_R(D.nCellsSolve)
    const int ne = D.nEdgesOnCell[i];
    // one edge of the cell per lane (lanes >= ne repeat the last edge): id, sign, mixing metadata
    const int le = min(lane, ne - 1);
    const int my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    const real my_sgn = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le];
    const int my_c1 = D.cellsOnEdge[2 * my_e], my_c2 = D.cellsOnEdge[2 * my_e + 1];
    const real my_dv = D.dvEdge[my_e];
    real my_d4 = 0.0, my_idc = 0.0;
    if (A.rk_step == 1) { my_d4 = D.meshScalingDel4[my_e]; my_idc = D.invDcEdge[my_e]; }
    const real invArea = D.invAreaCell[i];
    const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0);
    r2 tw = mk2(0.0, 0.0), tt = mk2(0.0, 0.0);
    // horizontal flux divergence of w and theta_m from the per-edge fluxes
#define CELL_F_EDGE(E)                                                                                      \
    {                                                                                                       \
        const int iEdge = BC(my_e, (E));                                                                    \
        const real sg = BC(my_sgn, (E));                                                                    \
        const r2 fxw = LD(D.adv_flux_w, iEdge), fxt = LD(D.adv_flux_theta, iEdge);                          \
        tw = selb((E) < ne, tw - sg * fxw, tw);                                                             \
        tt = selb((E) < ne, tt - sg * fxt, tt);                                                             \
    }
    if (HDIV) { tw = LD(D.hdiv_w, i); tt = LD(D.hdiv_theta, i); }
    else {
#pragma unroll
        for (int e = 0; e < CW_NE; e++) CELL_F_EDGE(e)
        for (int e = CW_NE; e < ne; e++) CELL_F_EDGE(e)
    }
#undef CELL_F_EDGE
    r2 twe = LD(D.tend_w_euler, i), tte = LD(D.tend_theta_euler, i);
    const r2 twe_in = twe;
    if (A.rk_step > 1) {          // perturbation flux for the rtheta_pp equation, TI:5995-6016
#define CELL_F_PERT(E)                                                                                      \
        {                                                                                                   \
            const int iEdge = BC(my_e, (E)), cell1 = BC(my_c1, (E)), cell2 = BC(my_c2, (E));                \
            const real sg = BC(my_sgn, (E)), dv = BC(my_dv, (E));                                           \
            const r2 flux = sg * dv * (LD(D.ru_save, iEdge) - LD(D.ru, iEdge)) * 0.5 * (LD(D.theta_m, cell2) + LD(D.theta_m, cell1)); \
            tt = selb((E) < ne, tt - flux, tt);                                                             \
        }
#pragma unroll 3
        for (int e = 0; e < CW_NE; e++) CELL_F_PERT(e)
        for (int e = CW_NE; e < ne; e++) CELL_F_PERT(e)
#undef CELL_F_PERT
    }
    const b2 k_ge1 = lv.ge(1), k_lt_nl = lv.lt(nl);
    if (A.rk_step == 1) {
#define CELL_F_DEL4(E, ACC, FIELD, RAREA)                                                                   \
        {                                                                                                   \
            const int cell1 = BC(my_c1, (E)), cell2 = BC(my_c2, (E));                                       \
            const real edge_sign = BC(my_d4, (E)) * (RAREA) * BC(my_dv, (E)) * BC(my_sgn, (E)) * BC(my_idc, (E)); \
            const r2 d = LD(FIELD, cell2) - LD(FIELD, cell1);                                               \
            ACC = selb((E) < ne, ACC - edge_sign * d, ACC);                                                 \
        }
        if (A.h_mom_eddy_visc4 > 0.0) {
            const real r_areaCell = A.h_mom_eddy_visc4 * invArea;
#pragma unroll
            for (int e = 0; e < CW_NE; e++) CELL_F_DEL4(e, twe, D.delsq_w, r_areaCell)
            for (int e = CW_NE; e < ne; e++) CELL_F_DEL4(e, twe, D.delsq_w, r_areaCell)
        }
        if (A.h_theta_eddy_visc4 > 0.0) {
            const real r_areaCell = A.h_theta_eddy_visc4 * A.prandtl_inv * invArea;
#pragma unroll
            for (int e = 0; e < CW_NE; e++) CELL_F_DEL4(e, tte, D.delsq_theta, r_areaCell)
            for (int e = CW_NE; e < ne; e++) CELL_F_DEL4(e, tte, D.delsq_theta, r_areaCell)
        }
#undef CELL_F_DEL4
    }
    // own-column operands of the vertical terms (loaded after the edge loops: keeps the kernel at <= 80
    // registers, i.e. 24 resident warps per SM, which matters more here than one extra exposed latency)
    const r2 rdzu = LD(D.rdzu, 0), rdzw = LD(D.rdzw, 0);
    const r2 rw = LD(D.rw, i), w = LD(D.w_2, i), t = LD(D.theta_m_2, i), ts = LD(D.theta_m, i), rws = LD(D.rw_save, i);
    const r2 rho = LD(D.rho_zz_2, i), tend_rho = LD(D.tend_rho, i), rtdiab = LD(D.rt_diabatic_tend, i);
    const r2 trp = LD(D.tend_rtheta_physics, i);
    r2 pp = mk2(0.0, 0.0), dpdz = mk2(0.0, 0.0), cqw = mk2(0.0, 0.0);
    if (A.rk_step == 1) { pp = LD(D.pressure_p, i); dpdz = LD(D.dpdz, i); cqw = LD(D.cqw, i); }
    const r2 rwm1 = up1(rw);
    const b2 kk_edge = lv.eq(1) || lv.eq(nl - 1);            // 2nd-order interfaces
    const b2 kk_zero = lv.lt(1) || lv.ge(nl);                // no flux through the boundaries
    // ---- w: vertical advection (TI:5878-5891), pressure gradient, buoyancy
    {
        const r2 wm1 = up1(w), wm2 = up2(w), wp1 = dn1(w);
        const r2 f2 = 0.25 * (rw + rwm1) * (w + wm1);
        const r2 f3 = flux3_2(wm2, wm1, w, wp1, 0.5 * (rw + rwm1), 1.0);
        const r2 fz = sel(kk_zero, mk2(0.0, 0.0), sel(kk_edge, f2, f3));     // flux at interface k
        const r2 f1 = dn1(fz);
        tw = tw * invArea - rdzu * (f1 - fz);
        if (A.rk_step == 1) {
            const r2 twe_new = twe - cqw * (rdzu * (pp - up1(pp)) - (fm * dpdz + fp * up1(dpdz)));
            twe = sel(k_ge1 && k_lt_nl, twe_new, twe_in);    // rows 0 and nl keep what they held
        }
    }
    const r2 out_tend_w = sel(k_ge1 && k_lt_nl, tw + twe, 0.0);
    // ---- theta_m: vertical advection (TI:6101-6116), mixing
    r2 out_rthdynten;
    {
        const r2 tm1 = up1(t), tm2 = up2(t), tp1 = dn1(t), tsm1 = up1(ts);
        const r2 ftop = rws * (fm * t + fp * tm1);                                   // kk == nl-1
        const r2 flow = rw * (fm * t + fp * tm1);                                    // kk == 1
        const r2 f3 = flux3_2(tm2, tm1, t, tp1, rw, A.coef_3rd_order);
        const r2 fpert = sel(lv.eq(1), flow, f3) + (rws - rw) * (fm * ts + fp * tsm1);
        const r2 fz = sel(kk_zero, mk2(0.0, 0.0), sel(lv.eq(nl - 1), ftop, fpert));
        const r2 f1 = dn1(fz);
        tt = tt * invArea - rdzw * (f1 - fz);
        out_rthdynten = sel(k_lt_nl, (tt - tend_rho * t) / rho, 0.0);
        tt = tt + rho * rtdiab;
    }
    if (A.rk_step == 1) {
        ST(D.tend_w_euler, i, twe);
        ST(D.tend_theta_euler, i, sel(k_lt_nl, tte, 0.0));
    }
    ST(D.tend_w, i, out_tend_w);
    ST(D.rthdynten, i, out_rthdynten);
    ST(D.tend_theta, i, sel(k_lt_nl, tt + tte + trp, 0.0));
}

// ---- the same cell tendency for the relaxed path: horizontal flux divergences from k5_flux_cell, persistent warps ----
// The kernel is a chain of dependent round trips (edgesOnCell -> cellsOnEdge -> gathered columns), i.e. latency bound at one
// column per warp.  Here a warp walks cells g, g + G, ..., fetches the NEXT cell's connectivity while the current cell's
// columns are in flight, and requests the own-column operands together with the gathers.
struct CfConn { int ne, e, c1, c2; real sgn, dv, d4, idc, invArea; };
__device__ __forceinline__ CfConn cf_conn(const Dev& D, int i, int lane, bool rk1) {
    CfConn c;
    c.ne = D.nEdgesOnCell[i];
    const int le = min(lane, c.ne - 1);
    c.e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    c.sgn = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le];
    c.c1 = D.cellsOnEdge[2 * c.e]; c.c2 = D.cellsOnEdge[2 * c.e + 1];
    c.dv = D.dvEdge[c.e];
    c.d4 = 0.0; c.idc = 0.0;
    if (rk1) { c.d4 = D.meshScalingDel4[c.e]; c.idc = D.invDcEdge[c.e]; }
    c.invArea = D.invAreaCell[i];
    return c;
}
#ifndef CF7_MINB
#define CF7_MINB 2
#endif
template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k7_dt_cell_f(const Dev D, const DynTendArgs A) {
    pdl_trigger();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int LDK = D.LDK, nl = D.nl;
    Lv lv; lv.k0 = 2 * lane;
    const int k0 = lv.k0; const bool act = k0 < D.LDKA;
    const unsigned uLDK = (unsigned)LDK, kc = (unsigned)min(k0, D.LDKA - 2);
    const bool rk1 = A.rk_step == 1;
    const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0), rdzu = LD(D.rdzu, 0), rdzw = LD(D.rdzw, 0);
    const b2 k_ge1 = lv.ge(1), k_lt_nl = lv.lt(nl);
    const b2 kk_edge = lv.eq(1) || lv.eq(nl - 1);            // 2nd-order interfaces
    const b2 kk_zero = lv.lt(1) || lv.ge(nl);                // no flux through the boundaries
    const Pers ps = pers_init<WARPS, MINB>(D.nCellsSolve, wib);
    const int G = ps.stride, nSolve = ps.end;
#define CF7_COL(j) (D.nCellsSolve - 1 - (j))                  /* backward sweep (see CW_SETUP_R) */
    int jcol = ps.j;
    CfConn cn; cn.ne = 1; cn.e = 0; cn.c1 = 0; cn.c2 = 0; cn.sgn = 0.0; cn.dv = 0.0; cn.d4 = 0.0; cn.idc = 0.0; cn.invArea = 0.0;
    if (jcol < nSolve) cn = cf_conn(D, CF7_COL(jcol), lane, rk1);
    pdl_wait();                                          // everything above is static mesh data
    for (; jcol < nSolve; jcol += G) {
        const int i = CF7_COL(jcol);
        const int ne = cn.ne, my_e = cn.e, my_c1 = cn.c1, my_c2 = cn.c2;
        const real my_sgn = cn.sgn, my_dv = cn.dv, my_d4 = cn.d4, my_idc = cn.idc, invArea = cn.invArea;
        // own-column operands: requested up front, together with the gathers below
        r2 tw = LD(D.hdiv_w, i), tt = LD(D.hdiv_theta, i);
        r2 twe = LD(D.tend_w_euler, i), tte = LD(D.tend_theta_euler, i);
        const r2 rw = LD(D.rw, i), w = LD(D.w_2, i), t = LD(D.theta_m_2, i), ts = LD(D.theta_m, i), rws = LD(D.rw_save, i);
        const r2 rho = LD(D.rho_zz_2, i), tend_rho = LD(D.tend_rho, i), rtdiab = LD(D.rt_diabatic_tend, i);
        const r2 trp = LD(D.tend_rtheta_physics, i);
        r2 pp = mk2(0.0, 0.0), dpdz = mk2(0.0, 0.0), cqw = mk2(0.0, 0.0);
        if (rk1) { pp = LD(D.pressure_p, i); dpdz = LD(D.dpdz, i); cqw = LD(D.cqw, i); }
        const r2 twe_in = twe;
        if (!rk1) {               // perturbation flux for the rtheta_pp equation, TI:5995-6016
#define CF7_PERT(E)                                                                                         \
            {                                                                                               \
                const int iEdge = BC(my_e, (E)), cell1 = BC(my_c1, (E)), cell2 = BC(my_c2, (E));            \
                const real sg = BC(my_sgn, (E)), dv = BC(my_dv, (E));                                       \
                const r2 flux = sg * dv * (LD(D.ru_save, iEdge) - LD(D.ru, iEdge)) * 0.5 * (LD(D.theta_m, cell2) + LD(D.theta_m, cell1)); \
                tt = selb((E) < ne, tt - flux, tt);                                                         \
            }
#pragma unroll
            for (int e = 0; e < CW_NE; e++) CF7_PERT(e)
            for (int e = CW_NE; e < ne; e++) CF7_PERT(e)
#undef CF7_PERT
        } else {
#define CF7_DEL4(E, ACC, FIELD, RAREA)                                                                      \
            {                                                                                               \
                const int cell1 = BC(my_c1, (E)), cell2 = BC(my_c2, (E));                                   \
                const real edge_sign = BC(my_d4, (E)) * (RAREA) * BC(my_dv, (E)) * BC(my_sgn, (E)) * BC(my_idc, (E)); \
                const r2 d = LD(FIELD, cell2) - LD(FIELD, cell1);                                           \
                ACC = selb((E) < ne, ACC - edge_sign * d, ACC);                                             \
            }
            if (A.h_mom_eddy_visc4 > 0.0) {
                const real r_areaCell = A.h_mom_eddy_visc4 * invArea;
#pragma unroll
                for (int e = 0; e < CW_NE; e++) CF7_DEL4(e, twe, D.delsq_w, r_areaCell)
                for (int e = CW_NE; e < ne; e++) CF7_DEL4(e, twe, D.delsq_w, r_areaCell)
            }
            if (A.h_theta_eddy_visc4 > 0.0) {
                const real r_areaCell = A.h_theta_eddy_visc4 * A.prandtl_inv * invArea;
#pragma unroll
                for (int e = 0; e < CW_NE; e++) CF7_DEL4(e, tte, D.delsq_theta, r_areaCell)
                for (int e = CW_NE; e < ne; e++) CF7_DEL4(e, tte, D.delsq_theta, r_areaCell)
            }
#undef CF7_DEL4
        }
        if (jcol + G < nSolve) {
            const int j = CF7_COL(jcol + G);
            cn = cf_conn(D, j, lane, rk1);              // next cell's connectivity, behind this cell's requests
            if (D.pf_next) {
                PF(D.hdiv_w, j); PF(D.hdiv_theta, j); PF(D.tend_w_euler, j); PF(D.tend_theta_euler, j); PF(D.rw, j); PF(D.w_2, j);
                PF(D.theta_m_2, j); PF(D.theta_m, j); PF(D.rw_save, j); PF(D.rho_zz_2, j); PF(D.tend_rho, j); PF(D.rt_diabatic_tend, j);
                PF(D.tend_rtheta_physics, j);
                if (rk1) { PF(D.pressure_p, j); PF(D.dpdz, j); PF(D.cqw, j); }
            }
        }
        const r2 rwm1 = up1(rw);
        {   // ---- w: vertical advection (TI:5878-5891), pressure gradient, buoyancy
            const r2 wm1 = up1(w), wm2 = up2(w), wp1 = dn1(w);
            const r2 f2 = 0.25 * (rw + rwm1) * (w + wm1);
            const r2 f3 = flux3_2(wm2, wm1, w, wp1, 0.5 * (rw + rwm1), 1.0);
            const r2 fz = sel(kk_zero, mk2(0.0, 0.0), sel(kk_edge, f2, f3));     // flux at interface k
            const r2 f1 = dn1(fz);
            tw = tw * invArea - rdzu * (f1 - fz);
            if (rk1) {
                const r2 twe_new = twe - cqw * (rdzu * (pp - up1(pp)) - (fm * dpdz + fp * up1(dpdz)));
                twe = sel(k_ge1 && k_lt_nl, twe_new, twe_in);    // rows 0 and nl keep what they held
            }
        }
        const r2 out_tend_w = sel(k_ge1 && k_lt_nl, tw + twe, 0.0);
        r2 out_rthdynten;
        {   // ---- theta_m: vertical advection (TI:6101-6116), mixing
            const r2 tm1 = up1(t), tm2 = up2(t), tp1 = dn1(t), tsm1 = up1(ts);
            const r2 ftop = rws * (fm * t + fp * tm1);                                   // kk == nl-1
            const r2 flow = rw * (fm * t + fp * tm1);                                    // kk == 1
            const r2 f3 = flux3_2(tm2, tm1, t, tp1, rw, A.coef_3rd_order);
            const r2 fpert = sel(lv.eq(1), flow, f3) + (rws - rw) * (fm * ts + fp * tsm1);
            const r2 fz = sel(kk_zero, mk2(0.0, 0.0), sel(lv.eq(nl - 1), ftop, fpert));
            const r2 f1 = dn1(fz);
            tt = tt * invArea - rdzw * (f1 - fz);
            out_rthdynten = sel(k_lt_nl, (tt - tend_rho * t) / rho, 0.0);
            tt = tt + rho * rtdiab;
        }
        if (rk1) {
            ST(D.tend_w_euler, i, twe);
            ST(D.tend_theta_euler, i, sel(k_lt_nl, tte, 0.0));
        }
        ST(D.tend_w, i, out_tend_w);
        ST(D.rthdynten, i, out_rthdynten);
        ST(D.tend_theta, i, sel(k_lt_nl, tt + tte + trp, 0.0));
    }
#undef CF7_COL
}

// ------------------------------------------------------------------ nonlinear Coriolis term, cell-centred partial sums
// (relaxed arithmetic; TI:5418-5428).  One warp per cell c of the block: u and pv_edge of its ne edges are gathered once and for
// every edge i of the cell P_c[i] = sum_{j != i} W[i][j] u_j (pv_i + pv_j)/2 is formed in registers, W[i][j] = weightsOnEdge of
// edge j in the edgesOnEdge list of edge i (tables built and verified on the host, build_coriolis_tables).  The edge kernel adds
// the two partial sums of its two cells: 2 gathered columns per edge instead of 2 (ne1 + ne2 - 2).
__global__ void __launch_bounds__(CW_THREADS, 2) k8_coriolis_cell(const Dev D) {
    __shared__ __align__(16) real s_w[CW_WARPS][64];
    CW_SETUP(D.nCells)
    const int ne = D.nEdgesOnCell[i];
    const int le = min(lane, ne - 1);
    const int my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    {   // 64 weights of the cell: two per lane
        const r2 w2 = *reinterpret_cast<const r2*>(D.cor_w + (size_t)i * 64 + 2 * lane);
        *reinterpret_cast<r2*>(&s_w[wib][2 * lane]) = w2;
    }
    r2 u[CW_MAXNE], pv[CW_MAXNE];
#pragma unroll
    for (int e = 0; e < CW_MAXNE; e++) {
        if (e < CW_NE || e < ne) { const int iEdge = BC(my_e, min(e, ne - 1)); u[e] = LD(D.u_2, iEdge); pv[e] = LD(D.pv_edge, iEdge); }
        else { u[e] = mk2(0.0, 0.0); pv[e] = u[e]; }
    }
    __syncwarp();
    const real* __restrict__ W = s_w[wib];
    const b2 k_lt_nl = lv.lt(nl);
#pragma unroll
    for (int e = 0; e < CW_MAXNE; e++) {
        if (e < ne) {                                       // warp-uniform
            r2 acc = mk2(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < CW_MAXNE; j++) {
                if (j != e && (j < CW_NE || j < ne)) {
                    const real w = W[e * 8 + j];            // zero for slots beyond ne
                    const r2 upv = u[j] * (0.5 * (pv[e] + pv[j]));
                    acc = fma2(w, upv, acc);
                }
            }
            st2(D.cor_part, ((unsigned)i * (unsigned)D.maxEdges + (unsigned)e) * uLDK + kc, act, sel(k_lt_nl, acc, 0.0));
        }
    }
}

// ------------------------------------------------------------------ atm_compute_dyn_tend_work, part (b)
// edge-all: [rk 1] PGF (5379-5387), delsq_u + del2 mixing (5467-5503); [owned edges] vertical transport,
// nonlinear Coriolis, KE gradient (5391-5447); [rk > 1] final sum with tend_u_euler (5694-5701).
// Restriction (host falls back to k_dt_edge_b otherwise): config_rayleigh_damp_u off.
// This kernel is bound by the L1 data pipe (46 gathered columns per edge), not by latency: it is kept at
// ~64 registers for 32 resident warps per SM rather than unrolled for more loads in flight (measured).
// COR: the nonlinear Coriolis sum arrives as two partial sums, one from each adjacent cell (k8_coriolis_cell, relaxed
// arithmetic), instead of being gathered here over edgesOnEdge
template <bool COR>
__global__ void __launch_bounds__(EB_WARPS * 32, MB_EDGE_B) k2_dt_edge_b(const Dev D, const DynTendArgs A) {
    CW_ENTER_R(D.nEdges)
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    const bool solve = i < D.nEdgesSolve;
    const real invDc = D.invDcEdge[i];
    const b2 k_lt_nl = lv.lt(nl);
    pdl_wait();
    const r2 rho_e = LD(D.rho_edge, i);
    if (A.rk_step == 1) {
        r2 tue = mk2(0.0, 0.0);
        if (solve)
            tue = -LD(D.cqu, i) * ((LD(D.pressure_p, cell2) - LD(D.pressure_p, cell1)) * invDc / (.5 * (LD(D.zz, cell2) + LD(D.zz, cell1)))
                                   - 0.5 * LD(D.zxu, i) * (LD(D.dpdz, cell1) + LD(D.dpdz, cell2)));
        const int vertex1 = D.verticesOnEdge[2 * i], vertex2 = D.verticesOnEdge[2 * i + 1];
        const real r_dc = invDc;
        const real r_dv = rmin(D.invDvEdge[i], 4 * invDc);
        const r2 u_diffusion = (LD(D.divergence, cell2) - LD(D.divergence, cell1)) * r_dc
                               - (LD(D.vorticity, vertex2) - LD(D.vorticity, vertex1)) * r_dv;
        ST(D.delsq_u, i, sel(k_lt_nl, 0.0 + u_diffusion, 0.0));
        const r2 kdiffu = 0.5 * (LD(D.kdiff, cell1) + LD(D.kdiff, cell2));
        tue = tue + rho_e * kdiffu * u_diffusion * D.meshScalingDel2[i];
        ST(D.tend_u_euler, i, sel(k_lt_nl, tue, 0.0));
    }
    if (!solve) return;
    const r2 u = LD(D.u_2, i);
    r2 tu;
    {   // vertical transport of u, TI:5391-5408
        const r2 rwf = LD(D.rw, cell1) + LD(D.rw, cell2);
        const r2 um1 = up1(u), um2 = up2(u), up = dn1(u);
        const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0);
        const r2 f2 = 0.5 * (rwf) * (fm * u + fp * um1);
        const r2 f3 = flux3_2(um2, um1, u, up, 0.5 * (rwf), 1.0);
        const b2 kk_edge = lv.eq(1) || lv.eq(nl - 1);
        const b2 kk_zero = lv.lt(1) || lv.ge(nl);
        const r2 fz = sel(kk_zero, mk2(0.0, 0.0), sel(kk_edge, f2, f3));
        const r2 f1 = dn1(fz);
        tu = -LD(D.rdzw, 0) * (f1 - fz);
    }
    r2 q = mk2(0.0, 0.0);
    if (COR) {
        const int sl = D.cor_slot[i];
        const r2 q1 = ld2(D.cor_part, ((unsigned)cell1 * (unsigned)D.maxEdges + (unsigned)(sl & 255)) * uLDK + kc);
        const r2 q2 = ld2(D.cor_part, ((unsigned)cell2 * (unsigned)D.maxEdges + (unsigned)(sl >> 8)) * uLDK + kc);
        q = q1 + q2;
    } else {
        const int neoe = D.nEdgesOnEdge[i];
        int my_eoe = 0; real my_woe = 0.0;
        if (lane < neoe) { my_eoe = D.edgesOnEdge[(unsigned)i * D.maxEdges2 + lane]; my_woe = D.weightsOnEdge[(unsigned)i * D.maxEdges2 + lane]; }
        const r2 pv_e = LD(D.pv_edge, i);
#ifndef EB_UNROLL
#define EB_UNROLL 5
#endif
        constexpr int eb_unroll = EB_UNROLL;
#pragma unroll eb_unroll
        for (int j = 0; j < neoe; j++) {
            const int eoe = BC(my_eoe, j);
            const real woe = BC(my_woe, j);
            const r2 workpv = 0.5 * (pv_e + LD(D.pv_edge, eoe));
            q = q + woe * LD(D.u_2, eoe) * workpv;
        }
    }
    tu = tu + rho_e * (q - (LD(D.ke, cell2) - LD(D.ke, cell1)) * invDc)
         - u * 0.5 * (LD(D.h_divergence, cell1) + LD(D.h_divergence, cell2));
    if (A.rk_step != 1) tu = tu + LD(D.tend_u_euler, i) + LD(D.tend_ru_physics, i);
    ST(D.tend_u, i, sel(k_lt_nl, tu, 0.0));
}

// (an earlier one-warp-per-cell version of the acoustic cell step, with lane 0 sweeping the column out of shared memory --
//  741 dynamic instructions per column in the sweep against 34 in the block-tiled k3_acoustic_cell below -- was removed)

// ------------------------------------------------------------------ atm_set_smlstep_pert_variables_work  TI:2427-2508
// zb_cell/zb3_cell are [cell][edge slot][LDK]; requires maxEdges >= CW_NE (slots beyond nEdgesOnCell exist and are skipped)
#define PFZ(p, col, E) pf2((p), ((unsigned)(col) * (unsigned)D.maxEdges + (unsigned)(E)) * uLDK + kc)
#define LDZ(p, E) ld2((p), ((unsigned)i * (unsigned)D.maxEdges + (unsigned)(E)) * uLDK + kc)
__global__ void __launch_bounds__(CW_THREADS, MB_SML) k2_smlstep_pert(const Dev D) {
    CW_ENTER(D.nCellsSolve)
    const int ne = D.nEdgesOnCell[i];
    const int le = min(lane, ne - 1);
    const int my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    const real my_sgn = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le];
    const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0);
    pdl_wait();
    const r2 wt_in = LD(D.tend_w, i);
    const r2 zz = LD(D.zz, i);
    r2 wt = wt_in;
    if (D.zb_any[i]) {
#define SML_EDGE(E)                                                                                         \
        {                                                                                                   \
            const int iEdge = BC(my_e, (E));                                                                \
            const real sg = BC(my_sgn, (E));                                                                \
            const r2 tu = LD(D.tend_u, iEdge);                                                              \
            const r2 zb = LDZ(D.zb_cell, (E)), zb3 = LDZ(D.zb3_cell, (E));                                  \
            const r2 flux = sg * (fm * tu + fp * up1(tu));                                                  \
            const r2 szb3 = mk2(signbit(tu.x) ? -zb3.x : zb3.x, signbit(tu.y) ? -zb3.y : zb3.y);   /* sign(1,tend_u) * zb3 */ \
            wt = selb((E) < ne, wt - (zb + szb3) * flux, wt);                                               \
        }
#pragma unroll
        for (int e = 0; e < CW_NE; e++) SML_EDGE(e)
        for (int e = CW_NE; e < ne; e++) SML_EDGE(e)
#undef SML_EDGE
    }
    // regional run: no conversion in the specified zone, TI:2482.  Only the store is predicated, with the (static, cache-resident)
    // mask read last: an early exit puts the mask load in front of every other load of the kernel (one more exposed memory
    // round trip, measured +5 %), a flag read first and kept costs the register-capped kernel another spill
    st2(D.tend_w, (unsigned)i * uLDK + kc, act && D.bdyMaskCell[i] <= 5, sel(lv.ge(1) && lv.lt(nl), (fm * zz + fp * up1(zz)) * wt, wt_in));
}

// ------------------------------------------------------------------ atm_recover_large_step_variables_work, part 3  TI:3379-3416
__global__ void __launch_bounds__(CW_THREADS, MB_REC2) k2_recover_cell2(const Dev D, real cf1, real cf2, real cf3) {
    CW_ENTER(D.nCells)
    const int ne = D.nEdgesOnCell[i];
    const int le = min(lane, ne - 1);
    const int my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    const real my_sgn = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le];
    const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0);
    pdl_wait();
    const r2 w_in = LD(D.w_2, i);
    const r2 rho = LD(D.rho_zz_2, i);
    const b2 k_eq0 = lv.eq(0);
    r2 w = w_in;
    if (D.zb_any[i]) {
#define REC_EDGE(E)                                                                                         \
        {                                                                                                   \
            const int iEdge = BC(my_e, (E));                                                                \
            const real sg = BC(my_sgn, (E));                                                                \
            const r2 ru = LD(D.ru, iEdge);                                                                  \
            const r2 zb = LDZ(D.zb_cell, (E)), zb3 = LDZ(D.zb3_cell, (E));                                  \
            const r2 flux = sel(k_eq0, cf1 * ru + cf2 * dn1(ru) + cf3 * dn2(ru), fm * ru + fp * up1(ru));   \
            const r2 szb3 = mk2(signbit(flux.x) ? -zb3.x : zb3.x, signbit(flux.y) ? -zb3.y : zb3.y);        \
            w = selb((E) < ne, w + sg * (zb + szb3) * flux, w);                                             \
        }
#pragma unroll
        for (int e = 0; e < CW_NE; e++) REC_EDGE(e)
        for (int e = CW_NE; e < ne; e++) REC_EDGE(e)
#undef REC_EDGE
    }
    const r2 den = sel(k_eq0, cf1 * rho + cf2 * dn1(rho) + cf3 * dn2(rho), fm * rho + fp * up1(rho));
    // regional run: no update in the specified zone, TI:3385 (store predicated, mask read last: see k2_smlstep_pert)
    st2(D.w_2, (unsigned)i * uLDK + kc, act && D.bdyMaskCell[i] <= 5, sel(lv.lt(nl), w / den, w_in));
}

// ------------------------------------------------------------------ atm_compute_solve_diagnostics_work  TI:6337-6773
// (1) vertex-all: vorticity (6452-6472), ke_vertex (6548-6561, ke_edge recomputed inline), pv_vertex (6647-6659)
__global__ void __launch_bounds__(CW_THREADS) k2_diag_vertex(const Dev D, const real* u) {
    CW_ENTER_R(D.nVertices)
    int my_e = 0; real my_s = 0.0, my_efac = 0.0;
    {
        const int l3 = min(lane, 2);
        my_e = D.edgesOnVertex[3 * i + l3];
        const real dc = D.dcEdge[my_e];
        my_s = D.edgesOnVertex_sign[3 * i + l3] * dc;
        my_efac = dc * D.dvEdge[my_e];
    }
    pdl_wait();
    const r2 u0 = LD(u, BC(my_e, 0)), u1 = LD(u, BC(my_e, 1)), u2 = LD(u, BC(my_e, 2));
    r2 vort = mk2(0.0, 0.0);
    vort = vort + BC(my_s, 0) * u0;
    vort = vort + BC(my_s, 1) * u1;
    vort = vort + BC(my_s, 2) * u2;
    const r2 ke0 = BC(my_efac, 0) * (u0 * u0), ke1 = BC(my_efac, 1) * (u1 * u1), ke2 = BC(my_efac, 2) * (u2 * u2);
    const real iat = D.invAreaTriangle[i];
    vort = vort * iat;
    const b2 k_lt_nl = lv.lt(nl);
    const real r = 0.25 * iat;
    ST(D.vorticity, i, sel(k_lt_nl, vort, 0.0));
    ST(D.ke_vertex, i, sel(k_lt_nl, (ke0 + ke1 + ke2) * r, 0.0));
    ST(D.pv_vertex, i, sel(k_lt_nl, D.fVertex[i] + vort, 0.0));
}
// (2) cell-all: divergence (6479-6499), ke (6515-6534) + Hollingsworth blend (6569-6593), pv_cell (6693-6709)
__global__ void __launch_bounds__(CW_THREADS, MB_DIAG_C) k2_diag_cell(const Dev D, const real* u, int apvm) {
    CW_ENTER(D.nCells)
    const int ne = D.nEdgesOnCell[i];
    const int le = min(lane, ne - 1);
    const real r = D.invAreaCell[i];
    const int my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    const real my_dv = D.dvEdge[my_e];
    const real my_s = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le] * my_dv;
    const real my_efac = D.dcEdge[my_e] * my_dv;
    pdl_wait();
    const int my_v = D.verticesOnCell[(unsigned)i * D.maxEdges + le];
    const real my_kite = D.kiteAreasOnVertex[3 * my_v + D.kiteForCell[(unsigned)i * D.maxEdges + le]];
    r2 div = mk2(0.0, 0.0), ke = mk2(0.0, 0.0);
#define DIAG_C_EDGE(E)                                                                                      \
    {                                                                                                       \
        const r2 uu = LD(u, BC(my_e, (E)));                                                                 \
        div = selb((E) < ne, div + BC(my_s, (E)) * uu, div);                                                \
        ke = selb((E) < ne, ke + 0.25 * (BC(my_efac, (E)) * (uu * uu)), ke);                                \
    }
#pragma unroll
    for (int e = 0; e < CW_NE; e++) DIAG_C_EDGE(e)
    for (int e = CW_NE; e < ne; e++) DIAG_C_EDGE(e)
#undef DIAG_C_EDGE
    ke = ke * r;
    const real ke_fact = 1.0 - .375;
    ke = ke_fact * ke;
    r2 pvc = mk2(0.0, 0.0);
#define DIAG_C_VTX(E)                                                                                       \
    {                                                                                                       \
        const int iVertex = BC(my_v, (E));                                                                  \
        const real kite = BC(my_kite, (E));                                                                 \
        const r2 kev = LD(D.ke_vertex, iVertex);                                                            \
        ke = selb((E) < ne, ke + (1. - ke_fact) * kite * kev * r, ke);                                      \
        if (apvm) pvc = selb((E) < ne, pvc + kite * LD(D.pv_vertex, iVertex) * r, pvc);                     \
    }
#pragma unroll
    for (int e = 0; e < CW_NE; e++) DIAG_C_VTX(e)
    for (int e = CW_NE; e < ne; e++) DIAG_C_VTX(e)
#undef DIAG_C_VTX
    const b2 k_lt_nl = lv.lt(nl);
    ST(D.divergence, i, sel(k_lt_nl, div * r, 0.0));
    ST(D.ke, i, sel(k_lt_nl, ke, 0.0));
    if (apvm) ST(D.pv_cell, i, sel(k_lt_nl, pvc, 0.0));
}
// (3) edge-all: h_edge (6428-6435), tangential velocity v (6618-6632, rk 3 only), pv_edge with APVM upwinding (6673-6745)
__global__ void __launch_bounds__(CW_THREADS, MB_DIAG_E) k2_diag_edge(const Dev D, const real* u, const real* h,
                                                           int reconstruct_v, int apvm, real apvm_dt) {
    CW_ENTER_R(D.nEdges)
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    const int vertex1 = D.verticesOnEdge[2 * i], vertex2 = D.verticesOnEdge[2 * i + 1];
    const b2 k_lt_nl = lv.lt(nl);
    pdl_wait();
    const r2 rho_edge = 0.5 * (LD(h, cell1) + LD(h, cell2));
    const r2 pv1 = LD(D.pv_vertex, vertex1), pv2 = LD(D.pv_vertex, vertex2);
    r2 vv;
    if (reconstruct_v) {
        vv = mk2(0.0, 0.0);
        const int neoe = D.nEdgesOnEdge[i];
        int my_eoe = 0; real my_woe = 0.0;
        if (lane < neoe) { my_eoe = D.edgesOnEdge[(unsigned)i * D.maxEdges2 + lane]; my_woe = D.weightsOnEdge[(unsigned)i * D.maxEdges2 + lane]; }
#pragma unroll 5
        for (int j = 0; j < neoe; j++) vv = vv + BC(my_woe, j) * LD(u, BC(my_eoe, j));
    } else {
        vv = LD(D.v, i);
    }
    r2 pve = 0.5 * (pv1 + pv2);
    r2 gt = mk2(0.0, 0.0), gn = mk2(0.0, 0.0);
    if (apvm) {
        const real r1 = 1.0 * D.invDvEdge[i];
        const real r2_ = 1.0 * D.invDcEdge[i];
        gt = (pv2 - pv1) * r1;
        gn = (LD(D.pv_cell, cell2) - LD(D.pv_cell, cell1)) * r2_;
        pve = pve - apvm_dt * (vv * gt + LD(u, i) * gn);
    }
    ST(D.rho_edge, i, sel(k_lt_nl, rho_edge, 0.0));
    if (reconstruct_v) ST(D.v, i, sel(k_lt_nl, vv, 0.0));
    if (apvm) { ST(D.gradPVt, i, sel(k_lt_nl, gt, 0.0)); ST(D.gradPVn, i, sel(k_lt_nl, gn, 0.0)); }
    ST(D.pv_edge, i, sel(k_lt_nl, pve, 0.0));
}

// ------------------------------------------------------------------ atm_compute_dyn_tend_work, part (a)
// cell-all: Smagorinsky kdiff (rk 1, TI:5226-5296), h_divergence (5307-5338), tend_rho + dpdz (rk 1, 5345-5362).
// Restriction (host falls back to k_dt_cell_a otherwise): config_mpas_cam_coef == 0.
__global__ void __launch_bounds__(CW_THREADS, MB_CELL_A) k2_dt_cell_a(const Dev D, const DynTendArgs A) {
    CW_ENTER(D.nCells)
    const int ne = D.nEdgesOnCell[i];
    const int le = min(lane, ne - 1);
    const int my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    const real my_es = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le] * D.dvEdge[my_e];
    const b2 k_lt_nl = lv.lt(nl);
    pdl_wait();
    r2 kd = mk2(A.fixed_visc2, A.fixed_visc2);
    if (A.rk_step == 1 && A.smag) {
        const real my_da = D.defc_a[(unsigned)i * D.maxEdges + le], my_db = D.defc_b[(unsigned)i * D.maxEdges + le];
        r2 d_diag = mk2(0.0, 0.0), d_off_diag = mk2(0.0, 0.0);
#define CELL_A_DEF(E)                                                                                       \
        {                                                                                                   \
            const int iEdge = BC(my_e, (E));                                                                \
            const real da = BC(my_da, (E)), db = BC(my_db, (E));                                            \
            const r2 uu = LD(D.u_2, iEdge), vv = LD(D.v, iEdge);                                            \
            d_diag = selb((E) < ne, d_diag + da * uu - db * vv, d_diag);                                    \
            d_off_diag = selb((E) < ne, d_off_diag + db * uu + da * vv, d_off_diag);                        \
        }
#pragma unroll
        for (int e = 0; e < CW_NE; e++) CELL_A_DEF(e)
        for (int e = CW_NE; e < ne; e++) CELL_A_DEF(e)
#undef CELL_A_DEF
        const r2 dd = d_diag * d_diag + d_off_diag * d_off_diag;
        kd = mk2(rmin(A.cs_len2 * sqrt(dd.x), A.kdiff_cap), rmin(A.cs_len2 * sqrt(dd.y), A.kdiff_cap));
    }
    r2 hd = mk2(0.0, 0.0);
#define CELL_A_DIV(E) { const r2 ru = LD(D.ru, BC(my_e, (E))); hd = selb((E) < ne, hd + BC(my_es, (E)) * ru, hd); }
#pragma unroll
    for (int e = 0; e < CW_NE; e++) CELL_A_DIV(e)
    for (int e = CW_NE; e < ne; e++) CELL_A_DIV(e)
#undef CELL_A_DIV
    hd = hd * D.invAreaCell[i];
    if (A.rk_step == 1) {
        const r2 rw = LD(D.rw, i), qt = LD(D.qtot, i);
        const r2 tend_rho = -hd - LD(D.rdzw, 0) * (dn1(rw) - rw) + LD(D.tend_rho_physics, i);
        const r2 dpdz = -GRAVITY * (LD(D.rho_base, i) * (qt) + LD(D.rho_p_save, i) * (1. + qt));
        ST(D.kdiff, i, sel(k_lt_nl, kd, 0.0));
        ST(D.tend_rho, i, sel(k_lt_nl, tend_rho, 0.0));
        ST(D.dpdz, i, sel(k_lt_nl, dpdz, 0.0));
    }
    ST(D.h_divergence, i, sel(k_lt_nl, hd, 0.0));
}

// ------------------------------------------------------------------ atm_compute_dyn_tend_work, part (e)
// rk 1, cell-all: first del^2 of w (5795-5829) and of theta_m (6027-6057) with their 2nd-order mixing tendencies
__global__ void __launch_bounds__(CW_THREADS, MB_CELL_E) k2_dt_cell_e(const Dev D, const DynTendArgs A) {
    CW_ENTER_R(D.nCells)
    const int ne = D.nEdgesOnCell[i];
    const int le = min(lane, ne - 1);
    const int my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    const real my_sgn = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le];
    const int my_c1 = D.cellsOnEdge[2 * my_e], my_c2 = D.cellsOnEdge[2 * my_e + 1];
    pdl_wait();
    // one of the two cells of every edge is this cell: its columns are loaded once, only the other cell is gathered
    const bool my_is1 = my_c1 == i, my_is2 = my_c2 == i;
    const int my_oth = my_is1 ? my_c2 : my_c1;
    const real my_dv = D.dvEdge[my_e], my_idc = D.invDcEdge[my_e], my_msd2 = D.meshScalingDel2[my_e];
    const real r_areaCell = D.invAreaCell[i];
    const r2 kd_own = LD(D.kdiff, i), w_own = LD(D.w_2, i), th_own = LD(D.theta_m_2, i);
    r2 dsw = mk2(0.0, 0.0), twe = mk2(0.0, 0.0), dst = mk2(0.0, 0.0), tte = mk2(0.0, 0.0);
#define CELL_E_EDGE(E)                                                                                      \
    {                                                                                                       \
        const int iEdge = BC(my_e, (E)), oth = BC(my_oth, (E));                                             \
        const bool is1 = BC(my_is1, (E)), is2 = BC(my_is2, (E));                                            \
        const real sg = BC(my_sgn, (E)), dv = BC(my_dv, (E)), idc = BC(my_idc, (E)), msd2 = BC(my_msd2, (E)); \
        const r2 rho_e = LD(D.rho_edge, iEdge);                                                             \
        const r2 kd_o = LD(D.kdiff, oth), w_o = LD(D.w_2, oth), th_o = LD(D.theta_m_2, oth);                \
        const r2 kd1 = selb(is1, kd_own, kd_o), kd2 = selb(is2, kd_own, kd_o);                              \
        const r2 dw = selb(is2, w_own, w_o) - selb(is1, w_own, w_o);                                        \
        const r2 dth = selb(is2, th_own, th_o) - selb(is1, th_own, th_o);                                   \
        {                                                                                                   \
            const real edge_sign = 0.5 * r_areaCell * sg * dv * idc;                                        \
            r2 w_turb_flux = edge_sign * (rho_e + up1(rho_e)) * dw;                                         \
            dsw = selb((E) < ne, dsw + w_turb_flux, dsw);                                                   \
            w_turb_flux = w_turb_flux * msd2 * 0.25 *                                                       \
                          (kd1 + kd2 + up1(kd1) + up1(kd2));                                                \
            twe = selb((E) < ne, twe + w_turb_flux, twe);                                                   \
        }                                                                                                   \
        {                                                                                                   \
            const real edge_sign = r_areaCell * sg * dv * idc;                                              \
            const real pr_scale = A.prandtl_inv * msd2;                                                     \
            r2 theta_turb_flux = edge_sign * dth * rho_e;                                                   \
            dst = selb((E) < ne, dst + theta_turb_flux, dst);                                               \
            theta_turb_flux = theta_turb_flux * 0.5 * (kd1 + kd2) * pr_scale;                               \
            tte = selb((E) < ne, tte + theta_turb_flux, tte);                                               \
        }                                                                                                   \
    }
#pragma unroll 3
    for (int e = 0; e < CW_NE; e++) CELL_E_EDGE(e)
    for (int e = CW_NE; e < ne; e++) CELL_E_EDGE(e)
#undef CELL_E_EDGE
    const b2 k_lt_nl = lv.lt(nl), k_mid = lv.ge(1) && lv.lt(nl);
    ST(D.delsq_w, i, sel(k_mid, dsw, 0.0));
    ST(D.tend_w_euler, i, sel(k_mid, twe, 0.0));
    ST(D.delsq_theta, i, sel(k_lt_nl, dst, 0.0));
    ST(D.tend_theta_euler, i, sel(k_lt_nl, tte, 0.0));
}

// ------------------------------------------------------------------ atm_advance_acoustic_step_work, cell part (block-tiled)
// TI:2824-2973.  A block owns AC3_COLS consecutive cells.  Phase 1: each warp assembles the right-hand sides of
// AC3_COLS/CW_WARPS columns (one after the other) into shared memory.  Phase 2: ONE warp runs the tridiagonal
// sweeps (TI:2922-2930) of all AC3_COLS columns at once, lane = column, rows padded to an odd stride so that the
// lanes hit distinct banks -- the serial recurrence costs one warp-instruction per level for 32 columns instead of
// one per column.  Phase 3: each warp finishes its columns (Rayleigh damping, wwAvg, rho_pp, rtheta_pp).
// Operation order inside every column is the reference's: bit-identical results.
// Tiling measured on B200 (x1.40962 x 55): 32 columns / 8 warps / 2 blocks per SM (128 registers) 154 us; 16 columns / 4 warps with
// 4 blocks (128 registers) 150 us, 3 blocks (167 registers: more loads in flight per warp) 144 us, 2 blocks (207 registers) 176 us.
#ifndef AC3_COLS
#define AC3_COLS 16                     // columns per block = lanes of the sweeping warp
#endif
#ifndef AC3_MINB
#define AC3_MINB 3
#endif
#define AC3_ARRAYS 6                    // rw (rhs / solution), a_tri, alpha_tri, gamma_tri, ts, rs
#ifndef AC3_WARPS
#define AC3_WARPS 4                     // warps per block: AC3_COLS / AC3_WARPS columns per warp
#endif
__global__ void __launch_bounds__(AC3_WARPS * 32, AC3_MINB) k3_acoustic_cell(const Dev D, real dts, int small_step, real epssm, real resm) {
    extern __shared__ __align__(16) real sm3[];
    PDL_ENTER
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int LDK = D.LDK, nl = D.nl;
    const int S = LDK | 1;                                  // odd row stride (LDK is even)
    real* s_rw = sm3;
    real* s_a = s_rw + AC3_COLS * S;
    real* s_al = s_a + AC3_COLS * S;
    real* s_ga = s_al + AC3_COLS * S;
    real* s_ts = s_ga + AC3_COLS * S;
    real* s_rs = s_ts + AC3_COLS * S;
    Lv lv; lv.k0 = 2 * lane;
    const int k0 = lv.k0; const bool act = k0 < D.LDKA;
    const unsigned uLDK = (unsigned)LDK, kc = (unsigned)min(k0, D.LDKA - 2);
    const bool first = small_step == 1;
    const b2 k_lt_nl = lv.lt(nl), k_le_nl = lv.lt(nl + 1), k_mid = lv.ge(1) && lv.lt(nl);
    const int base = (gridDim.x - 1 - blockIdx.x) * AC3_COLS;      // backward sweep (see CW_SETUP_R)
    const r2 rdzw = LD(D.rdzw, 0), cofrz = LD(D.cofrz, 0);
    // connectivity of ALL columns of this warp in one pass: lane = (column slot << 3) | edge slot, so the two
    // dependent index loads (edgesOnCell -> cellsOnEdge/dvEdge) are exposed once per warp, not once per column
    static_assert(AC3_COLS / AC3_WARPS <= 4 && CW_MAXNE <= 8, "one lane per (column, edge slot)");
    int m_ne = 1, m_e = 0, m_oth = 0; bool m_is1 = false, m_is2 = false; real m_f = 0.0, m_invArea = 0.0;
    {
        const int mi = base + (lane >> 3) * AC3_WARPS + wib;
        if ((lane >> 3) < AC3_COLS / AC3_WARPS && mi < D.nCellsSolve) {
            m_ne = D.nEdgesOnCell[mi];
            m_invArea = D.invAreaCell[mi];
            const int le = min(lane & 7, m_ne - 1);
            m_e = D.edgesOnCell[(unsigned)mi * D.maxEdges + le];
            // one of the two cells of every edge is the column itself: only the other cell's theta_m is gathered
            const int c1 = D.cellsOnEdge[2 * m_e], c2 = D.cellsOnEdge[2 * m_e + 1];
            m_is1 = c1 == mi; m_is2 = c2 == mi; m_oth = m_is1 ? c2 : c1;
            m_f = D.edgesOnCell_sign[(unsigned)mi * D.maxEdges + le] * dts * D.dvEdge[m_e];
        }
    }
    // ---------------- phase 1: right-hand sides
    for (int cc = 0; cc < AC3_COLS / AC3_WARPS; cc++) {
        const int c = cc * AC3_WARPS + wib;
        const int i = base + c;
        if (i >= D.nCellsSolve) continue;                   // warp-uniform
        const int ne = BC(m_ne, cc << 3);
        const real invArea = BC(m_invArea, cc << 3);
        // operands of phase 3 (after the solve): start them towards L2 now
        PF(D.dss, i); PF(D.rw_save, i); PF(D.rw, i); PF(D.rho_zz_2, i); PF(D.w_2, i);
        if (!first) PF(D.wwAvg, i);
        r2 rtheta_pp = mk2(0.0, 0.0), rho_pp = mk2(0.0, 0.0), rw_p = mk2(0.0, 0.0);
        if (!first) {
            rtheta_pp = sel(k_lt_nl, LD(D.rtheta_pp, i), 0.0);
            rw_p = sel(k_le_nl, LD(D.rw_p, i), 0.0);
            rho_pp = sel(k_lt_nl, LD(D.rho_pp, i), 0.0);
        }
        const r2 tend_rho = LD(D.tend_rho, i), tend_theta = LD(D.tend_theta, i), tend_w = LD(D.tend_w, i);
        const r2 coftz = LD(D.coftz, i), cofwz = LD(D.cofwz, i), cofwr = LD(D.cofwr, i), cofwt = LD(D.cofwt, i);
        const r2 zz = LD(D.zz, i);
        const r2 a_tri = LD(D.a_tri, i), al_tri = LD(D.alpha_tri, i), ga_tri = LD(D.gamma_tri, i);
        const r2 th_own = LD(D.theta_m, i);
        r2 rs = mk2(0.0, 0.0), ts = mk2(0.0, 0.0);
#define AC_EDGE(E)                                                                                          \
        {                                                                                                   \
            const int src = (cc << 3) + (E);                                                                \
            const int iEdge = BC(m_e, src);                                                                 \
            const bool is1 = BC(m_is1, src), is2 = BC(m_is2, src);                                          \
            const r2 th_o = LD(D.theta_m, BC(m_oth, src));                                                  \
            const r2 ru_p = first ? dts * LD(D.tend_u, iEdge) : LD(D.ru_p, iEdge);   /* TI:2798-2806 */     \
            const r2 flux = BC(m_f, src) * ru_p * invArea;                                                  \
            const r2 th = selb(is2, th_own, th_o) + selb(is1, th_own, th_o);                                \
            rs = selb((E) < ne, rs - flux, rs);                                                             \
            ts = selb((E) < ne, ts - flux * 0.5 * th, ts);                                                  \
        }
#pragma unroll
        for (int e = 0; e < CW_NE; e++) AC_EDGE(e)
        for (int e = CW_NE; e < ne; e++) AC_EDGE(e)
#undef AC_EDGE
        const r2 rw_p1 = dn1(rw_p);
        const r2 coftz1 = dn1(coftz);
        rs = rho_pp + dts * tend_rho + rs
             - cofrz * resm * (rw_p1 - rw_p);
        ts = rtheta_pp + dts * tend_theta + ts
             - resm * rdzw * (coftz1 * rw_p1
                              - coftz * rw_p);
        rs = sel(k_lt_nl, rs, 0.0); ts = sel(k_lt_nl, ts, 0.0);
        const r2 zzm = up1(zz);
        const r2 tsm = up1(ts), rsm = up1(rs), rtm = up1(rtheta_pp), rhm = up1(rho_pp), cofwtm = up1(cofwt);
        const r2 r = rw_p + dts * tend_w
                     - cofwz * ((zz * ts
                                 - zzm * tsm)
                                + resm * (zz * rtheta_pp
                                          - zzm * rtm))
                     - cofwr * ((rs + rsm)
                                + resm * (rho_pp + rhm))
                     + cofwt * (ts + resm * rtheta_pp)
                     + cofwtm * (tsm + resm * rtm);
        const r2 rhs = sel(k_mid, r, rw_p);
        if (act) {
            const int o = c * S + k0;
            s_rw[o] = rhs.x; s_rw[o + 1] = rhs.y; s_a[o] = a_tri.x; s_a[o + 1] = a_tri.y;
            s_al[o] = al_tri.x; s_al[o + 1] = al_tri.y; s_ga[o] = ga_tri.x; s_ga[o + 1] = ga_tri.y;
            s_ts[o] = ts.x; s_ts[o + 1] = ts.y; s_rs[o] = rs.x; s_rs[o + 1] = rs.y;
        }
    }
    __syncthreads();
    // ---------------- phase 2: all columns of the block swept by one warp, lane = column
    if (wib == 0 && lane < AC3_COLS && base + lane < D.nCellsSolve) {
        real* rwv = s_rw + lane * S;
        const real* av = s_a + lane * S; const real* alv = s_al + lane * S; const real* gav = s_ga + lane * S;
        real prev = rwv[0];
#pragma unroll 4
        for (int kk = 1; kk < nl; kk++) {
            prev = (rwv[kk] - av[kk] * prev) * alv[kk];
            rwv[kk] = prev;
        }
        real next = rwv[nl];
#pragma unroll 4
        for (int kk = nl - 1; kk >= 0; kk--) {
            next = rwv[kk] - gav[kk] * next;
            rwv[kk] = next;
        }
    }
    __syncthreads();
    // ---------------- phase 3: damping, averages, back-substitution of rho_pp and rtheta_pp
    const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0);
    for (int cc = 0; cc < AC3_COLS / AC3_WARPS; cc++) {
        const int c = cc * AC3_WARPS + wib;
        const int i = base + c;
        if (i >= D.nCells) continue;
        // old values of the perturbation variables (zero on the first small step, TI:2850-2860)
        r2 rtheta_pp_old = mk2(0.0, 0.0);
        if (!first) rtheta_pp_old = sel(k_lt_nl, LD(D.rtheta_pp, i), 0.0);
        if (i >= D.nCellsSolve) { ST(D.rtheta_pp_old, i, rtheta_pp_old); continue; }
        r2 rw_p = mk2(0.0, 0.0), wwAvg = mk2(0.0, 0.0);
        if (!first) {
            rw_p = sel(k_le_nl, LD(D.rw_p, i), 0.0);
            wwAvg = sel(k_le_nl, LD(D.wwAvg, i), 0.0);
        }
        const r2 coftz = LD(D.coftz, i), zz = LD(D.zz, i);
        const r2 dss = LD(D.dss, i), rw_save = LD(D.rw_save, i), rw_now = LD(D.rw, i), rho = LD(D.rho_zz_2, i), w_now = LD(D.w_2, i);
        const int o = c * S + (int)kc;
        r2 r = mk2(s_rw[o], s_rw[o + 1]);
        const r2 ts = mk2(s_ts[o], s_ts[o + 1]), rs = mk2(s_rs[o], s_rs[o + 1]);
        wwAvg = sel(k_mid, wwAvg + 0.5 * (1.0 - epssm) * rw_p, wwAvg);
        const r2 zzm = up1(zz);
        const r2 coftz1 = dn1(coftz);
        {   // implicit Rayleigh damping on w, TI:2936-2942
            const r2 dw = rw_save - rw_now;
            const r2 rd = (r + dw - dts * dss *
                           (fm * zz + fp * zzm)
                           * (fm * rho + fp * up1(rho))
                           * w_now) / (1.0 + dts * dss)
                          - dw;
            r = sel(k_mid, rd, r);
            wwAvg = sel(k_mid, wwAvg + 0.5 * (1.0 + epssm) * r, wwAvg);
        }
        r = sel(k_le_nl, r, 0.0);
        const r2 r1 = dn1(r);
        ST(D.rtheta_pp_old, i, rtheta_pp_old);
        ST(D.rw_p, i, r);
        ST(D.wwAvg, i, sel(k_le_nl, wwAvg, 0.0));
        ST(D.rho_pp, i, sel(k_lt_nl, rs - cofrz * (r1 - r), 0.0));
        ST(D.rtheta_pp, i, sel(k_lt_nl, ts - rdzw * (coftz1 * r1
                                                      - coftz * r), 0.0));
    }
}

// ---- the same routine with the column solve done by the warp that owns the column (relaxed arithmetic) ----
// The two sweeps of TI:2922-2930 are first-order linear recurrences, x_k = A_k x_{k-1} + B_k with A_k = -a_k alpha_k,
// B_k = rhs_k alpha_k, and y_k = G_k y_{k+1} + x_k with G_k = -gamma_k.  Affine maps compose associatively, so each is a
// parallel prefix over the levels: every lane composes its own level pair, five shuffle steps combine the 32 lanes, one more
// shuffle hands each lane its neighbour's end value.  One warp then carries a column from its gathers to its stores with
// everything in registers: no shared-memory tiles, no block-wide barriers between the phases, no operand read twice
// (k3_acoustic_cell re-reads coftz, zz, rw_p, rtheta_pp after the solve), and occupancy is set by registers alone.
// |A_k| < 0.7 and |G_k| < 1.1 on the JW cases (the system is diagonally dominant), so the products the prefixes form stay of
// order one and the re-associated solve agrees with the serial sweep to rounding (tests: TOL_ROUTINE_FAST per routine, 1e-11
// per step; tests/test_oracle_cpu.py::test_scan_form_of_the_column_solve repeats the re-association in numpy: 7e-17).
#ifndef AC6_MINB
#define AC6_MINB 2
#endif
struct aff { real a, b; };          // x -> a * x + b
__device__ __forceinline__ aff aff_after(aff later, aff earlier) { aff r; r.a = later.a * earlier.a; r.b = fma(later.a, earlier.b, later.b); return r; }
struct Ac6Conn { int ne, e, oth, is12; real f, invArea; };
// connectivity of cell i for its lane: edge slot min(lane, ne-1) -- two dependent index loads (edgesOnCell -> cellsOnEdge, dvEdge)
__device__ __forceinline__ Ac6Conn ac6_conn(const Dev& D, int i, int lane, real dts) {
    Ac6Conn c;
    c.ne = D.nEdgesOnCell[i];
    const int le = min(lane, c.ne - 1);
    c.e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    const int c1 = D.cellsOnEdge[2 * c.e], c2 = D.cellsOnEdge[2 * c.e + 1];
    c.is12 = (c1 == i ? 1 : 0) | (c2 == i ? 2 : 0);
    c.oth = (c1 == i) ? c2 : c1;
    c.f = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le] * dts * D.dvEdge[c.e];
    c.invArea = D.invAreaCell[i];
    return c;
}
// LISTED: the kernel works through the nlist columns of `list` instead of all cells -- a decomposed block may run it twice per
// small step, first over the columns a neighbour rank needs (and its own halo columns), then, while their exchange is on its
// way, over the rest (MPASB_SPLIT=1).  REGIONAL: with the specified-zone branch of TI:2862.  Both are template parameters
// because either costs the plain kernel its last free registers (168, 24 bytes of spills otherwise).
template <int WARPS, int MINB, bool LISTED, bool REGIONAL>
__global__ void __launch_bounds__(WARPS * 32, MINB) k6_acoustic_cell(const Dev D, real dts, int small_step, real epssm, real resm,
                                                                     const int* list, int nlist) {
    pdl_trigger();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int LDK = D.LDK, nl = D.nl;
    Lv lv; lv.k0 = 2 * lane;
    const int k0 = lv.k0; const bool act = k0 < D.LDKA;
    const unsigned uLDK = (unsigned)LDK, kc = (unsigned)min(k0, D.LDKA - 2);
    const bool first = small_step == 1;
    const b2 k_lt_nl = lv.lt(nl), k_le_nl = lv.lt(nl + 1), k_mid = lv.ge(1) && lv.lt(nl);
    const r2 rdzw = LD(D.rdzw, 0);
    const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0);
    // persistent warps (warp g: cells g, g + G, ...): the connectivity of the NEXT cell -- a chain of two dependent index
    // loads -- is fetched while this cell's columns are in flight, so each cell exposes one memory round trip, not three
    const int nTot = LISTED ? nlist : D.nCells;
    const Pers ps = pers_init<WARPS, MINB>(nTot, wib);
    const int G = ps.stride, nAll = ps.end;
#define AC6_POS(j) (nTot - 1 - (j))                           /* backward sweep (see CW_SETUP_R) */
#define AC6_COL(j) (LISTED ? list[AC6_POS(j)] : AC6_POS(j))
    int jcol = ps.j;
    Ac6Conn cn; cn.ne = 1; cn.e = 0; cn.oth = 0; cn.is12 = 0; cn.f = 0.0; cn.invArea = 0.0;
    if (jcol < nAll && AC6_COL(jcol) < D.nCellsSolve) cn = ac6_conn(D, AC6_COL(jcol), lane, dts);
    pdl_wait();                                          // everything above is static mesh data
    const r2 cofrz = LD(D.cofrz, 0);                     // (written by the vertical-coefficient kernel)
    for (; jcol < nAll; jcol += G) {
    const int i = AC6_COL(jcol);
    const int inext = jcol + G < nAll ? AC6_COL(jcol + G) : D.nCells; // the next column of this warp (nCells: none)
    r2 rtheta_pp = mk2(0.0, 0.0), rho_pp = mk2(0.0, 0.0), rw_p = mk2(0.0, 0.0), wwAvg = mk2(0.0, 0.0);
    if (!first) rtheta_pp = sel(k_lt_nl, LD(D.rtheta_pp, i), 0.0);
    if (i >= D.nCellsSolve) {                                                          // halo cells: TI:2827-2842 only
        ST(D.rtheta_pp_old, i, rtheta_pp);
        if (inext < D.nCellsSolve) cn = ac6_conn(D, inext, lane, dts);                 // (backward sweep: halo cells come first)
        continue;
    }
    const int ne = cn.ne;
    const int my_e = cn.e, my_oth = cn.oth;
    const bool my_is1 = (cn.is12 & 1) != 0, my_is2 = (cn.is12 & 2) != 0;
    const real my_f = cn.f, invArea = cn.invArea;
    if (!first) {
        rw_p = sel(k_le_nl, LD(D.rw_p, i), 0.0);
        rho_pp = sel(k_lt_nl, LD(D.rho_pp, i), 0.0);
        wwAvg = sel(k_le_nl, LD(D.wwAvg, i), 0.0);
    }
    const r2 tend_rho = LD(D.tend_rho, i), tend_theta = LD(D.tend_theta, i), tend_w = LD(D.tend_w, i);
    if (REGIONAL && D.specZoneMaskCell[i] != 0.0) {
        // specified zone of a regional run (TI:2862, 2962-2971): no implicit solve, the perturbation variables follow their
        // (driving) tendencies; row nl is zeroed on the first small step and left alone afterwards
        const r2 rw_new = rw_p + dts * tend_w;
        ST(D.rtheta_pp_old, i, rtheta_pp);
        ST(D.rho_pp, i, sel(k_lt_nl, rho_pp + dts * tend_rho, 0.0));
        ST(D.rtheta_pp, i, sel(k_lt_nl, rtheta_pp + dts * tend_theta, 0.0));
        ST(D.rw_p, i, sel(k_lt_nl, rw_new, sel(k_le_nl, rw_p, 0.0)));
        ST(D.wwAvg, i, sel(k_lt_nl, wwAvg + 0.5 * (1.0 + epssm) * rw_new, sel(k_le_nl, wwAvg, 0.0)));
        if (inext < D.nCellsSolve) cn = ac6_conn(D, inext, lane, dts);
        continue;
    }
    const r2 coftz = LD(D.coftz, i), cofwz = LD(D.cofwz, i), cofwr = LD(D.cofwr, i), cofwt = LD(D.cofwt, i);
    const r2 zz = LD(D.zz, i);
    const r2 a_tri = LD(D.a_tri, i), al_tri = LD(D.alpha_tri, i), ga_tri = LD(D.gamma_tri, i);
    const r2 th_own = LD(D.theta_m, i);
    r2 rs = mk2(0.0, 0.0), ts = mk2(0.0, 0.0);
#define AC6_EDGE(E)                                                                                         \
    {                                                                                                       \
        const int iEdge = BC(my_e, (E));                                                                    \
        const bool is1 = BC(my_is1, (E)), is2 = BC(my_is2, (E));                                            \
        const r2 th_o = LD(D.theta_m, BC(my_oth, (E)));                                                     \
        const r2 ru_p = first ? dts * LD(D.tend_u, iEdge) : LD(D.ru_p, iEdge);   /* TI:2798-2806 */         \
        const r2 flux = BC(my_f, (E)) * ru_p * invArea;                                                     \
        const r2 th = selb(is2, th_own, th_o) + selb(is1, th_own, th_o);                                    \
        rs = selb((E) < ne, rs - flux, rs);                                                                 \
        ts = selb((E) < ne, ts - flux * 0.5 * th, ts);                                                      \
    }
#pragma unroll
    for (int e = 0; e < CW_NE; e++) AC6_EDGE(e)
    for (int e = CW_NE; e < ne; e++) AC6_EDGE(e)
#undef AC6_EDGE
    // operands of the part after the solve: issued here so that they are in flight during the solve
    const r2 dss = LD(D.dss, i), rw_save = LD(D.rw_save, i), rw_now = LD(D.rw, i), rho = LD(D.rho_zz_2, i), w_now = LD(D.w_2, i);
    if (inext < D.nCellsSolve) {
        cn = ac6_conn(D, inext, lane, dts);          // next cell's connectivity, in flight with this cell's columns
        if (D.pf_next) {                             // and its own-column operands on their way into L2
            const int j = inext;
            PF(D.tend_rho, j); PF(D.tend_theta, j); PF(D.tend_w, j); PF(D.coftz, j); PF(D.cofwz, j); PF(D.cofwr, j); PF(D.cofwt, j);
            PF(D.zz, j); PF(D.a_tri, j); PF(D.alpha_tri, j); PF(D.gamma_tri, j); PF(D.theta_m, j); PF(D.dss, j); PF(D.rw_save, j);
            PF(D.rw, j); PF(D.rho_zz_2, j); PF(D.w_2, j);
            if (!first) { PF(D.rtheta_pp, j); PF(D.rw_p, j); PF(D.rho_pp, j); PF(D.wwAvg, j); }
        }
    }
    const r2 rw_p1 = dn1(rw_p);
    const r2 coftz1 = dn1(coftz);
    rs = rho_pp + dts * tend_rho + rs - cofrz * resm * (rw_p1 - rw_p);
    ts = rtheta_pp + dts * tend_theta + ts - resm * rdzw * (coftz1 * rw_p1 - coftz * rw_p);
    rs = sel(k_lt_nl, rs, 0.0); ts = sel(k_lt_nl, ts, 0.0);
    const r2 zzm = up1(zz);
    const r2 tsm = up1(ts), rsm = up1(rs), rtm = up1(rtheta_pp), rhm = up1(rho_pp), cofwtm = up1(cofwt);
    const r2 rr = rw_p + dts * tend_w
                  - cofwz * ((zz * ts - zzm * tsm) + resm * (zz * rtheta_pp - zzm * rtm))
                  - cofwr * ((rs + rsm) + resm * (rho_pp + rhm))
                  + cofwt * (ts + resm * rtheta_pp)
                  + cofwtm * (tsm + resm * rtm);
    const r2 rhs = sel(k_le_nl, sel(k_mid, rr, rw_p), 0.0);
    // ---- forward sweep: x_k = alpha_k (rhs_k - a_k x_{k-1}), k = 1 .. nl-1; x_0 = rhs_0, x_nl = rhs_nl
    r2 x;
    {
        aff m0, m1;
        m0.a = k_mid.x ? -(a_tri.x * al_tri.x) : (real)0.0; m0.b = k_mid.x ? rhs.x * al_tri.x : rhs.x;
        m1.a = k_mid.y ? -(a_tri.y * al_tri.y) : (real)0.0; m1.b = k_mid.y ? rhs.y * al_tri.y : rhs.y;
        aff m = aff_after(m1, m0);                              // x_{k0-1} -> x_{k0+1}
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            aff p; p.a = __shfl_up_sync(CW_FULL, m.a, d); p.b = __shfl_up_sync(CW_FULL, m.b, d);
            if (lane >= d) m = aff_after(m, p);
        }
        real xprev = __shfl_up_sync(CW_FULL, m.b, 1);           // x_{k0-1}: the previous lane's upper level (lane 0: m0.a == 0)
        if (lane == 0) xprev = 0.0;
        x.x = fma(m0.a, xprev, m0.b);
        x.y = m.b;
    }
    // ---- backward sweep: y_k = x_k - gamma_k y_{k+1}, k = nl-1 .. 0; y_nl = x_nl
    r2 r;
    {
        aff m0, m1;
        m0.a = k_lt_nl.x ? -ga_tri.x : (real)0.0; m0.b = x.x;
        m1.a = k_lt_nl.y ? -ga_tri.y : (real)0.0; m1.b = x.y;
        aff m = aff_after(m0, m1);                              // y_{k0+2} -> y_{k0}
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            aff p; p.a = __shfl_down_sync(CW_FULL, m.a, d); p.b = __shfl_down_sync(CW_FULL, m.b, d);
            if (lane + d < 32) m = aff_after(m, p);
        }
        real ynext = __shfl_down_sync(CW_FULL, m.b, 1);         // y_{k0+2}
        if (lane == 31) ynext = 0.0;
        r.x = m.b;
        r.y = fma(m1.a, ynext, m1.b);
    }
    // ---- damping, averages, back-substitution of rho_pp and rtheta_pp (TI:2936-2959)
    wwAvg = sel(k_mid, wwAvg + 0.5 * (1.0 - epssm) * rw_p, wwAvg);
    {
        const r2 dw = rw_save - rw_now;
        const r2 rd = (r + dw - dts * dss * (fm * zz + fp * zzm) * (fm * rho + fp * up1(rho)) * w_now) / (1.0 + dts * dss) - dw;
        r = sel(k_mid, rd, r);
        wwAvg = sel(k_mid, wwAvg + 0.5 * (1.0 + epssm) * r, wwAvg);
    }
    r = sel(k_le_nl, r, 0.0);
    const r2 r1 = dn1(r);
    ST(D.rtheta_pp_old, i, rtheta_pp);
    ST(D.rw_p, i, r);
    ST(D.wwAvg, i, sel(k_le_nl, wwAvg, 0.0));
    ST(D.rho_pp, i, sel(k_lt_nl, rs - cofrz * (r1 - r), 0.0));
    ST(D.rtheta_pp, i, sel(k_lt_nl, ts - rdzw * (coftz1 * r1 - coftz * r), 0.0));
    }
#undef AC6_COL
#undef AC6_POS
}

// ---- the same column solve with its OWN-COLUMN operands staged by bulk-asynchronous copies (TMA) ----
// k6_acoustic_cell reads 21 of its 33 columns from the cell's own column; a block of W warps that takes W consecutive cells per
// trip therefore needs, per field, ONE contiguous slab of W columns (W * LDK reals).  Here warp 0 requests the slabs of the NEXT
// trip -- one cp.async.bulk (SASS UBLKCP) per field, each issued by its own lane, all completing on one mbarrier -- into the
// other half of a double buffer while the block works on the current trip, and the warps read their operands from shared
// memory at the point of use: a whole trip of prefetch distance (the kernel is bound by memory latency at 12 warps per SM, not
// by bytes), no registers held by loads in flight, no LSU issue slots for 20 of the 33 columns.  Gathers of neighbour columns
// (theta_m of the <= 6 neighbours, ru_p of the <= 6 edges) and the cell's own theta_m stay LDG.128.
// Protocol: full[s] (1 arrival + transaction bytes) is waited for by every warp before its first read of stage s; empty[s]
// (W arrivals, one per warp after its last read) is waited for by warp 0 before it refills stage s.  Trips are block-uniform.
// Arithmetic and results are those of k6_acoustic_cell.  Not for listed or regional runs (k6 serves those).
#ifndef AC9_MINB
#define AC9_MINB 3
#endif
#ifndef AC9_WARPS
#define AC9_WARPS 4
#endif
#ifndef AC9_DEFAULT
#define AC9_DEFAULT 0                   // MPASB_AC9=1 selects it at run time
#endif
enum { A9_rtheta_pp, A9_rw_p, A9_rho_pp, A9_wwAvg,        // the first four are not read on the first small step
       A9_tend_rho, A9_tend_theta, A9_tend_w, A9_coftz, A9_cofwz, A9_cofwr, A9_cofwt, A9_zz, A9_a_tri, A9_alpha_tri, A9_gamma_tri,
       A9_dss, A9_rw_save, A9_rw, A9_rho_zz_2, A9_w_2, AC9_NF };
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k9_acoustic_cell(const Dev D, real dts, int small_step, real epssm, real resm) {
    extern __shared__ __align__(128) unsigned char a9_raw[];          // [2 stages][AC9_NF fields][WARPS columns][LDK]
    __shared__ __align__(8) unsigned long long a9_bar[4];             // full[0], full[1], empty[0], empty[1]
    __shared__ const real* a9_src[AC9_NF];
    pdl_trigger();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int LDK = D.LDK, nl = D.nl;
    Lv lv; lv.k0 = 2 * lane;
    const int k0 = lv.k0; const bool act = k0 < D.LDKA;
    const unsigned uLDK = (unsigned)LDK, kc = (unsigned)min(k0, D.LDKA - 2);
    const bool first = small_step == 1;
    const b2 k_lt_nl = lv.lt(nl), k_le_nl = lv.lt(nl + 1), k_mid = lv.ge(1) && lv.lt(nl);
    const r2 rdzw = LD(D.rdzw, 0);
    const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0);
    const unsigned colB = uLDK * (unsigned)sizeof(real), fieldB = WARPS * colB, stageB = AC9_NF * fieldB;
    const unsigned bar = smem_u32(a9_bar), raw = smem_u32(a9_raw);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1); mbar_init(bar + 8, 1); mbar_init(bar + 16, WARPS); mbar_init(bar + 24, WARPS);
        a9_src[A9_rtheta_pp] = D.rtheta_pp; a9_src[A9_rw_p] = D.rw_p; a9_src[A9_rho_pp] = D.rho_pp; a9_src[A9_wwAvg] = D.wwAvg;
        a9_src[A9_tend_rho] = D.tend_rho; a9_src[A9_tend_theta] = D.tend_theta; a9_src[A9_tend_w] = D.tend_w;
        a9_src[A9_coftz] = D.coftz; a9_src[A9_cofwz] = D.cofwz; a9_src[A9_cofwr] = D.cofwr; a9_src[A9_cofwt] = D.cofwt;
        a9_src[A9_zz] = D.zz; a9_src[A9_a_tri] = D.a_tri; a9_src[A9_alpha_tri] = D.alpha_tri; a9_src[A9_gamma_tri] = D.gamma_tri;
        a9_src[A9_dss] = D.dss; a9_src[A9_rw_save] = D.rw_save; a9_src[A9_rw] = D.rw; a9_src[A9_rho_zz_2] = D.rho_zz_2; a9_src[A9_w_2] = D.w_2;
    }
    __syncthreads();
    // trip t of this block: group g = blockIdx + t * gridDim of W consecutive columns [lo, hi), walked from the last column down
    // (backward sweep, see CW_SETUP_R); warp wib takes column hi - 1 - wib
    const int nAll = D.nCells, nG = (nAll + WARPS - 1) / WARPS, NB = (int)gridDim.x;
    const int f0 = first ? 4 : 0;
    // request the slabs of group g into stage s (warp 0 only; lane f copies field f)
#define AC9_FILL(g, s)                                                                                           \
    {                                                                                                             \
        const int hi_ = nAll - (g) * WARPS, lo_ = max(0, hi_ - WARPS);                                            \
        const unsigned bytes_ = (unsigned)(hi_ - lo_) * colB;                                                     \
        if (lane == 0) mbar_expect_tx(bar + 8u * (s), (unsigned)(AC9_NF - f0) * bytes_);                          \
        __syncwarp();                                                                                             \
        if (lane >= f0 && lane < AC9_NF)                                                                          \
            bulk_g2s(raw + (s) * stageB + (unsigned)lane * fieldB, a9_src[lane] + (size_t)lo_ * uLDK, bytes_, bar + 8u * (s)); \
    }
    int g = (int)blockIdx.x;
    Ac6Conn cn; cn.ne = 1; cn.e = 0; cn.oth = 0; cn.is12 = 0; cn.f = 0.0; cn.invArea = 0.0;
    {
        const int i0 = nAll - g * WARPS - 1 - wib;
        if (g < nG && i0 >= 0 && i0 < D.nCellsSolve) cn = ac6_conn(D, i0, lane, dts);
    }
    pdl_wait();                                          // everything above is static mesh data
    const r2 cofrz = LD(D.cofrz, 0);                     // (written by the vertical-coefficient kernel)
    if (wib == 0 && g < nG) AC9_FILL(g, 0u)
    for (unsigned t = 0; g < nG; g += NB, t++) {
        const unsigned s = t & 1u, ph = (t >> 1) & 1u;
        const int hi = nAll - g * WARPS, lo = max(0, hi - WARPS);
        const int i = hi - 1 - wib;                       // this warp's column (i < lo: none, in the last group)
        const int gn = g + NB;
        const int inext = gn < nG ? nAll - gn * WARPS - 1 - wib : -1;
        if (wib == 0 && gn < nG) {                       // next trip's slabs into the other stage, once every warp has left it
            if (t >= 1) { if (lane == 0) mbar_wait(bar + 16 + 8u * (s ^ 1u), ((t - 1) >> 1) & 1u); __syncwarp(); }
            AC9_FILL(gn, s ^ 1u)
        }
        mbar_wait(bar + 8u * s, ph);
        const unsigned char* st = a9_raw + s * stageB + (unsigned)max(i - lo, 0) * colB + kc * (unsigned)sizeof(real);
#define S9(F) (*reinterpret_cast<const r2*>(st + (unsigned)(F) * fieldB))
#define AC9_RELEASE { __syncwarp(); if (lane == 0) mbar_arrive(bar + 16 + 8u * s); }
        if (i < lo) { AC9_RELEASE continue; }
        r2 rtheta_pp = mk2(0.0, 0.0), rho_pp = mk2(0.0, 0.0), rw_p = mk2(0.0, 0.0), wwAvg = mk2(0.0, 0.0);
        if (!first) rtheta_pp = sel(k_lt_nl, S9(A9_rtheta_pp), 0.0);
        if (i >= D.nCellsSolve) {                                                          // halo cells: TI:2827-2842 only
            AC9_RELEASE
            ST(D.rtheta_pp_old, i, rtheta_pp);
            if (inext >= 0 && inext < D.nCellsSolve) cn = ac6_conn(D, inext, lane, dts);
            continue;
        }
        const int ne = cn.ne;
        const int my_e = cn.e, my_oth = cn.oth;
        const bool my_is1 = (cn.is12 & 1) != 0, my_is2 = (cn.is12 & 2) != 0;
        const real my_f = cn.f, invArea = cn.invArea;
        const r2 th_own = LD(D.theta_m, i);
        r2 rs = mk2(0.0, 0.0), ts = mk2(0.0, 0.0);
#define AC9_EDGE(E)                                                                                         \
        {                                                                                                   \
            const int iEdge = BC(my_e, (E));                                                                \
            const bool is1 = BC(my_is1, (E)), is2 = BC(my_is2, (E));                                        \
            const r2 th_o = LD(D.theta_m, BC(my_oth, (E)));                                                 \
            const r2 ru_p = first ? dts * LD(D.tend_u, iEdge) : LD(D.ru_p, iEdge);   /* TI:2798-2806 */     \
            const r2 flux = BC(my_f, (E)) * ru_p * invArea;                                                 \
            const r2 th = selb(is2, th_own, th_o) + selb(is1, th_own, th_o);                                \
            rs = selb((E) < ne, rs - flux, rs);                                                             \
            ts = selb((E) < ne, ts - flux * 0.5 * th, ts);                                                  \
        }
#pragma unroll
        for (int e = 0; e < CW_NE; e++) AC9_EDGE(e)
        for (int e = CW_NE; e < ne; e++) AC9_EDGE(e)
#undef AC9_EDGE
        if (inext >= 0 && inext < D.nCellsSolve) cn = ac6_conn(D, inext, lane, dts);   // next cell's connectivity, in flight with the gathers
        if (!first) {
            rw_p = sel(k_le_nl, S9(A9_rw_p), 0.0);
            rho_pp = sel(k_lt_nl, S9(A9_rho_pp), 0.0);
            wwAvg = sel(k_le_nl, S9(A9_wwAvg), 0.0);
        }
        const r2 coftz = S9(A9_coftz), zz = S9(A9_zz);
        const r2 rw_p1 = dn1(rw_p);
        const r2 coftz1 = dn1(coftz);
        rs = rho_pp + dts * S9(A9_tend_rho) + rs - cofrz * resm * (rw_p1 - rw_p);
        ts = rtheta_pp + dts * S9(A9_tend_theta) + ts - resm * rdzw * (coftz1 * rw_p1 - coftz * rw_p);
        rs = sel(k_lt_nl, rs, 0.0); ts = sel(k_lt_nl, ts, 0.0);
        const r2 zzm = up1(zz);
        r2 rhs;
        {
            const r2 cofwt = S9(A9_cofwt);
            const r2 tsm = up1(ts), rsm = up1(rs), rtm = up1(rtheta_pp), rhm = up1(rho_pp), cofwtm = up1(cofwt);
            const r2 rr = rw_p + dts * S9(A9_tend_w)
                          - S9(A9_cofwz) * ((zz * ts - zzm * tsm) + resm * (zz * rtheta_pp - zzm * rtm))
                          - S9(A9_cofwr) * ((rs + rsm) + resm * (rho_pp + rhm))
                          + cofwt * (ts + resm * rtheta_pp)
                          + cofwtm * (tsm + resm * rtm);
            rhs = sel(k_le_nl, sel(k_mid, rr, rw_p), 0.0);
        }
        // ---- forward sweep: x_k = alpha_k (rhs_k - a_k x_{k-1}), k = 1 .. nl-1; x_0 = rhs_0, x_nl = rhs_nl
        r2 x;
        {
            const r2 a_tri = S9(A9_a_tri), al_tri = S9(A9_alpha_tri);
            aff m0, m1;
            m0.a = k_mid.x ? -(a_tri.x * al_tri.x) : (real)0.0; m0.b = k_mid.x ? rhs.x * al_tri.x : rhs.x;
            m1.a = k_mid.y ? -(a_tri.y * al_tri.y) : (real)0.0; m1.b = k_mid.y ? rhs.y * al_tri.y : rhs.y;
            aff m = aff_after(m1, m0);                              // x_{k0-1} -> x_{k0+1}
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                aff p; p.a = __shfl_up_sync(CW_FULL, m.a, d); p.b = __shfl_up_sync(CW_FULL, m.b, d);
                if (lane >= d) m = aff_after(m, p);
            }
            real xprev = __shfl_up_sync(CW_FULL, m.b, 1);           // x_{k0-1}: the previous lane's upper level (lane 0: m0.a == 0)
            if (lane == 0) xprev = 0.0;
            x.x = fma(m0.a, xprev, m0.b);
            x.y = m.b;
        }
        // ---- backward sweep: y_k = x_k - gamma_k y_{k+1}, k = nl-1 .. 0; y_nl = x_nl
        r2 r;
        {
            const r2 ga_tri = S9(A9_gamma_tri);
            aff m0, m1;
            m0.a = k_lt_nl.x ? -ga_tri.x : (real)0.0; m0.b = x.x;
            m1.a = k_lt_nl.y ? -ga_tri.y : (real)0.0; m1.b = x.y;
            aff m = aff_after(m0, m1);                              // y_{k0+2} -> y_{k0}
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                aff p; p.a = __shfl_down_sync(CW_FULL, m.a, d); p.b = __shfl_down_sync(CW_FULL, m.b, d);
                if (lane + d < 32) m = aff_after(m, p);
            }
            real ynext = __shfl_down_sync(CW_FULL, m.b, 1);         // y_{k0+2}
            if (lane == 31) ynext = 0.0;
            r.x = m.b;
            r.y = fma(m1.a, ynext, m1.b);
        }
        // ---- damping, averages, back-substitution of rho_pp and rtheta_pp (TI:2936-2959)
        wwAvg = sel(k_mid, wwAvg + 0.5 * (1.0 - epssm) * rw_p, wwAvg);
        {
            const r2 dss = S9(A9_dss), rho = S9(A9_rho_zz_2);
            const r2 dw = S9(A9_rw_save) - S9(A9_rw);
            const r2 rd = (r + dw - dts * dss * (fm * zz + fp * zzm) * (fm * rho + fp * up1(rho)) * S9(A9_w_2)) / (1.0 + dts * dss) - dw;
            r = sel(k_mid, rd, r);
            wwAvg = sel(k_mid, wwAvg + 0.5 * (1.0 + epssm) * r, wwAvg);
        }
        AC9_RELEASE                                      // last read of this stage
        r = sel(k_le_nl, r, 0.0);
        const r2 r1 = dn1(r);
        ST(D.rtheta_pp_old, i, rtheta_pp);
        ST(D.rw_p, i, r);
        ST(D.wwAvg, i, sel(k_le_nl, wwAvg, 0.0));
        ST(D.rho_pp, i, sel(k_lt_nl, rs - cofrz * (r1 - r), 0.0));
        ST(D.rtheta_pp, i, sel(k_lt_nl, ts - rdzw * (coftz1 * r1 - coftz * r), 0.0));
    }
#undef S9
#undef AC9_RELEASE
#undef AC9_FILL
}

// ------------------------------------------------------------------ atm_divergence_damping_3d  TI:2987-3075
// first != 0: also performs the first-small-step edge update of atm_advance_acoustic_step_work (TI:2798-2806:
// ru_p = dts * tend_u, ruAvg = ru_p), which k3_acoustic_cell only evaluated on the fly.
// list != nullptr: only the nlist edges of `list` (a decomposed block: the owned edges a neighbour rank receives; the damping
// of every other edge is folded into k2_recover_edge)
__global__ void __launch_bounds__(CW_THREADS) k2_divergence_damping(const Dev D, real coef_divdamp, int first, real dts,
                                                                    const int* list, int nlist) {
    pdl_trigger();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int ncols = list ? nlist : D.nEdges;
    const int i_fwd = blockIdx.x * CW_WARPS + wib;
    const int LDK = D.LDK, nl = D.nl;
    if (i_fwd >= ncols) return;
    const int i_pos = i_fwd;
    const int i = list ? list[i_pos] : i_pos;
    Lv lv; lv.k0 = 2 * lane;
    const int k0 = lv.k0; const bool act = k0 < D.LDKA; (void)nl;
    const unsigned uLDK = (unsigned)LDK, kc = (unsigned)min(k0, D.LDKA - 2);
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    pdl_wait();
    if (!(cell1 < D.nCellsSolve || cell2 < D.nCellsSolve)) return;
    const real mask = 1.0 - D.specZoneMaskEdge[i];
    const r2 divCell1 = -(LD(D.rtheta_pp, cell1) - LD(D.rtheta_pp_old, cell1));
    const r2 divCell2 = -(LD(D.rtheta_pp, cell2) - LD(D.rtheta_pp_old, cell2));
    const r2 th = LD(D.theta_m, cell1) + LD(D.theta_m, cell2);
    r2 ru_p;
    if (first) ru_p = dts * LD(D.tend_u, i);
    else ru_p = LD(D.ru_p, i);
    const b2 k_lt_nl = lv.lt(nl);
    if (first) ST(D.ruAvg, i, sel(k_lt_nl, ru_p, 0.0));
    ST(D.ru_p, i, sel(k_lt_nl, ru_p + coef_divdamp * (divCell2 - divCell1) * mask / th, 0.0));
}

// ------------------------------------------------------------------ atm_recover_large_step_variables_work, parts 1 and 2
// (1) cell-all, TI:3294-3350 (+ the garbage cell, TI:3282-3284)
__global__ void __launch_bounds__(CW_THREADS, MB_REC1) k2_recover_cell1(const Dev D, real dt, real invNs, int rk_step, real rcv, real rgas_p0) {
    CW_SETUP(D.nCells + 1)
    const b2 k_lt_nl = lv.lt(nl), k_mid = lv.ge(1) && lv.lt(nl);
    if (i == D.nCells) { ST(D.rho_zz_2, i, sel(k_lt_nl, mk2(1.0, 1.0), LD(D.rho_zz_2, i))); return; }
    const r2 rho_p = LD(D.rho_p_save, i) + LD(D.rho_pp, i);
    const r2 rho_zz = rho_p + LD(D.rho_base, i);
    const r2 rtb = LD(D.rtheta_base, i);
    const r2 zz = LD(D.zz, i);
    const r2 rw_save = LD(D.rw_save, i);
    const r2 wwAvg_in = LD(D.wwAvg, i), rw_in = LD(D.rw, i), w_in = LD(D.w_2, i);
    const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0);
    r2 rtheta_p;
    if (rk_step == 3) {
        rtheta_p = LD(D.rtheta_p_save, i) + LD(D.rtheta_pp, i)
                   - dt * rho_zz * LD(D.rt_diabatic_tend, i);
        const r2 arg = zz * (rgas_p0) * (rtheta_p + rtb);
        const r2 ex = mk2(pow_cr(arg.x, rcv), pow_cr(arg.y, rcv));
        ST(D.exner, i, sel(k_lt_nl, ex, 0.0));
        ST(D.pressure_p, i, sel(k_lt_nl, zz * RGAS * (ex * rtheta_p + rtb
                                                       * (ex - LD(D.exner_base, i))), 0.0));
    } else {
        rtheta_p = LD(D.rtheta_p_save, i) + LD(D.rtheta_pp, i);
    }
    ST(D.rho_p, i, sel(k_lt_nl, rho_p, 0.0));
    ST(D.rho_zz_2, i, sel(k_lt_nl, rho_zz, 0.0));
    ST(D.rtheta_p, i, sel(k_lt_nl, rtheta_p, 0.0));
    ST(D.theta_m_2, i, sel(k_lt_nl, (rtheta_p + rtb) / rho_zz, 0.0));
    // rows 0 and nl: rw = w = 0, wwAvg untouched; rows beyond nl (padding) keep what they held
    const b2 k_ends = lv.eq(0) || lv.eq(nl);
    const r2 rw = rw_save + LD(D.rw_p, i);
    ST(D.wwAvg, i, sel(k_mid, rw_save + (wwAvg_in * invNs), wwAvg_in));
    ST(D.rw, i, sel(k_mid, rw, sel(k_ends, mk2(0.0, 0.0), rw_in)));
    ST(D.w_2, i, sel(k_mid, rw / (fm * zz + fp * up1(zz)), sel(k_ends, mk2(0.0, 0.0), w_in)));
}
// (2) edge-all, TI:3360-3372
// dd_mode != 0: the divergence damping of the last small step (TI:3040-3060) has not been applied to ru_p yet and is applied
// here on the fly (1: ru_p holds the undamped value; 2: first small step, ru_p = ruAvg = dts * tend_u was never stored either).
// With halos the damped ru_p must exist before the exchange of TI:1322 on every edge a neighbour receives: those edges are damped
// by k2_divergence_damping (list form) and flagged in dd_done; the fold applies to the other edges this block computes itself
// (an owned cell on either side), and edges further out keep the value the exchange delivered.
__device__ __forceinline__ r2 dd_term(const Dev& D, int cell1, int cell2, int i, real coef_divdamp, unsigned uLDK, unsigned kc) {
    const real mask = 1.0 - D.specZoneMaskEdge[i];
    const r2 divCell1 = -(LD(D.rtheta_pp, cell1) - LD(D.rtheta_pp_old, cell1));
    const r2 divCell2 = -(LD(D.rtheta_pp, cell2) - LD(D.rtheta_pp_old, cell2));
    const r2 th = LD(D.theta_m, cell1) + LD(D.theta_m, cell2);
    return coef_divdamp * (divCell2 - divCell1) * mask / th;
}
__global__ void __launch_bounds__(CW_THREADS) k2_recover_edge(const Dev D, real invNs, int dd_mode_in, real coef_divdamp, real dts,
                                                              const unsigned char* dd_done) {
    CW_ENTER_R(D.nEdges)
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    const b2 k_lt_nl = lv.lt(nl);
    const int dd_mode = (dd_done && (dd_done[i] || !(cell1 < D.nCellsSolve || cell2 < D.nCellsSolve))) ? 0 : dd_mode_in;   // warp-uniform
    pdl_wait();
    const r2 rus = LD(D.ru_save, i);
    r2 ru_p, ruAvg;
    if (dd_mode == 2) { ru_p = dts * LD(D.tend_u, i); ruAvg = sel(k_lt_nl, ru_p, 0.0); }
    else { ru_p = LD(D.ru_p, i); ruAvg = LD(D.ruAvg, i); }
    if (dd_mode) ru_p = sel(k_lt_nl, ru_p + dd_term(D, cell1, cell2, i, coef_divdamp, uLDK, kc), 0.0);
    const r2 ru = rus + ru_p;
    const r2 rho2 = LD(D.rho_zz_2, cell1) + LD(D.rho_zz_2, cell2);
    ST(D.ruAvg, i, sel(k_lt_nl, rus + (ruAvg * invNs), 0.0));
    ST(D.ru, i, sel(k_lt_nl, ru, 0.0));
    ST(D.u_2, i, sel(k_lt_nl, 2. * ru / rho2, 0.0));
}

// ------------------------------------------------------------------ atm_advance_acoustic_step_work, edge part (small_step > 1)  TI:2751-2796
// dd_mode != 0: the divergence damping of the PREVIOUS small step is applied first, in registers (see k2_recover_edge): the
// separate damping kernel between two small steps disappears (its operands rtheta_pp, theta_m are gathered here anyway)
__global__ void __launch_bounds__(CW_THREADS) k2_acoustic_edge(const Dev D, real dts, real c2, int dd_mode, real coef_divdamp) {
    CW_ENTER(D.nEdges)
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    pdl_wait();
    if (!(cell1 < D.nCellsSolve || cell2 < D.nCellsSolve)) return;
    const b2 k_lt_nl = lv.lt(nl);
    const r2 tend_u = LD(D.tend_u, i);
    r2 ru_p, ruAvg;
    if (dd_mode == 2) { ru_p = dts * tend_u; ruAvg = sel(k_lt_nl, ru_p, 0.0); }
    else { ru_p = LD(D.ru_p, i); ruAvg = LD(D.ruAvg, i); }
    if (dd_mode) ru_p = sel(k_lt_nl, ru_p + dd_term(D, cell1, cell2, i, coef_divdamp, uLDK, kc), 0.0);
    r2 pgrad = ((LD(D.rtheta_pp, cell2) - LD(D.rtheta_pp, cell1)) * D.invDcEdge[i]) / (.5 * (LD(D.zz, cell2) + LD(D.zz, cell1)));
    pgrad = LD(D.cqu, i) * 0.5 * c2 * (LD(D.exner, cell1) + LD(D.exner, cell2)) * pgrad;
    pgrad = pgrad + 0.5 * LD(D.zxu, i) * GRAVITY * (LD(D.rho_pp, cell1) + LD(D.rho_pp, cell2));
    const r2 rup = ru_p + dts * (tend_u - (1.0 - D.specZoneMaskEdge[i]) * pgrad);
    ST(D.ru_p, i, sel(k_lt_nl, rup, 0.0));
    ST(D.ruAvg, i, sel(k_lt_nl, ruAvg + rup, 0.0));
}

// ------------------------------------------------------------------ atm_advance_scalars_work  TI:3575-3855
// edge value of every scalar ("horiz_flux_arr"), TI:3670-3751: one warp per edge, the stencil indices and the two
// possible weights per entry live one per lane and are broadcast; scalars are separate level-contiguous planes
__global__ void __launch_bounds__(CW_THREADS, 6 * 8 / CW_WARPS) k2_scalars_edge(const Dev D) {
    CW_SETUP_R(D.nEdges)
    const int nadv = D.nAdvCellsForEdge[i];
    int my_c = 0; real my_wp = 0.0, my_wm = 0.0;
    if (lane < nadv) {
        my_c = D.advCellsForEdge[(unsigned)i * 15 + lane];
        const real a = D.adv_coefs[(unsigned)i * 15 + lane], b = D.adv_coefs_3rd[(unsigned)i * 15 + lane];
        my_wp = a + b; my_wm = a - b;
    }
    const r2 ruavg = LD(D.ruAvg, i);
    const b2 pos = nonneg_sign(ruavg);
    const b2 k_lt_nl = lv.lt(nl);
    if (D.apply_lbcs && D.bdyMaskEdge[i] >= 4) {           // regional run, TI:3732-3750 (warp-uniform)
        if (D.bdyMaskEdge[i] > 5) return;                   // edges of the specified zone: nothing
        // the two outermost rings of the relaxation zone take the upwind cell value
        const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
        const real dv = D.dvEdge[i];
        const r2 u_direction = mk2(copysign((real)0.5, ruavg.x), copysign((real)0.5, ruavg.y));
        const r2 u_positive = dv * abs2(u_direction + 0.5), u_negative = dv * abs2(u_direction - 0.5);
        for (int s = 0; s < D.num_scalars; s++) {
            const real* q = D.scalars_2 + (size_t)s * D.cellPlane;
            ST(D.horiz_flux_arr + (size_t)s * D.edgePlane, i, sel(k_lt_nl, u_positive * LD(q, cell1) + u_negative * LD(q, cell2), 0.0));
        }
        return;
    }
    for (int s = 0; s < D.num_scalars; s++) {
        const real* __restrict__ q = D.scalars_2 + (size_t)s * D.cellPlane;
        r2 acc = mk2(0.0, 0.0);
        int j0 = 0;
        if (nadv == 10) {                                   // single expression, TI:3694-3704: no leading 0 +
            const real wp = BC(my_wp, 0), wm = BC(my_wm, 0);
            const r2 q2 = LD(q, BC(my_c, 0));
            acc = mk2((pos.x ? wp : wm) * q2.x, (pos.y ? wp : wm) * q2.y);
            j0 = 1;
        }
#pragma unroll 3
        for (int j = j0; j < nadv; j++) {
            const real wp = BC(my_wp, j), wm = BC(my_wm, j);
            const r2 q2 = LD(q, BC(my_c, j));
            acc.x = acc.x + (pos.x ? wp : wm) * q2.x;
            acc.y = acc.y + (pos.y ? wp : wm) * q2.y;
        }
        ST(D.horiz_flux_arr + (size_t)s * D.edgePlane, i, sel(k_lt_nl, acc, 0.0));
    }
}
// owned cells: flux divergence + vertical flux + update, TI:3773-3846
__global__ void __launch_bounds__(CW_THREADS, MB_SC_CELL) k2_scalars_cell(const Dev D, real dt, real weight_time_old, real weight_time_new, real coef3) {
    CW_SETUP(D.nCellsSolve)
    const int ne = D.nEdgesOnCell[i];
    const int le = min(lane, ne - 1);
    const int my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    const real my_sgn = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le];
    const real invArea = D.invAreaCell[i];
    if (D.bdyMaskCell[i] > 5) return;                  // regional run: the specified zone is not updated here, TI:3775
    const r2 rho_old = LD(D.rho_zz, i), rho_new = LD(D.rho_zz_2, i);
    const r2 rho_zz_new_inv = 1.0 / (weight_time_old * rho_old + weight_time_new * rho_new);
    const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0), rdzw = LD(D.rdzw, 0);
    const r2 ww = LD(D.wwAvg, i);
    const b2 k_lt_nl = lv.lt(nl);
    const b2 kk_edge = lv.eq(1) || lv.eq(nl - 1), kk_zero = lv.lt(1) || lv.ge(nl);
    for (int s = 0; s < D.num_scalars; s++) {
        real* qn = D.scalars_2 + (size_t)s * D.cellPlane;
        const real* __restrict__ hf = D.horiz_flux_arr + (size_t)s * D.edgePlane;
        r2 tend = mk2(0.0, 0.0);
#define SC_EDGE(E)                                                                                          \
        {                                                                                                   \
            const int iEdge = BC(my_e, (E));                                                                \
            const r2 term = BC(my_sgn, (E)) * LD(D.ruAvg, iEdge) * LD(hf, iEdge);                           \
            tend = selb((E) < ne, tend - term, tend);                                                       \
        }
#pragma unroll
        for (int e = 0; e < CW_NE; e++) SC_EDGE(e)
        for (int e = CW_NE; e < ne; e++) SC_EDGE(e)
#undef SC_EDGE
        tend = tend * invArea + 0.0;        // + scalar_tend_save, zero without physics (TI:3781-3783)
        const r2 q = LD(qn, i);
        const r2 qm1 = up1(q), qm2 = up2(q), qp1 = dn1(q);
        const r2 f2 = ww * (fm * q + fp * qm1);
        const r2 f3 = flux3_2(qm2, qm1, q, qp1, ww, coef3);
        const r2 fz = sel(kk_zero, mk2(0.0, 0.0), sel(kk_edge, f2, f3));
        const r2 f1 = dn1(fz);
        const r2 val = (LD(D.scalars + (size_t)s * D.cellPlane, i) * rho_old
                        + dt * (tend - rdzw * (f1 - fz))) * rho_zz_new_inv;
        ST(D.scalars_tend + (size_t)s * D.cellPlane, i, mk2(0.0, 0.0));
        ST(qn, i, sel(k_lt_nl, val, 0.0));
    }
}

// ------------------------------------------------------------------ atm_compute_vert_imp_coefs_work  TI:2225-2366 (block-tiled)
// Same tiling as k3_acoustic_cell.  Phase 1: each warp computes the acoustic coefficients cofwr, cofwz, coftz, cofwt and
// the tridiagonal rows (a, b, c) of VIC_COLS/VIC_WARPS columns, lane = level pair, vertical neighbours by shuffle.
// Phase 2: ONE warp runs the LU recurrence alpha = 1/(b - a*gamma), gamma = c*alpha (TI:2352-2355) of all VIC_COLS
// columns at once, lane = column, rows of odd stride in shared memory (alpha/gamma overwrite b/c).
// Phase 3: the warps write alpha_tri and gamma_tri.  Operation order inside a column is the reference's.
#ifndef VIC_COLS
#define VIC_COLS 32
#endif
#ifndef VIC_WARPS
#define VIC_WARPS 8
#endif
#ifndef MB_VIC
#define MB_VIC 3
#endif
__global__ void __launch_bounds__(VIC_WARPS * 32, MB_VIC) k3_vert_imp_coefs(const Dev D, real dtseps, real c2, real rcv) {
    extern __shared__ __align__(16) real smv[];
    PDL_ENTER
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int LDK = D.LDK, nl = D.nl;
    const int S = LDK | 1;
    real* s_a = smv;
    real* s_b = s_a + VIC_COLS * S;
    real* s_c = s_b + VIC_COLS * S;
    Lv lv; lv.k0 = 2 * lane;
    const int k0 = lv.k0; const bool act = k0 < D.LDKA;
    const unsigned uLDK = (unsigned)LDK, kc = (unsigned)min(k0, D.LDKA - 2);
    const b2 k_lt_nl = lv.lt(nl), k_mid = lv.ge(1) && lv.lt(nl);
    const int base = (gridDim.x - 1 - blockIdx.x) * VIC_COLS;      // backward sweep (see CW_SETUP_R)
    const r2 fzm = LD(D.fzm, 0), fzp = LD(D.fzp, 0), rdzw = LD(D.rdzw, 0), rdzu = LD(D.rdzu, 0);
    const r2 cofrz = dtseps * rdzw;
    if (blockIdx.x == 0 && wib == 0) ST(D.cofrz, 0, sel(k_lt_nl, cofrz, 0.0));
    const r2 rdzwm = up1(rdzw), cofrzm = up1(cofrz);
    for (int cc = 0; cc < VIC_COLS / VIC_WARPS; cc++) {
        const int c = cc * VIC_WARPS + wib;
        const int i = base + c;
        if (i >= D.nCellsSolve) continue;                   // warp-uniform
        const r2 zz = LD(D.zz, i), p = LD(D.exner, i), t = LD(D.theta_m_2, i), cqw = LD(D.cqw, i);
        const r2 qtotal = LD(D.qtot, i), rb = LD(D.rho_base, i), rtb = LD(D.rtheta_base, i), rt = LD(D.rtheta_p, i), pb = LD(D.exner_base, i);
        const r2 zzm = up1(zz);
        const r2 zint = fzm * zz + fzp * zzm;
        const r2 cofwr = sel(k_mid, .5 * dtseps * GRAVITY * zint, 0.0);
        const r2 cofwz = sel(k_mid, dtseps * c2 * zint
                                    * rdzu * cqw * (fzm * p + fzp * up1(p)), 0.0);
        const r2 coftz = sel(k_mid, dtseps * (fzm * t + fzp * up1(t)), 0.0);
        const r2 cofwt = sel(k_lt_nl, .5 * dtseps * rcv * zz * GRAVITY * rb / (1. + qtotal)
                                      * p / ((rtb + rt) * pb), 0.0);
        const r2 coftzm = up1(coftz), coftzp = dn1(coftz), cofwtm = up1(cofwt);
        const r2 a = -cofwz * coftzm * rdzwm * zzm
                     + cofwr * cofrzm
                     - cofwtm * coftzm * rdzwm;
        const r2 b = 1.
                     + cofwz * (coftz * rdzw * zz
                                + coftz * rdzwm * zzm)
                     - coftz * (cofwt * rdzw
                                - cofwtm * rdzwm)
                     + cofwr * (cofrz - cofrzm);
        const r2 cc_ = -cofwz * coftzp * rdzw * zz
                       - cofwr * cofrz
                       + cofwt * coftzp * rdzw;
        if (act) {
            const int o = c * S + k0;
            s_a[o] = a.x; s_a[o + 1] = a.y; s_b[o] = b.x; s_b[o + 1] = b.y; s_c[o] = cc_.x; s_c[o + 1] = cc_.y;
        }
        ST(D.cofwr, i, cofwr); ST(D.cofwz, i, cofwz); ST(D.coftz, i, coftz); ST(D.cofwt, i, cofwt);
        ST(D.a_tri, i, sel(k_mid, a, 0.0));
    }
    __syncthreads();
    if (wib == 0 && lane < VIC_COLS && base + lane < D.nCellsSolve) {
        const real* av = s_a + lane * S; real* bv = s_b + lane * S; real* cv = s_c + lane * S;
        real g = 0.;
#pragma unroll 4
        for (int kk = 1; kk < nl; kk++) {
            const real al = 1. / (bv[kk] - av[kk] * g);
            g = cv[kk] * al;
            bv[kk] = al; cv[kk] = g;
        }
    }
    __syncthreads();
    for (int cc = 0; cc < VIC_COLS / VIC_WARPS; cc++) {
        const int c = cc * VIC_WARPS + wib;
        const int i = base + c;
        if (i >= D.nCellsSolve) continue;
        const int o = c * S + (int)kc;
        ST(D.alpha_tri, i, sel(k_mid, mk2(s_b[o], s_b[o + 1]), 0.0));
        ST(D.gamma_tri, i, sel(k_mid, mk2(s_c[o], s_c[o + 1]), 0.0));
    }
}

// ------------------------------------------------------------------ atm_compute_dyn_tend_work, parts (c) and (d), rk 1
// del^2 of delsq_u on vertices (TI:5512-5529) and on cells (TI:5532-5551)
__global__ void __launch_bounds__(CW_THREADS) k2_dt_delsq_vertex(const Dev D) {
    CW_SETUP(D.nVertices)
    const int lj = min(lane, 2);
    const int my_e = D.edgesOnVertex[3 * i + lj];
    const real my_s = D.invAreaTriangle[i] * D.dcEdge[my_e] * D.edgesOnVertex_sign[3 * i + lj];
    r2 acc = mk2(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < 3; j++) acc = acc + BC(my_s, j) * LD(D.delsq_u, BC(my_e, j));
    ST(D.delsq_vorticity, i, sel(lv.lt(nl), acc, 0.0));
}
__global__ void __launch_bounds__(CW_THREADS) k2_dt_delsq_cell(const Dev D) {
    CW_SETUP_R(D.nCells)
    const int ne = D.nEdgesOnCell[i];
    const int le = min(lane, ne - 1);
    const int my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    const real my_s = D.invAreaCell[i] * D.dvEdge[my_e] * D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le];
    r2 acc = mk2(0.0, 0.0);
#define DELSQ_C(E) { const r2 du = LD(D.delsq_u, BC(my_e, (E))); acc = selb((E) < ne, acc + BC(my_s, (E)) * du, acc); }
#pragma unroll
    for (int e = 0; e < CW_NE; e++) DELSQ_C(e)
    for (int e = CW_NE; e < ne; e++) DELSQ_C(e)
#undef DELSQ_C
    ST(D.delsq_divergence, i, sel(lv.lt(nl), acc, 0.0));
}
// owned edges: del^4 of u (TI:5558-5584) and the final sum (TI:5694-5701).
// Restrictions (the host falls back to k_dt_edge_d otherwise): v_mom_eddy_visc2 == 0, no Rayleigh damping of u.
__global__ void __launch_bounds__(CW_THREADS) k2_dt_edge_d(const Dev D, const DynTendArgs A) {
    CW_SETUP(D.nEdgesSolve)
    r2 tue = LD(D.tend_u_euler, i);
    if (A.h_mom_eddy_visc4 > 0.0) {
        const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
        const int vertex1 = D.verticesOnEdge[2 * i], vertex2 = D.verticesOnEdge[2 * i + 1];
        const real u_mix_scale = D.meshScalingDel4[i] * A.h_mom_eddy_visc4;
        const real r_dc = u_mix_scale * A.del4u_div_factor * D.invDcEdge[i];
        const real r_dv = u_mix_scale * rmin(D.invDvEdge[i], 4 * D.invDcEdge[i]);
        const r2 u_diffusion = LD(D.rho_edge, i) * ((LD(D.delsq_divergence, cell2) - LD(D.delsq_divergence, cell1)) * r_dc
                                                   - (LD(D.delsq_vorticity, vertex2) - LD(D.delsq_vorticity, vertex1)) * r_dv);
        tue = tue - u_diffusion;
    }
    const r2 tu = LD(D.tend_u, i) + tue + LD(D.tend_ru_physics, i);
    const b2 k_lt_nl = lv.lt(nl);
    ST(D.tend_u_euler, i, sel(k_lt_nl, tue, 0.0));
    ST(D.tend_u, i, sel(k_lt_nl, tu, 0.0));
}

// ------------------------------------------------------------------ atm_advance_scalars_mono_work, edge part (C2)
// high-order flux of scalar s (TI:4356-4413: one expression when the stencil has 10 cells, a running sum otherwise),
// upwind flux and their difference (TI:4467-4487); same arithmetic as k_mono_edge2
__device__ __forceinline__ r2 max0(r2 a) { return mk2(rmax(0.0, a.x), rmax(0.0, a.y)); }
__device__ __forceinline__ r2 min0(r2 a) { return mk2(rmin(0.0, a.x), rmin(0.0, a.y)); }
// The five kernels of the scalar loop take the scalar s = s0 + blockIdx.y and the plane pl of the per-scalar work arrays
// (pl0 < 0: plane s, the batched launch with gridDim.y = num_scalars; otherwise plane pl0 for every scalar -- the arrays of
// the field table are the last plane, Dev::mb_planes)
__global__ void __launch_bounds__(CW_THREADS) k2_mono_edge2(const Dev D, int s0, int pl0, real dt) {
    CW_SETUP_R(D.nEdges)
    const int s = s0 + (int)blockIdx.y;
    const size_t pl = (size_t)(pl0 < 0 ? s : pl0);
    real* const flux_tmp = D.mb_flux_tmp + pl * D.edgePlane;
    real* const flux_upwind_tmp = D.mb_flux_upwind_tmp + pl * D.edgePlane;
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    const real* __restrict__ so = D.scalars + (size_t)s * D.cellPlane;
    const real* __restrict__ sn = D.scalars_2 + (size_t)s * D.cellPlane;
    // every load is unconditional (an inactive edge, one without an owned cell, runs zero stencil iterations)
    const int nadv_i = D.nAdvCellsForEdge[i];
    const int nadv = (cell1 < D.nCellsSolve || cell2 < D.nCellsSolve) ? nadv_i : 0;
    const int lj = min(lane, 14);
    const int my_c = D.advCellsForEdge[(unsigned)i * 15 + lj];
    const real my_a = D.adv_coefs[(unsigned)i * 15 + lj], my_b = D.adv_coefs_3rd[(unsigned)i * 15 + lj];
    const real my_wp = my_a + my_b, my_wm = my_a - my_b;
    const r2 uh = LD(D.ruAvg, i);
    const r2 so1 = LD(so, cell1), so2 = LD(so, cell2);
    const bool ten = nadv == 10;                                // warp-uniform: TI:4371-4384 vs TI:4386-4399
    const bool px = uh.x > 0, py = uh.y > 0;
    const real sx = sign1(uh.x), sy = sign1(uh.y);
    r2 acc = mk2(0.0, 0.0);
#pragma unroll 5
    for (int j = 0; j < nadv; j++) {
        const r2 q2 = LD(sn, BC(my_c, j));
        const real a = BC(my_a, j), b = BC(my_b, j), wp = BC(my_wp, j), wm = BC(my_wm, j);
        const real wx = ten ? (px ? wp : wm) : uh.x * (a + sx * b);
        const real wy = ten ? (py ? wp : wm) : uh.y * (a + sy * b);
        acc.x = acc.x + wx * q2.x;
        acc.y = acc.y + wy * q2.y;
    }
    const r2 flux = ten ? uh * acc : acc;
    const r2 fup = D.dvEdge[i] * dt * (max0(uh) * so1 + min0(uh) * so2);
    const b2 k_lt_nl = lv.lt(nl);
    // TI:4479 (and :4592), as written there: `config_apply_lbcs .and. (m == nRelaxZone) .or. (m == nRelaxZone-1)`
    const int m_bdy = D.bdyMaskEdge[i];
    const bool upwind_only = (D.apply_lbcs && m_bdy == 5) || m_bdy == 4;
    ST(flux_upwind_tmp, i, sel(k_lt_nl, fup, 0.0));
    ST(flux_tmp, i, upwind_only ? mk2(0.0, 0.0) : sel(k_lt_nl, dt * flux - fup, 0.0));
}

// ------------------------------------------------------------------ atm_advance_scalars_mono_work, the other parts
__device__ __forceinline__ r2 max2(r2 a, r2 b) { return mk2(rmax(a.x, b.x), rmax(a.y, b.y)); }
__device__ __forceinline__ r2 min2(r2 a, r2 b) { return mk2(rmin(a.x, b.x), rmin(a.y, b.y)); }
// (B) owned cells: re-integrated density, TI:4177-4204
__global__ void __launch_bounds__(CW_THREADS) k2_mono_rho_int(const Dev D, real dt) {
    CW_SETUP_R(D.nCellsSolve)
    const int ne = D.nEdgesOnCell[i];
    const int le = min(lane, ne - 1);
    const int my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    const real my_sgn = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le], my_dv = D.dvEdge[my_e];
    const real invArea = D.invAreaCell[i];
    r2 r = mk2(0.0, 0.0);
#define RHO_INT_EDGE(E) { const r2 ru = LD(D.ruAvg, BC(my_e, (E))); r = selb((E) < ne, r - BC(my_sgn, (E)) * ru * BC(my_dv, (E)) * invArea, r); }
#pragma unroll
    for (int e = 0; e < CW_NE; e++) RHO_INT_EDGE(e)
    for (int e = CW_NE; e < ne; e++) RHO_INT_EDGE(e)
#undef RHO_INT_EDGE
    const r2 ww = LD(D.wwAvg, i);
    ST(D.rho_zz_int, i, sel(lv.lt(nl), LD(D.rho_zz, i) + dt * (r - LD(D.rdzw, 0) * (dn1(ww) - ww)), 0.0));
}
// (C1) owned cells: vertical fluxes, bounds, vertical part of the upwind update and of scale_arr  TI:4277-4344, 4426-4459
__global__ void __launch_bounds__(CW_THREADS) k2_mono_cell1(const Dev D, int s0, int pl0, real dt, real coef3) {
    CW_SETUP(D.nCellsSolve)
    const int s = s0 + (int)blockIdx.y;
    const size_t pl = (size_t)(pl0 < 0 ? s : pl0);
    real* const scale_in = D.mb_scale + 2 * pl * D.cellPlane;
    const real* __restrict__ so = D.scalars + (size_t)s * D.cellPlane;
    const real* __restrict__ sn = D.scalars_2 + (size_t)s * D.cellPlane;
    const int ne = D.nEdgesOnCell[i];
    const int my_c = D.cellsOnCell[(unsigned)i * D.maxEdges + min(lane, ne - 1)];
    const r2 q = LD(so, i), n = LD(sn, i), ww = LD(D.wwAvg, i), rho = LD(D.rho_zz, i);
    const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0), rdnw = LD(D.rdzw, 0);
    const r2 qm1 = up1(q), qp1 = dn1(q), nm1 = up1(n), nm2 = up2(n), np1 = dn1(n);
    const b2 k_lt_nl = lv.lt(nl), k_mid = lv.ge(1) && lv.lt(nl), kk_edge = lv.eq(1) || lv.eq(nl - 1);
    // interface k: upwind flux and (high-order - upwind) flux; zero at the boundaries
    const r2 fu = dt * (max0(ww) * qm1 + min0(ww) * q);
    const r2 raw = sel(kk_edge, ww * (fm * n + fp * nm1), flux3_2(nm2, nm1, n, np1, ww, coef3));
    const r2 wd0 = sel(k_mid, dt * raw - fu, 0.0);
    const r2 wd1 = dn1(wd0);
    // bounds over the column and the neighbouring cells
    r2 smax = sel(lv.eq(0), max2(q, qp1), sel(lv.eq(nl - 1), max2(q, qm1), max2(max2(qm1, q), qp1)));
    r2 smin = sel(lv.eq(0), min2(q, qp1), sel(lv.eq(nl - 1), min2(q, qm1), min2(min2(qm1, q), qp1)));
#define MONO_NBR(E) { const r2 v = LD(so, BC(my_c, (E))); smax = selb((E) < ne, max2(smax, v), smax); smin = selb((E) < ne, min2(smin, v), smin); }
#pragma unroll
    for (int e = 0; e < CW_NE; e++) MONO_NBR(e)
    for (int e = CW_NE; e < ne; e++) MONO_NBR(e)
#undef MONO_NBR
    // upwind vertical update, TI:4428-4446
    r2 snew = q * rho;
    snew = sel(lv.lt(nl - 1), snew - dn1(fu) * rdnw, snew);
    snew = sel(lv.ge(1), snew + fu * rdnw, snew);
    ST(D.mb_wdtn + pl * D.cellPlane, i, wd0);
    ST(D.mb_s_max + pl * D.cellPlane, i, sel(k_lt_nl, smax, 0.0));
    ST(D.mb_s_min + pl * D.cellPlane, i, sel(k_lt_nl, smin, 0.0));
    ST(D.mb_scalar_new + pl * D.cellPlane, i, sel(k_lt_nl, snew, 0.0));
    ST(scale_in, i, sel(k_lt_nl, -rdnw * (min0(wd1) - max0(wd0)), 0.0));                        // SCALE_IN
    ST(scale_in + D.cellPlane, i, sel(k_lt_nl, -rdnw * (max0(wd1) - min0(wd0)), 0.0));          // SCALE_OUT
}
// (C3) owned cells: horizontal part of the upwind update and of scale_arr (4496-4513) and the limiter (4523-4553)
__global__ void __launch_bounds__(CW_THREADS) k2_mono_cell3(const Dev D, int pl0, const real* rho_lim) {
    CW_SETUP(D.nCellsSolve)
    const size_t pl = (size_t)(pl0 < 0 ? (int)blockIdx.y : pl0);
    real* const scale_in = D.mb_scale + 2 * pl * D.cellPlane;
    real* const scalar_new = D.mb_scalar_new + pl * D.cellPlane;
    const real* const flux_tmp = D.mb_flux_tmp + pl * D.edgePlane;
    const real* const flux_upwind_tmp = D.mb_flux_upwind_tmp + pl * D.edgePlane;
    const int ne = D.nEdgesOnCell[i];
    const int le = min(lane, ne - 1);
    const int my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    const real my_sgn = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le];
    const real invArea = D.invAreaCell[i];
    r2 snew = LD(scalar_new, i), sin_ = LD(scale_in, i), sout = LD(scale_in + D.cellPlane, i);
#define MONO3_EDGE(E)                                                                                       \
    {                                                                                                       \
        const int iEdge = BC(my_e, (E)); const real sg = BC(my_sgn, (E));                                   \
        const r2 ft = LD(flux_tmp, iEdge), fup = LD(flux_upwind_tmp, iEdge);                                \
        snew = selb((E) < ne, snew - sg * fup * invArea, snew);                                             \
        sout = selb((E) < ne, sout - max0(sg * ft) * invArea, sout);                                        \
        sin_ = selb((E) < ne, sin_ - min0(sg * ft) * invArea, sin_);                                        \
    }
#pragma unroll
    for (int e = 0; e < CW_NE; e++) MONO3_EDGE(e)
    for (int e = CW_NE; e < ne; e++) MONO3_EDGE(e)
#undef MONO3_EDGE
    const real eps = 1.e-20;
    const r2 rl = LD(rho_lim, i);
    const r2 f_in = (LD(D.mb_s_max + pl * D.cellPlane, i) * rl - snew) / (sin_ + eps);
    const r2 f_out = (LD(D.mb_s_min + pl * D.cellPlane, i) * rl - snew) / (sout - eps);
    const b2 k_lt_nl = lv.lt(nl);
    ST(scalar_new, i, sel(k_lt_nl, snew, 0.0));
    ST(scale_in, i, sel(k_lt_nl, min2(splat(1.0), max0(f_in)), 0.0));
    ST(scale_in + D.cellPlane, i, sel(k_lt_nl, min2(splat(1.0), max0(f_out)), 0.0));
}
// (D1) edges of owned cells: rescale the anti-diffusive flux (4579-4623)
__global__ void __launch_bounds__(CW_THREADS) k2_mono_edge4(const Dev D, int pl0) {
    CW_ENTER_R(D.nEdges)
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    const size_t pl = (size_t)(pl0 < 0 ? (int)blockIdx.y : pl0);
    pdl_wait();
    if (!(cell1 < D.nCellsSolve || cell2 < D.nCellsSolve)) return;
    const real* s_in = D.mb_scale + 2 * pl * D.cellPlane; const real* s_out = s_in + D.cellPlane;
    const r2 flux = LD(D.mb_flux_tmp + pl * D.edgePlane, i);
    const r2 f = max0(flux) * min2(LD(s_out, cell1), LD(s_in, cell2))
               + min0(flux) * min2(LD(s_in, cell1), LD(s_out, cell2));
    ST(D.mb_flux_arr + pl * D.edgePlane, i, sel(lv.lt(nl), f, 0.0));
}
// (D2) all cells: rescaled vertical flux (4636-4645), final update (4651-4674), positive-definite copy-out (4708-4715)
__global__ void __launch_bounds__(CW_THREADS) k2_mono_cell5(const Dev D, int s0, int pl0, const real* rho_div) {
    CW_SETUP(D.nCells)
    const int s = s0 + (int)blockIdx.y;
    const size_t pl = (size_t)(pl0 < 0 ? s : pl0);
    const real* const flux_arr = D.mb_flux_arr + pl * D.edgePlane;
    real* out = D.scalars_2 + (size_t)s * D.cellPlane;
    const b2 k_lt_nl = lv.lt(nl);
    if (D.bdyMaskCell[i] > 2) return;                  // TI:4709 `bdyMaskCell <= nSpecZone`: these cells are set after the transport
    if (i >= D.nCellsSolve) { ST(out, i, sel(k_lt_nl, max0(LD(out, i)), 0.0)); return; }       // warp-uniform
    const int ne = D.nEdgesOnCell[i];
    const int le = min(lane, ne - 1);
    const int my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    const real my_sgn = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le];
    const real invArea = D.invAreaCell[i];
    const r2 s_in = LD(D.mb_scale + 2 * pl * D.cellPlane, i), s_out = LD(D.mb_scale + (2 * pl + 1) * D.cellPlane, i), wd = LD(D.mb_wdtn + pl * D.cellPlane, i);
    const r2 w0 = sel(lv.ge(1) && k_lt_nl, max0(wd) * min2(up1(s_out), s_in) + min0(wd) * min2(s_out, up1(s_in)), 0.0);
    const r2 w1 = dn1(w0);
    r2 snew = LD(D.mb_scalar_new + pl * D.cellPlane, i);
#define MONO5_EDGE(E) { const r2 fa = LD(flux_arr, BC(my_e, (E))); snew = selb((E) < ne, snew - BC(my_sgn, (E)) * fa * invArea, snew); }
#pragma unroll
    for (int e = 0; e < CW_NE; e++) MONO5_EDGE(e)
    for (int e = CW_NE; e < ne; e++) MONO5_EDGE(e)
#undef MONO5_EDGE
    snew = (snew + (-LD(D.rdzw, 0) * (w1 - w0))) / LD(rho_div, i);
    ST(out, i, sel(k_lt_nl, max0(snew), 0.0));
}

// ------------------------------------------------------------------ atm_compute_dyn_tend_work, part (f) split in two
// k2_dt_cell_f carries the operands of both the w and the theta_m tendency (80 registers with 192 bytes of spills, 24
// warps per SM).  The two tendencies share only rw and the cell's connectivity, so they are also available as two
// kernels with about half the live state each; same expressions, same order: bit-identical to k2_dt_cell_f.
#ifndef MB_CELL_FW
#define MB_CELL_FW 5
#endif
#ifndef MB_CELL_FT
#define MB_CELL_FT 4
#endif
__global__ void __launch_bounds__(CW_THREADS, MB_CELL_FW) k2_dt_cell_fw(const Dev D, const DynTendArgs A) {     // tend_w, TI:5713-5757, 5838-5945
    CW_SETUP_R(D.nCellsSolve)
    const int ne = D.nEdgesOnCell[i];
    const int le = min(lane, ne - 1);
    const int my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    const real my_sgn = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le];
    const real invArea = D.invAreaCell[i];
    r2 tw = mk2(0.0, 0.0);
#define CELL_FW_EDGE(E) { const r2 fxw = LD(D.adv_flux_w, BC(my_e, (E))); tw = selb((E) < ne, tw - BC(my_sgn, (E)) * fxw, tw); }
#pragma unroll
    for (int e = 0; e < CW_NE; e++) CELL_FW_EDGE(e)
    for (int e = CW_NE; e < ne; e++) CELL_FW_EDGE(e)
#undef CELL_FW_EDGE
    r2 twe = LD(D.tend_w_euler, i);
    const r2 twe_in = twe;
    const b2 k_ge1 = lv.ge(1), k_lt_nl = lv.lt(nl);
    if (A.rk_step == 1 && A.h_mom_eddy_visc4 > 0.0) {
        const int my_c1 = D.cellsOnEdge[2 * my_e], my_c2 = D.cellsOnEdge[2 * my_e + 1];
        const real my_f = D.meshScalingDel4[my_e];
        const real my_dv = D.dvEdge[my_e], my_idc = D.invDcEdge[my_e];
        const real r_areaCell = A.h_mom_eddy_visc4 * invArea;
#define CELL_FW_DEL4(E)                                                                                     \
        {                                                                                                   \
            const real edge_sign = BC(my_f, (E)) * r_areaCell * BC(my_dv, (E)) * BC(my_sgn, (E)) * BC(my_idc, (E)); \
            const r2 d = LD(D.delsq_w, BC(my_c2, (E))) - LD(D.delsq_w, BC(my_c1, (E)));                     \
            twe = selb((E) < ne, twe - edge_sign * d, twe);                                                 \
        }
#pragma unroll
        for (int e = 0; e < CW_NE; e++) CELL_FW_DEL4(e)
        for (int e = CW_NE; e < ne; e++) CELL_FW_DEL4(e)
#undef CELL_FW_DEL4
    }
    const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0), rdzu = LD(D.rdzu, 0);
    const r2 rw = LD(D.rw, i), w = LD(D.w_2, i);
    const r2 rwm1 = up1(rw);
    const b2 kk_edge = lv.eq(1) || lv.eq(nl - 1), kk_zero = lv.lt(1) || lv.ge(nl);
    const r2 wm1 = up1(w), wm2 = up2(w), wp1 = dn1(w);
    const r2 f2 = 0.25 * (rw + rwm1) * (w + wm1);
    const r2 f3 = flux3_2(wm2, wm1, w, wp1, 0.5 * (rw + rwm1), 1.0);
    const r2 fz = sel(kk_zero, mk2(0.0, 0.0), sel(kk_edge, f2, f3));
    const r2 f1 = dn1(fz);
    tw = tw * invArea - rdzu * (f1 - fz);
    if (A.rk_step == 1) {
        const r2 pp = LD(D.pressure_p, i), dpdz = LD(D.dpdz, i), cqw = LD(D.cqw, i);
        const r2 twe_new = twe - cqw * (rdzu * (pp - up1(pp)) - (fm * dpdz + fp * up1(dpdz)));
        twe = sel(k_ge1 && k_lt_nl, twe_new, twe_in);
        ST(D.tend_w_euler, i, twe);
    }
    ST(D.tend_w, i, sel(k_ge1 && k_lt_nl, tw + twe, 0.0));
}
__global__ void __launch_bounds__(CW_THREADS, MB_CELL_FT) k2_dt_cell_ft(const Dev D, const DynTendArgs A) {     // tend_theta, TI:5956-6016, 6066-6126, 6134-6197
    CW_SETUP_R(D.nCellsSolve)
    const int ne = D.nEdgesOnCell[i];
    const int le = min(lane, ne - 1);
    const int my_e = D.edgesOnCell[(unsigned)i * D.maxEdges + le];
    const real my_sgn = D.edgesOnCell_sign[(unsigned)i * D.maxEdges + le];
    const int my_c1 = D.cellsOnEdge[2 * my_e], my_c2 = D.cellsOnEdge[2 * my_e + 1];
    const real my_dv = D.dvEdge[my_e];
    const real invArea = D.invAreaCell[i];
    r2 tt = mk2(0.0, 0.0);
#define CELL_FT_EDGE(E) { const r2 fxt = LD(D.adv_flux_theta, BC(my_e, (E))); tt = selb((E) < ne, tt - BC(my_sgn, (E)) * fxt, tt); }
#pragma unroll
    for (int e = 0; e < CW_NE; e++) CELL_FT_EDGE(e)
    for (int e = CW_NE; e < ne; e++) CELL_FT_EDGE(e)
#undef CELL_FT_EDGE
    r2 tte = LD(D.tend_theta_euler, i);
    if (A.rk_step > 1) {          // perturbation flux for the rtheta_pp equation, TI:5995-6016
#define CELL_FT_PERT(E)                                                                                     \
        {                                                                                                   \
            const int iEdge = BC(my_e, (E)), cell1 = BC(my_c1, (E)), cell2 = BC(my_c2, (E));                \
            const real sg = BC(my_sgn, (E)), dv = BC(my_dv, (E));                                           \
            const r2 flux = sg * dv * (LD(D.ru_save, iEdge) - LD(D.ru, iEdge)) * 0.5 * (LD(D.theta_m, cell2) + LD(D.theta_m, cell1)); \
            tt = selb((E) < ne, tt - flux, tt);                                                             \
        }
#pragma unroll 3
        for (int e = 0; e < CW_NE; e++) CELL_FT_PERT(e)
        for (int e = CW_NE; e < ne; e++) CELL_FT_PERT(e)
#undef CELL_FT_PERT
    }
    const b2 k_lt_nl = lv.lt(nl);
    if (A.rk_step == 1 && A.h_theta_eddy_visc4 > 0.0) {
        const real my_f = D.meshScalingDel4[my_e], my_idc = D.invDcEdge[my_e];
        const real r_areaCell = A.h_theta_eddy_visc4 * A.prandtl_inv * invArea;
#define CELL_FT_DEL4(E)                                                                                     \
        {                                                                                                   \
            const real edge_sign = BC(my_f, (E)) * r_areaCell * BC(my_dv, (E)) * BC(my_sgn, (E)) * BC(my_idc, (E)); \
            const r2 d = LD(D.delsq_theta, BC(my_c2, (E))) - LD(D.delsq_theta, BC(my_c1, (E)));             \
            tte = selb((E) < ne, tte - edge_sign * d, tte);                                                 \
        }
#pragma unroll
        for (int e = 0; e < CW_NE; e++) CELL_FT_DEL4(e)
        for (int e = CW_NE; e < ne; e++) CELL_FT_DEL4(e)
#undef CELL_FT_DEL4
    }
    const r2 fm = LD(D.fzm, 0), fp = LD(D.fzp, 0), rdzw = LD(D.rdzw, 0);
    const r2 rw = LD(D.rw, i), t = LD(D.theta_m_2, i), ts = LD(D.theta_m, i), rws = LD(D.rw_save, i);
    const r2 rho = LD(D.rho_zz_2, i), tend_rho = LD(D.tend_rho, i), rtdiab = LD(D.rt_diabatic_tend, i);
    const r2 trp = LD(D.tend_rtheta_physics, i);
    const b2 kk_zero = lv.lt(1) || lv.ge(nl);
    const r2 tm1 = up1(t), tm2 = up2(t), tp1 = dn1(t), tsm1 = up1(ts);
    const r2 ftop = rws * (fm * t + fp * tm1);
    const r2 flow = rw * (fm * t + fp * tm1);
    const r2 f3 = flux3_2(tm2, tm1, t, tp1, rw, A.coef_3rd_order);
    const r2 fpert = sel(lv.eq(1), flow, f3) + (rws - rw) * (fm * ts + fp * tsm1);
    const r2 fz = sel(kk_zero, mk2(0.0, 0.0), sel(lv.eq(nl - 1), ftop, fpert));
    const r2 f1 = dn1(fz);
    tt = tt * invArea - rdzw * (f1 - fz);
    const r2 out_rthdynten = sel(k_lt_nl, (tt - tend_rho * t) / rho, 0.0);
    tt = tt + rho * rtdiab;
    if (A.rk_step == 1) ST(D.tend_theta_euler, i, sel(k_lt_nl, tte, 0.0));
    ST(D.rthdynten, i, out_rthdynten);
    ST(D.tend_theta, i, sel(k_lt_nl, tt + tte + trp, 0.0));
}
