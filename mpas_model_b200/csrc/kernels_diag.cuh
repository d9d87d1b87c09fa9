// kernels_diag.cuh -- solve diagnostics, coupled-diagnostics initialisation, small
// element-wise helpers, min/max summary and host<->device layout conversion.
#pragma once
#include "kernels_dyn.cuh"

// ------------------------------------------------------------------ atm_compute_solve_diagnostics_work  TI:6337-6773
// u, h are passed explicitly: the routine runs on time level 1 at init (mpas_atm_core.F:524) and 2 in the step.
// (1) vertex-all: vorticity (6452-6472), ke_vertex (6548-6561, ke_edge recomputed inline), pv_vertex (6647-6659)
__global__ void k_diag_vertex(const Dev D, const real* __restrict__ u) {
    KI;
    if (i >= D.nVertices || k >= nl) return;
    real vort = 0.0;
    real ke3[3];
    for (int j = 0; j < 3; j++) {
        const int iEdge = D.edgesOnVertex[3 * i + j];
        const real dc = D.dcEdge[iEdge];
        const real uu = AT(u, iEdge, k);
        const real s = D.edgesOnVertex_sign[3 * i + j] * dc;
        vort = vort + s * uu;
        const real efac = dc * D.dvEdge[iEdge];
        ke3[j] = efac * (uu * uu);
    }
    const real iat = D.invAreaTriangle[i];
    vort = vort * iat;
    AT(D.vorticity, i, k) = vort;
    const real r = 0.25 * iat;
    AT(D.ke_vertex, i, k) = (ke3[0] + ke3[1] + ke3[2]) * r;
    AT(D.pv_vertex, i, k) = (D.fVertex[i] + vort);
}
// (2) cell-all: divergence (6479-6499), ke (6515-6534) + Hollingsworth blend (6569-6593), pv_cell (6693-6709)
__global__ void k_diag_cell(const Dev D, const real* __restrict__ u, int apvm) {
    KI;
    if (i >= D.nCells || k >= nl) return;
    const int ne = D.nEdgesOnCell[i];
    const real r = D.invAreaCell[i];
    real div = 0.0, ke = 0.0;
    for (int e = 0; e < ne; e++) {
        const int iEdge = D.edgesOnCell[(size_t)i * D.maxEdges + e];
        const real dv = D.dvEdge[iEdge];
        const real uu = AT(u, iEdge, k);
        const real s = D.edgesOnCell_sign[(size_t)i * D.maxEdges + e] * dv;
        div = div + s * uu;
        const real efac = D.dcEdge[iEdge] * dv;
        ke = ke + 0.25 * (efac * (uu * uu));
    }
    AT(D.divergence, i, k) = div * r;
    ke = ke * r;
    const real ke_fact = 1.0 - .375;
    ke = ke_fact * ke;
    real pvc = 0.0;
    for (int e = 0; e < ne; e++) {
        const int iVertex = D.verticesOnCell[(size_t)i * D.maxEdges + e];
        const int j = D.kiteForCell[(size_t)i * D.maxEdges + e];
        const real kite = D.kiteAreasOnVertex[3 * iVertex + j];
        ke = ke + (1. - ke_fact) * kite * AT(D.ke_vertex, iVertex, k) * r;
        if (apvm) pvc = pvc + kite * AT(D.pv_vertex, iVertex, k) * r;
    }
    AT(D.ke, i, k) = ke;
    if (apvm) AT(D.pv_cell, i, k) = pvc;
}
// (3) edge-all: h_edge (6428-6435), tangential velocity v (6618-6632, rk 3 only), pv_edge with APVM upwinding (6673-6745)
__global__ void k_diag_edge(const Dev D, const real* __restrict__ u, const real* __restrict__ h,
                            int reconstruct_v, int apvm, real apvm_dt) {
    KI;
    if (i >= D.nEdges || k >= nl) return;
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    const int vertex1 = D.verticesOnEdge[2 * i], vertex2 = D.verticesOnEdge[2 * i + 1];
    AT(D.rho_edge, i, k) = 0.5 * (AT(h, cell1, k) + AT(h, cell2, k));
    real vv;
    if (reconstruct_v) {
        vv = 0.0;
        const int neoe = D.nEdgesOnEdge[i];
        for (int j = 0; j < neoe; j++) {
            const int eoe = D.edgesOnEdge[(size_t)i * D.maxEdges2 + j];
            vv = vv + D.weightsOnEdge[(size_t)i * D.maxEdges2 + j] * AT(u, eoe, k);
        }
        AT(D.v, i, k) = vv;
    } else {
        vv = AT(D.v, i, k);
    }
    const real pv1 = AT(D.pv_vertex, vertex1, k), pv2 = AT(D.pv_vertex, vertex2, k);
    real pve = 0.5 * (pv1 + pv2);
    if (apvm) {
        const real r1 = 1.0 * D.invDvEdge[i];
        const real r2 = 1.0 * D.invDcEdge[i];
        const real gt = (pv2 - pv1) * r1;
        const real gn = (AT(D.pv_cell, cell2, k) - AT(D.pv_cell, cell1, k)) * r2;
        AT(D.gradPVt, i, k) = gt;
        AT(D.gradPVn, i, k) = gn;
        pve = pve - apvm_dt * (vv * gt + AT(u, i, k) * gn);
    }
    AT(D.pv_edge, i, k) = pve;
}

// ------------------------------------------------------------------ atm_init_coupled_diagnostics  TI:6776-7010 (time level 1)
__global__ void k_initcd_cell1(const Dev D, real rvord, real rcv, real rgas_p0) {
    KI;
    if (i >= D.nCells || k >= nl) return;
    RP qv = D.scalars + (size_t)D.index_qv * D.cellPlane;
    const real zzk = AT(D.zz, i, k);
    const real theta_m = AT(D.theta, i, k) * (1. + rvord * AT(qv, i, k));
    const real rho_zz = AT(D.rho, i, k) / zzk;
    AT(D.theta_m, i, k) = theta_m;
    AT(D.rho_zz, i, k) = rho_zz;
    const real rb = AT(D.rho_base, i, k), tb = AT(D.theta_base, i, k);
    const real rho_p = rho_zz - rb;
    const real rtb = tb * rb;
    const real rtp = theta_m * rho_p + rb * (theta_m - tb);
    const real ex = pow_cr(zzk * (rgas_p0) * (rtp + rtb), rcv);
    const real exb = pow_cr(zzk * (rgas_p0) * (rtb), rcv);
    AT(D.rho_p, i, k) = rho_p;
    AT(D.rtheta_base, i, k) = rtb;
    AT(D.rtheta_p, i, k) = rtp;
    AT(D.exner, i, k) = ex;
    AT(D.exner_base, i, k) = exb;
    AT(D.pressure_p, i, k) = zzk * RGAS * (ex * rtp + rtb * (ex - exb));
    AT(D.pressure_base, i, k) = zzk * RGAS * exb * rtb;
}
__global__ void k_initcd_edge(const Dev D) {
    KI;
    if (i >= D.nEdges || k >= nl) return;
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    AT(D.ru, i, k) = 0.5 * AT(D.u, i, k) * (AT(D.rho_zz, cell1, k) + AT(D.rho_zz, cell2, k));
}
__global__ void k_initcd_cell2(const Dev D) {
    KI;
    if (i >= D.nCells || k > nl) return;
    if (k == 0 || k == nl) { AT(D.rw, i, k) = 0.0; return; }
    const real fm = D.fzm[k], fp = D.fzp[k];
    const real zzf = (fp * AT(D.zz, i, k - 1) + fm * AT(D.zz, i, k));
    real rw = AT(D.w, i, k)
              * (fp * AT(D.rho_zz, i, k - 1) + fm * AT(D.rho_zz, i, k))
              * zzf;
    const int ne = D.nEdgesOnCell[i];
    for (int e = 0; e < ne; e++) {
        const int iEdge = D.edgesOnCell[(size_t)i * D.maxEdges + e];
        const real flux = (fm * AT(D.ru, iEdge, k) + fp * AT(D.ru, iEdge, k - 1));
        const size_t zi = ((size_t)i * D.maxEdges + e) * LDK + k;
        rw = rw - D.edgesOnCell_sign[(size_t)i * D.maxEdges + e] * (D.zb_cell[zi] + sign1(flux) * D.zb3_cell[zi]) * flux
                  * zzf;
    }
    AT(D.rw, i, k) = rw;
}

// ------------------------------------------------------------------ element-wise helpers (atm_rk_dynamics_substep_finish TI:7121-7172)
// zb_any(i) = any non-zero zb_cell / zb3_cell entry of cell i (derived once per upload of those fields)
__global__ void k_zb_flags(const Dev D) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > (size_t)D.nCells) return;
    const size_t n = (size_t)D.maxEdges * D.LDK;
    const real* a = D.zb_cell + i * n; const real* b = D.zb3_cell + i * n;
    int any = 0;
    for (size_t j = 0; j < n; j++) any |= (a[j] != 0.0) | (b[j] != 0.0);
    D.zb_any[i] = any;
}
// One launch for a whole list of element-wise column copies/accumulations (atm_rk_integration_setup, TI:1930-2039, and
// atm_rk_dynamics_substep_finish, TI:7013-7191): blockIdx.y selects the segment, blockIdx.x strides over its
// 16-byte element pairs.  op 0: d = s;  1: d = s + d;  2: d = s, then s = d * scale;  3: d = s + d, then s = d * scale.
#define SEG_MAX 16
struct Seg { real* d; real* s; unsigned n2; int op; };        // n2 = number of 16-byte pairs (LDK is even)
struct SegList { Seg seg[SEG_MAX]; real scale; };
#define SEG_BLOCKS 296                                        // blocks per segment: 2 per SM
__global__ void __launch_bounds__(256) k_segments(const SegList L) {
    PDL_ENTER
    const Seg g = L.seg[blockIdx.y];
#ifdef MPASB_SINGLE
    typedef float2 v2;
#else
    typedef double2 v2;
#endif
    v2* __restrict__ d = reinterpret_cast<v2*>(g.d);
    v2* __restrict__ s = reinterpret_cast<v2*>(g.s);
    const unsigned stride = gridDim.x * blockDim.x;
#define SEG_AT(t) (t)
    if (g.op == 0) {
        for (unsigned t0 = blockIdx.x * blockDim.x + threadIdx.x; t0 < g.n2; t0 += stride) { const unsigned t = SEG_AT(t0); d[t] = s[t]; }
    } else if (g.op == 1) {
        for (unsigned t0 = blockIdx.x * blockDim.x + threadIdx.x; t0 < g.n2; t0 += stride) {
            const unsigned t = SEG_AT(t0);
            const v2 x = s[t], y = d[t]; v2 r; r.x = x.x + y.x; r.y = x.y + y.y; d[t] = r;
        }
    } else {
        for (unsigned t0 = blockIdx.x * blockDim.x + threadIdx.x; t0 < g.n2; t0 += stride) {
            const unsigned t = SEG_AT(t0);
            v2 r = s[t];
            if (g.op == 3) { const v2 y = d[t]; r.x = r.x + y.x; r.y = r.y + y.y; }
            d[t] = r;
            v2 q; q.x = r.x * L.scale; q.y = r.y * L.scale; s[t] = q;
        }
    }
#undef SEG_AT
}

// ------------------------------------------------------------------ summarize_timestep  TI:8286-8319
__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a, assumed;
    do { assumed = old; if (__longlong_as_double(assumed) <= v) break;
         old = atomicCAS(a, assumed, __double_as_longlong(v)); } while (assumed != old);
}
__device__ __forceinline__ void atomic_max_f64(double* addr, double v) {
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a, assumed;
    do { assumed = old; if (__longlong_as_double(assumed) >= v) break;
         old = atomicCAS(a, assumed, __double_as_longlong(v)); } while (assumed != old);
}
// out[0..1] = min/max over x[0:n_items][0:nl]; reductions start from 0.0 as the reference's do (TI:8291-8292)
__global__ void k_minmax(const real* __restrict__ x, int n_items, int nl, int LDK, double* out) {
    real mn = 0.0, mx = 0.0;
    const size_t total = (size_t)n_items * LDK;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        if ((int)(t % LDK) < nl) { const real v = x[t]; mn = rmin(mn, v); mx = rmax(mx, v); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = rmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = rmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { atomic_min_f64(out, (double)mn); atomic_max_f64(out + 1, (double)mx); }
}

// the same reduction plus a NaN count (fmin/fmax drop NaNs; TI:8258-8281 tests every element with ieee_is_nan)
__global__ void k_minmax_nan(const real* __restrict__ x, int n_items, int nl, int LDK, double* out, unsigned long long* n_nan) {
    real mn = 0.0, mx = 0.0; unsigned bad = 0;
    const size_t total = (size_t)n_items * LDK;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        if ((int)(t % LDK) < nl) { const real v = x[t]; mn = rmin(mn, v); mx = rmax(mx, v); bad += (v != v); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = rmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = rmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomic_min_f64(out, (double)mn); atomic_max_f64(out + 1, (double)mx);
        if (bad) atomicAdd(n_nan, (unsigned long long)bad);
    }
}

// ------------------------------------------------------------------ dense host layout <-> padded device layout
// dst[o][0:LDK] <- src[o][0:ninner] (zero padded)
__global__ void k_pad(real* __restrict__ dst, const real* __restrict__ src, size_t nouter, int ninner, int LDK) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nouter * LDK) return;
    const size_t o = t / LDK; const int k = (int)(t % LDK);
    dst[t] = k < ninner ? src[o * ninner + k] : 0.0;
}
__global__ void k_unpad(real* __restrict__ dst, const real* __restrict__ src, size_t nouter, int ninner, int LDK) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nouter * ninner) return;
    const size_t o = t / ninner; const int k = (int)(t % ninner);
    dst[t] = src[o * LDK + k];
}
// planes: host [n][ninner][P] (Fortran (P, ninner, n)) <-> device [P][n][LDK]
__global__ void k_pad_planes(real* __restrict__ dst, const real* __restrict__ src, size_t n, int ninner, int P, int LDK) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)P * n * LDK) return;
    const int k = (int)(t % LDK); const size_t o = (t / LDK) % n; const int p = (int)(t / ((size_t)LDK * n));
    dst[t] = k < ninner ? src[(o * ninner + k) * P + p] : 0.0;
}
__global__ void k_unpad_planes(real* __restrict__ dst, const real* __restrict__ src, size_t n, int ninner, int P, int LDK) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)P * n * ninner) return;
    const int p = (int)(t % P); const int k = (int)((t / P) % ninner); const size_t o = t / ((size_t)P * ninner);
    dst[t] = src[((size_t)p * n + o) * LDK + k];
}
// planes with the plane index between: host [n][P][ninner] (Fortran (ninner, P, n)) <-> device [P][n][LDK]
__global__ void k_pad_midplanes(real* __restrict__ dst, const real* __restrict__ src, size_t n, int ninner, int P, int LDK) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)P * n * LDK) return;
    const int k = (int)(t % LDK); const size_t o = (t / LDK) % n; const int p = (int)(t / ((size_t)LDK * n));
    dst[t] = k < ninner ? src[(o * P + p) * ninner + k] : 0.0;
}
__global__ void k_unpad_midplanes(real* __restrict__ dst, const real* __restrict__ src, size_t n, int ninner, int P, int LDK) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)P * n * ninner) return;
    const int k = (int)(t % ninner); const int p = (int)((t / ninner) % P); const size_t o = t / ((size_t)P * ninner);
    dst[t] = src[((size_t)p * n + o) * LDK + k];
}
__global__ void k_int_to_zero_based(int* __restrict__ dst, const int* __restrict__ src, size_t n, int sub) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dst[t] = src[t] - sub;
}

// a few values from device memory into pinned (device-accessible) host memory, without a copy engine
__global__ void k_copy_to_host(double* __restrict__ host, const double* __restrict__ dev, int n) {
    for (int t = threadIdx.x; t < n; t += blockDim.x) host[t] = dev[t];
    __threadfence_system();
}

// ------------------------------------------------------------------ halo exchange pack / unpack
// One launch packs every (field, list element, level) of one exchange group.  seg[] describes
// contiguous runs of the send buffer: run r moves `count` columns of field `fld` listed in
// idx[idx_off : idx_off+count] to buf[buf_off + j*width + k], width = levels moved per column.
struct HaloSeg { real* field; int idx_off; int count; int width; int peer; size_t buf_off; int stride; };     // peer: index into the plan's peer list
__global__ void k_halo_pack(const HaloSeg* __restrict__ seg, const int* __restrict__ idx, real* __restrict__ buf, int nseg) {
    for (int s = blockIdx.y; s < nseg; s += gridDim.y) {
        const HaloSeg g = seg[s];
        // 32-bit index arithmetic: a segment holds count * width < 2^31 reals, and a 64-bit division per element costs more than the copy
        const unsigned total = (unsigned)g.count * (unsigned)g.width, w = (unsigned)g.width;
        for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
            const unsigned j = t / w, k = t - j * w;
            buf[g.buf_off + t] = g.field[(size_t)idx[g.idx_off + j] * g.stride + k];
        }
    }
}

// ---- the same exchange without NCCL: stores straight into the neighbour's mailbox over NVLink (CUDA IPC peer memory) ----
// Every rank owns a mailbox [source rank][2 buffers][slot] and a flag array {arrived[source], consumed[destination]} of
// 64-bit message counters; both are mapped into every peer with cudaIpcOpenMemHandle.  k_halo_put packs the send lists
// directly into the peers' mailboxes (buffer = message number & 1), and its last block publishes the message number in
// each peer's arrived[] after a system-scope fence.  k_halo_get waits for arrived[], unpacks from the local mailbox
// (ld.global.cg: the lines were written by another GPU) and its last block publishes consumed[] to the senders, which
// is what a sender checks before it reuses a buffer two messages later.
#define P2P_MAXP 8
struct P2PPeers {
    int n;
    real* remote[P2P_MAXP];                          // start of my message in peer p's mailbox
    const real* local[P2P_MAXP];                     // start of peer p's message in my mailbox
    unsigned long long* remote_arrived[P2P_MAXP];    // peer p's arrived[my rank]
    unsigned long long* remote_consumed[P2P_MAXP];   // peer p's consumed[my rank]
    const unsigned long long* local_arrived[P2P_MAXP];   // my arrived[p]
    const unsigned long long* local_consumed[P2P_MAXP];  // my consumed[p]
    unsigned long long seq_send[P2P_MAXP], seq_recv[P2P_MAXP];   // 0: nothing to send / receive in this exchange
    size_t send_off[P2P_MAXP], recv_off[P2P_MAXP];   // start of the peer's part in the plan's contiguous layout (elements)
};
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__global__ void k_halo_put(const HaloSeg* __restrict__ seg, const int* __restrict__ idx, int nseg, const P2PPeers pp, unsigned* done) {
    for (int s = blockIdx.y; s < nseg; s += gridDim.y) {
        const HaloSeg g = seg[s];
        const int p = g.peer;
        if (threadIdx.x == 0 && pp.seq_send[p] > 2)            // the buffer last held message seq - 2: wait until it was unpacked
            while (ld_acquire_sys(pp.local_consumed[p]) < pp.seq_send[p] - 2) { }
        __syncthreads();
        real* dst = pp.remote[p] + (g.buf_off - pp.send_off[p]);
        // 32-bit index arithmetic: a segment holds count * width < 2^31 reals, and a 64-bit division per element costs more than the copy
        const unsigned total = (unsigned)g.count * (unsigned)g.width, w = (unsigned)g.width;
        for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
            const unsigned j = t / w, k = t - j * w;
            dst[t] = g.field[(size_t)idx[g.idx_off + j] * g.stride + k];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(done, 1u);
        if (prev == gridDim.x * gridDim.y - 1) {               // last block: every store of the grid is visible system-wide
            *done = 0;
            __threadfence_system();
            for (int p = 0; p < pp.n; p++) if (pp.seq_send[p]) st_release_sys(pp.remote_arrived[p], pp.seq_send[p]);
        }
    }
}
__global__ void k_halo_get(const HaloSeg* __restrict__ seg, const int* __restrict__ idx, int nseg, const P2PPeers pp, unsigned* done) {
    for (int s = blockIdx.y; s < nseg; s += gridDim.y) {
        const HaloSeg g = seg[s];
        const int p = g.peer;
        if (threadIdx.x == 0) while (ld_acquire_sys(pp.local_arrived[p]) < pp.seq_recv[p]) { }
        __syncthreads();
        const real* src = pp.local[p] + (g.buf_off - pp.recv_off[p]);
        // 32-bit index arithmetic: a segment holds count * width < 2^31 reals, and a 64-bit division per element costs more than the copy
        const unsigned total = (unsigned)g.count * (unsigned)g.width, w = (unsigned)g.width;
        for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
            const unsigned j = t / w, k = t - j * w;
            g.field[(size_t)idx[g.idx_off + j] * g.stride + k] = __ldcg(src + t);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(done, 1u);
        if (prev == gridDim.x * gridDim.y - 1) {               // last block: the mailbox buffers of this exchange are free again
            *done = 0;
            for (int p = 0; p < pp.n; p++) if (pp.seq_recv[p]) st_release_sys(pp.remote_consumed[p], pp.seq_recv[p]);
        }
    }
}
__global__ void k_halo_unpack(const HaloSeg* __restrict__ seg, const int* __restrict__ idx, const real* __restrict__ buf, int nseg) {
    for (int s = blockIdx.y; s < nseg; s += gridDim.y) {
        const HaloSeg g = seg[s];
        // 32-bit index arithmetic: a segment holds count * width < 2^31 reals, and a 64-bit division per element costs more than the copy
        const unsigned total = (unsigned)g.count * (unsigned)g.width, w = (unsigned)g.width;
        for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
            const unsigned j = t / w, k = t - j * w;
            g.field[(size_t)idx[g.idx_off + j] * g.stride + k] = buf[g.buf_off + t];
        }
    }
}
