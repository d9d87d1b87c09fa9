// mpasb.cu -- host side of the B200-native MPAS-A dycore step and its C ABI (include/mpasb.h).
//
// atm_srk3 (mpas_atm_time_integration.F:803-1725) is re-expressed as a fixed sequence of
// kernel launches on one CUDA stream over device-resident, level-contiguous fields; the
// reference's per-routine host<->device copies (e.g. TI:2739-2749 / 2976-2982) do not exist
// here.  Halo exchanges are pack-kernel -> NCCL send/recv -> unpack-kernel on the same stream.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <string>
#include <vector>
#include <algorithm>
#include <dlfcn.h>
#include <cuda_runtime.h>
#include "../../include/mpasb.h"
#include "kernels_dyn.cuh"
#include "kernels_acoustic.cuh"
#include "kernels_diag.cuh"
#include "kernels_transport.cuh"
#include "kernels_col.cuh"
#include "kernels_output.cuh"
#include "kernels_lbc.cuh"
#include "halo.cuh"

enum { T_REAL = 0, T_INT = 1 };
enum { TG_NONE = 0, TG_LOCAL, TG_CELL, TG_EDGE, TG_VERTEX };

struct FieldRec {
    const char* name; int loc, inner, levels, type, target;
    void* d[2];            // device buffers per time level
    void* alloc = nullptr; // start of the allocation when d[0] is not (work arrays of the batched monotonic transport: d[0] is the last plane)
    size_t dev_count;      // elements per buffer
    long host_count;       // dense host elements
};

struct ProfRec { double ms = 0; long count = 0; };

struct mpasb_handle_s {
    mpasb_dims dims; mpasb_config cfg; int device = 0;
    cudaStream_t stream = nullptr;
    Dev D;
    std::vector<FieldRec> fields;
    std::map<std::string, int> index;
    void* staging = nullptr; size_t staging_bytes = 0;
    double* d_minmax = nullptr;
    // asynchronous summarize_timestep: device results (2 * (2 + num_scalars) doubles + 2 NaN counters as doubles' bit
    // patterns), pinned host copy, completion event
    // (two slots, so that the summary of one request can be fetched while the next request is already enqueued)
    enum { SUMMARY_RING = 4 };      // summaries that may be pending at once (requests in flight, mpasb_summarize_timestep_async)
    double* d_summary[SUMMARY_RING] = {}; double* h_summary[SUMMARY_RING] = {}; cudaEvent_t ev_summary[SUMMARY_RING] = {}; long summary_head = 0, summary_tail = 0;
    std::string err;
    long launches = 0;
    int cpb = 4;
    bool colwarp = false;          // LDK <= 64 and <= CW_MAXNE edges per cell: the column-warp kernels apply
    int max_ne = 0;                // max(nEdgesOnCell), known once the mesh is uploaded
    bool zb_dirty = true;          // zb_any must be recomputed before the next step
    bool ru_p_pending = false;     // first-small-step ru_p/ruAvg still to be written by the divergence-damping kernel
    bool dd_deferred = false;      // divergence damping of the last small step still to be applied (by the next edge kernel)
    real dd_coef = 0.0, dd_dts = 0.0;
    real lbc_dt_end = 0.0;         // regional runs: seconds from the start of the next step to the end of the LBC interval
    bool fuse_dd = true;           // MPASB_NO_DD_FUSE=1: always run the damping as its own kernel
    bool pdl = true;               // MPASB_PDL=0: no programmatic dependent launch
    int* d_dd_edges = nullptr; unsigned char* d_dd_done = nullptr; int n_dd_edges = 0; bool dd_lists_ok = false, dd_partial = false;   // build_dd_lists
    int* d_ac_bnd = nullptr; int* d_ac_int = nullptr; int n_ac_bnd = 0, n_ac_int = 0; bool ac_lists_ok = false;   // build_acoustic_lists
    bool profile = false;
    bool smem_attr_vic = false, smem_attr_ac = false, smem_attr_ac9 = false;   // opt-in to > 48 KB of dynamic shared memory, per handle because it is per device
    std::map<std::string, ProfRec> prof;
    // stencil-union tiles of the TMA-staged advective flux kernel (k4_dt_edge_flux): host copies of the lists they
    // are derived from, and the derived device tables
    std::vector<int> hc_advCells, hc_nAdv, hc_cellsOnEdge;
    // host copies for the canonical-neighbourhood tables of the cell-centred flux sweep (build_flux_rings)
    std::vector<int> hc_cellsOnCell, hc_edgesOnCell, hc_nEdgesOnCell;
    std::vector<real> hc_adv_coefs, hc_adv_coefs_3rd, hc_weightsOnEdge;
    std::vector<int> hc_edgesOnEdge, hc_nEdgesOnEdge;
    bool cor_dirty = true, cor_ok = false;
    bool rings_dirty = true, rings_ok = false;
    bool relaxed = true;           // re-associated / FMA kernels allowed (parity bar 1e-11, not bit equality); MPASB_STRICT=1 turns it off
    long n_regular = 0;
    bool tiles_dirty = true, tiles_ok = false;
    int4* d_tile_hdr = nullptr; int4* d_tile_runs = nullptr; unsigned char* d_tile_slot = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, kev0 = nullptr, kev1 = nullptr, tev0 = nullptr, tev1 = nullptr;
    // halo exchanges that do not feed the next kernel run on comm_stream, concurrently with that kernel
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
    bool comm_pending = false, overlap = true;
    HaloState halo;
    // batched, stream-ordered field transfers (mpasb_set_fields_async / mpasb_get_fields_async): their own copy streams and
    // staging areas, so that the upload of the next request, the step and the download of the previous one overlap
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    void* stage_in = nullptr; void* stage_out = nullptr; size_t stage_in_bytes = 0, stage_out_bytes = 0;
    enum { XFER_RING = 4 };         // completion events of the last XFER_RING batches per direction (mpasb_wait_*_lag)
    cudaEvent_t ev_in_ready[XFER_RING] = {}, ev_in_free = nullptr, ev_out_ready = nullptr, ev_out_done[XFER_RING] = {};
    long n_in = 0, n_out = 0;       // batches enqueued so far
};
typedef mpasb_handle_s H;

#define CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return 1; } } while (0)

static int inner_dense1(const H* h, int in) {
    const mpasb_dims& d = h->dims;
    switch (in) {
        case IN_ONE: return 1; case IN_NL: return d.nVertLevels; case IN_NL1: return d.nVertLevels + 1;
        case IN_ME: return d.maxEdges; case IN_ME2: return d.maxEdges2; case IN_VD: return d.vertexDegree;
        case IN_TWO: return 2; case IN_F15: return 15; case IN_NL1_ME: return d.nVertLevels + 1;
        case IN_S_NL: return d.num_scalars; case IN_NL_TWO: return d.nVertLevels; case IN_THREE_ME: return 3;
    }
    return 1;
}
static int inner_dense2(const H* h, int in) {
    switch (in) { case IN_NL1_ME: return h->dims.maxEdges; case IN_S_NL: return h->dims.nVertLevels; case IN_NL_TWO: return 2; case IN_THREE_ME: return h->dims.maxEdges; default: return 1; }
}
static size_t outer_of(const H* h, int loc) {
    switch (loc) { case LOC_CELL: return (size_t)h->dims.nCells + 1; case LOC_EDGE: return (size_t)h->dims.nEdges + 1;
                   case LOC_VERTEX: return (size_t)h->dims.nVertices + 1; default: return 1; }
}
static size_t dev_count_of(const H* h, int loc, int in) {
    const size_t o = outer_of(h, loc); const size_t L = h->D.LDK;
    switch (in) {
        case IN_NL: case IN_NL1: return o * L;
        case IN_NL1_ME: return o * h->dims.maxEdges * L;
        case IN_S_NL: return o * L * h->dims.num_scalars;
        case IN_NL_TWO: return o * L * 2;
        default: return o * inner_dense1(h, in) * inner_dense2(h, in);
    }
}

static void set_dev_ptr(H* h, const FieldRec& f);

extern "C" int mpasb_create(const mpasb_dims* dims, const mpasb_config* cfg, int device, mpasb_handle* out) {
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return 2;      // fail loudly: no CPU fallback
    if (device < 0 || device >= ndev) return 3;
    if (dims->nVertLevels < 4 || dims->maxEdges < 1) return 4;
    H* h = new H();
    h->dims = *dims; h->cfg = *cfg; h->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return 5; }
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return 6; }
    cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1); cudaEventCreate(&h->kev0); cudaEventCreate(&h->kev1);
    cudaEventCreate(&h->tev0); cudaEventCreate(&h->tev1);
    {   // highest priority: the small pack / NCCL / unpack kernels must get SM slots while a full-grid compute kernel runs
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, greatest);
    }
    cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming); cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming);
    h->overlap = !getenv("MPASB_NO_OVERLAP");
    h->relaxed = !mpasb_strict_arithmetic();
    h->fuse_dd = !getenv("MPASB_NO_DD_FUSE");
    if (const char* e = getenv("MPASB_PDL")) h->pdl = atoi(e) != 0;
    memset(&h->D, 0, sizeof(Dev));
    h->D.pf_next = 1;
    if (const char* pf = getenv("MPASB_PF_NEXT")) h->D.pf_next = atoi(pf);
    Dev& D = h->D;
    D.nCells = dims->nCells; D.nEdges = dims->nEdges; D.nVertices = dims->nVertices;
    D.nCellsSolve = dims->nCellsSolve; D.nEdgesSolve = dims->nEdgesSolve; D.nVerticesSolve = dims->nVerticesSolve;
    D.nl = dims->nVertLevels; D.LDKA = (dims->nVertLevels + 1 + 1) / 2 * 2; D.LDK = D.LDKA;
    if (const char* al = getenv("MPASB_LDK_ALIGN")) { const int a = std::max(2, atoi(al)) / 2 * 2; D.LDK = (D.LDKA + a - 1) / a * a; }
    D.maxEdges = dims->maxEdges; D.maxEdges2 = dims->maxEdges2; D.num_scalars = dims->num_scalars;
    D.apply_lbcs = cfg->config_apply_lbcs != 0;
    D.index_qv = dims->index_qv - 1; D.moist_start = dims->moist_start - 1; D.moist_end = dims->moist_end - 1;
    D.cellPlane = (size_t)(dims->nCells + 1) * D.LDK; D.edgePlane = (size_t)(dims->nEdges + 1) * D.LDK;
    {   // batched monotonic transport (MPASB_MONO_BATCH=0: one scalar at a time, as the reference's loop TI:4220)
        const char* e = getenv("MPASB_MONO_BATCH");
        const bool on = (!e || atoi(e) != 0) && dims->num_scalars > 1 && cfg->config_scalar_advection && (cfg->config_monotonic || cfg->config_positive_definite);
        D.mb_planes = on ? dims->num_scalars : 1;
    }
    h->cpb = std::max(1, 256 / D.LDK);
#define F(name_, loc_, inner_, lev_, type_, tgt_) { FieldRec f; f.name = #name_; f.loc = LOC_##loc_; f.inner = IN_##inner_; \
        f.levels = lev_; f.type = T_##type_; f.target = TG_##tgt_; f.d[0] = f.d[1] = nullptr; h->fields.push_back(f); }
#include "../../include/mpasb_fields.def"
#undef F
    size_t max_bytes = 0;
    for (size_t n = 0; n < h->fields.size(); n++) {
        FieldRec& f = h->fields[n];
        h->index[f.name] = (int)n;
        f.dev_count = dev_count_of(h, f.loc, f.inner);
        f.host_count = (long)(outer_of(h, f.loc) * inner_dense1(h, f.inner) * inner_dense2(h, f.inner));
        const size_t esz = f.type == T_REAL ? sizeof(real) : sizeof(int);
        // work arrays of the monotonic transport: one plane per scalar when the transport is batched (Dev::mb_planes)
        real** mb = nullptr;
        for (const auto& kv : std::initializer_list<std::pair<const char*, real**>>{{"wdtn", &D.mb_wdtn}, {"s_max", &D.mb_s_max}, {"s_min", &D.mb_s_min},
                 {"scalar_new", &D.mb_scalar_new}, {"scale_arr", &D.mb_scale}, {"flux_tmp", &D.mb_flux_tmp}, {"flux_upwind_tmp", &D.mb_flux_upwind_tmp},
                 {"flux_arr", &D.mb_flux_arr}}) if (!strcmp(f.name, kv.first)) mb = kv.second;
        const size_t planes = mb ? (size_t)D.mb_planes : 1;
        for (int l = 0; l < f.levels; l++) {
            if (cudaMalloc(&f.d[l], planes * f.dev_count * esz) != cudaSuccess) { h->err = "cudaMalloc failed"; mpasb_destroy(h); return 7; }
            cudaMemsetAsync(f.d[l], 0, planes * f.dev_count * esz, h->stream);
        }
        if (mb) { *mb = (real*)f.d[0]; if (planes > 1) { f.alloc = f.d[0]; f.d[0] = (real*)f.d[0] + (planes - 1) * f.dev_count; } }
        max_bytes = std::max(max_bytes, (size_t)f.host_count * esz);
        set_dev_ptr(h, f);
    }
    h->staging_bytes = max_bytes;
    if (cudaMalloc(&h->staging, max_bytes) != cudaSuccess || cudaMalloc(&h->d_minmax, 4 * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->D.zb_any, ((size_t)dims->nCells + 1) * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&h->D.adv_flux_w, h->D.edgePlane * sizeof(real)) != cudaSuccess ||
        cudaMalloc(&h->D.adv_flux_theta, h->D.edgePlane * sizeof(real)) != cudaSuccess) {
        mpasb_destroy(h); return 8;
    }
    cudaStreamSynchronize(h->stream);
    *out = h;
    return 0;
}

static void set_dev_ptr(H* h, const FieldRec& f) {
    Dev& D = h->D;
#define FIELD_REAL(name_) if (!strcmp(f.name, #name_)) { D.name_ = (real*)f.d[0]; D.name_##_2 = (real*)f.d[1]; return; }
#define FIELD_INT(name_) if (!strcmp(f.name, #name_)) { D.name_ = (int*)f.d[0]; return; }
#define F(name_, loc_, inner_, lev_, type_, tgt_) FIELD_##type_(name_)
#include "../../include/mpasb_fields.def"
#undef F
#undef FIELD_REAL
#undef FIELD_INT
}

extern "C" int mpasb_destroy(mpasb_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    halo_destroy(h->halo);
    for (FieldRec& f : h->fields) for (int l = 0; l < 2; l++) if (f.d[l]) cudaFree(l == 0 && f.alloc ? f.alloc : f.d[l]);
    if (h->staging) cudaFree(h->staging);
    if (h->stage_in) cudaFree(h->stage_in);
    if (h->stage_out) cudaFree(h->stage_out);
    for (cudaEvent_t e : {h->ev_in_free, h->ev_out_ready}) if (e) cudaEventDestroy(e);
    for (int q = 0; q < H::XFER_RING; q++) { if (h->ev_in_ready[q]) cudaEventDestroy(h->ev_in_ready[q]); if (h->ev_out_done[q]) cudaEventDestroy(h->ev_out_done[q]); }
    if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    if (h->d_minmax) cudaFree(h->d_minmax);
    if (h->d_ac_bnd) { cudaFree(h->d_ac_bnd); cudaFree(h->d_ac_int); }
    if (h->d_dd_edges) { cudaFree(h->d_dd_edges); cudaFree(h->d_dd_done); }
    for (int q = 0; q < H::SUMMARY_RING; q++) {
        if (h->d_summary[q]) cudaFree(h->d_summary[q]);
        if (h->h_summary[q]) cudaFreeHost(h->h_summary[q]);
        if (h->ev_summary[q]) cudaEventDestroy(h->ev_summary[q]);
    }
    if (h->D.zb_any) cudaFree(h->D.zb_any);
    if (h->D.adv_flux_w) cudaFree(h->D.adv_flux_w);
    if (h->D.adv_flux_theta) cudaFree(h->D.adv_flux_theta);
    for (void* p : {(void*)h->D.fx_ring, (void*)h->D.fx_w, (void*)h->D.hdiv_w, (void*)h->D.hdiv_theta, (void*)h->D.cor_w, (void*)h->D.cor_slot, (void*)h->D.cor_part}) if (p) cudaFree(p);
    if (h->d_tile_hdr) cudaFree(h->d_tile_hdr);
    if (h->d_tile_runs) cudaFree(h->d_tile_runs);
    if (h->d_tile_slot) cudaFree(h->d_tile_slot);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (cudaEvent_t e : {h->kev0, h->kev1, h->tev0, h->tev1}) if (e) cudaEventDestroy(e);
    if (h->ev_ready) cudaEventDestroy(h->ev_ready);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

extern "C" const char* mpasb_last_error(mpasb_handle h) { return h ? h->err.c_str() : "null handle"; }
extern "C" long mpasb_kernel_launch_count(mpasb_handle h) { return h->launches; }
// 1 (MPASB_STRICT=1 in the environment when the handle is created): every kernel keeps the reference's operation order and
// nothing is contracted into FMAs, so results are bit-identical to the fp64 CPU arithmetic.  0 (default): the kernels that
// re-associate sums for register-level reuse (the cell-centred flux sweep, explicit fma()) are used where their tables
// apply; results then agree with the strict path to rounding (north-star bar: rel-L2 <= 1e-11 after one step).
extern "C" int mpasb_strict_arithmetic(void) { const char* s = getenv("MPASB_STRICT"); return (s && atoi(s) != 0) ? 1 : 0; }
extern "C" int mpasb_real_bytes(void) { return (int)sizeof(mpasb_real); }
extern "C" int mpasb_synchronize(mpasb_handle h) { cudaSetDevice(h->device); CUDA_OK(cudaStreamSynchronize(h->stream)); CUDA_OK(cudaGetLastError()); return 0; }

static FieldRec* find_field(H* h, const char* name) {
    auto it = h->index.find(name);
    if (it == h->index.end()) { h->err = std::string("unknown field ") + name; return nullptr; }
    return &h->fields[it->second];
}
extern "C" int mpasb_field_count(mpasb_handle h, const char* name, long* count) {
    FieldRec* f = find_field(h, name); if (!f) return 1; *count = f->host_count; return 0;
}

static inline unsigned nblk(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

static bool is_padded(const FieldRec* f) {
    return f->inner == IN_NL || f->inner == IN_NL1 || f->inner == IN_NL1_ME || f->inner == IN_S_NL || f->inner == IN_NL_TWO;
}
// dense host layout (in `st`, device memory) -> padded device field, and back; on the compute stream
static void launch_pad(H* h, const FieldRec* f, real* dst, const real* st) {
    const int LDK = h->D.LDK; const size_t o = outer_of(h, f->loc); const int n1 = inner_dense1(h, f->inner);
    if (f->inner == IN_NL || f->inner == IN_NL1) k_pad<<<nblk(o * LDK), 256, 0, h->stream>>>(dst, st, o, n1, LDK);
    else if (f->inner == IN_NL1_ME) k_pad<<<nblk(o * h->dims.maxEdges * LDK), 256, 0, h->stream>>>(dst, st, o * h->dims.maxEdges, n1, LDK);
    else if (f->inner == IN_S_NL) k_pad_planes<<<nblk(f->dev_count), 256, 0, h->stream>>>(dst, st, o, h->dims.nVertLevels, h->dims.num_scalars, LDK);
    else k_pad_midplanes<<<nblk(f->dev_count), 256, 0, h->stream>>>(dst, st, o, h->dims.nVertLevels, 2, LDK);
    h->launches++;
}
static void launch_unpad(H* h, const FieldRec* f, real* st, const real* src) {
    const int LDK = h->D.LDK; const size_t o = outer_of(h, f->loc); const int n1 = inner_dense1(h, f->inner); const long count = f->host_count;
    if (f->inner == IN_NL || f->inner == IN_NL1) k_unpad<<<nblk(count), 256, 0, h->stream>>>(st, src, o, n1, LDK);
    else if (f->inner == IN_NL1_ME) k_unpad<<<nblk(count), 256, 0, h->stream>>>(st, src, o * h->dims.maxEdges, n1, LDK);
    else if (f->inner == IN_S_NL) k_unpad_planes<<<nblk(count), 256, 0, h->stream>>>(st, src, o, h->dims.nVertLevels, h->dims.num_scalars, LDK);
    else k_unpad_midplanes<<<nblk(count), 256, 0, h->stream>>>(st, src, o, h->dims.nVertLevels, 2, LDK);
    h->launches++;
}

// ------------------------------------------------------------------ batched asynchronous transfers
// A request = "these host arrays in, one step, those host arrays out".  With the three calls below the upload of request n+1
// (PCIe host->device on h2d_stream), the step of request n (compute stream) and the download of request n-1 (device->host on
// d2h_stream) overlap; within one request the order upload -> step -> download is kept by events.  One staging area per
// direction holds the dense host layout of every field of a batch; pad / unpad kernels run on the compute stream, so the order
// of device-field accesses is simply the order of the calls.
static int ensure_transfer_state(H* h, size_t in_bytes, size_t out_bytes) {
    if (!h->h2d_stream) {
        CUDA_OK(cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
        CUDA_OK(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
        for (cudaEvent_t* e : {&h->ev_in_free, &h->ev_out_ready}) CUDA_OK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        for (int q = 0; q < H::XFER_RING; q++) { CUDA_OK(cudaEventCreateWithFlags(&h->ev_in_ready[q], cudaEventDisableTiming)); CUDA_OK(cudaEventCreateWithFlags(&h->ev_out_done[q], cudaEventDisableTiming)); }
    }
    if (in_bytes > h->stage_in_bytes) {
        CUDA_OK(cudaStreamSynchronize(h->stream)); CUDA_OK(cudaStreamSynchronize(h->h2d_stream));
        if (h->stage_in) cudaFree(h->stage_in);
        CUDA_OK(cudaMalloc(&h->stage_in, in_bytes)); h->stage_in_bytes = in_bytes;
    }
    if (out_bytes > h->stage_out_bytes) {
        CUDA_OK(cudaStreamSynchronize(h->stream)); CUDA_OK(cudaStreamSynchronize(h->d2h_stream));
        if (h->stage_out) cudaFree(h->stage_out);
        CUDA_OK(cudaMalloc(&h->stage_out, out_bytes)); h->stage_out_bytes = out_bytes;
    }
    return 0;
}
static int batch_fields(H* h, int n, const char* const* names, const int* levels, const long* counts, std::vector<FieldRec*>& fs, size_t& bytes) {
    bytes = 0;
    for (int q = 0; q < n; q++) {
        FieldRec* f = find_field(h, names[q]); if (!f) return 1;
        if (f->type != T_REAL || levels[q] < 1 || levels[q] > f->levels || counts[q] != f->host_count) { h->err = std::string("bad field in batch: ") + names[q]; return 2; }
        fs.push_back(f); bytes += ((size_t)f->host_count * sizeof(real) + 255) / 256 * 256;
    }
    return 0;
}
// Host arrays -> device fields.  The host arrays may be reused after mpasb_wait_uploads (or any later synchronising call).
extern "C" int mpasb_set_fields_async(mpasb_handle h, int n, const char* const* names, const int* levels, const mpasb_real* const* src, const long* counts) {
    cudaSetDevice(h->device);
    std::vector<FieldRec*> fs; size_t bytes = 0;
    if (int rc = batch_fields(h, n, names, levels, counts, fs, bytes)) return rc;
    if (ensure_transfer_state(h, bytes, 0)) return 3;
    if (h->n_in) CUDA_OK(cudaStreamWaitEvent(h->h2d_stream, h->ev_in_free, 0));             // the previous batch has been unpacked
    size_t off = 0;
    for (int q = 0; q < n; q++) {
        CUDA_OK(cudaMemcpyAsync((char*)h->stage_in + off, src[q], (size_t)fs[q]->host_count * sizeof(real), cudaMemcpyHostToDevice, h->h2d_stream));
        off += ((size_t)fs[q]->host_count * sizeof(real) + 255) / 256 * 256;
    }
    cudaEvent_t ready = h->ev_in_ready[h->n_in % H::XFER_RING];
    CUDA_OK(cudaEventRecord(ready, h->h2d_stream));
    CUDA_OK(cudaStreamWaitEvent(h->stream, ready, 0));
    off = 0;
    for (int q = 0; q < n; q++) {
        FieldRec* f = fs[q];
        real* dst = (real*)f->d[levels[q] - 1];
        const real* st = (const real*)((char*)h->stage_in + off);
        if (f->inner == IN_NL1_ME) h->zb_dirty = true;
        if (is_padded(f)) launch_pad(h, f, dst, st);
        else CUDA_OK(cudaMemcpyAsync(dst, st, (size_t)f->host_count * sizeof(real), cudaMemcpyDeviceToDevice, h->stream));
        off += ((size_t)f->host_count * sizeof(real) + 255) / 256 * 256;
    }
    CUDA_OK(cudaEventRecord(h->ev_in_free, h->stream));
    h->n_in++;
    return 0;
}
// lag = 0: every upload enqueued so far has left the host arrays; lag = k: all but the last k batches
extern "C" int mpasb_wait_uploads_lag(mpasb_handle h, int lag) {
    cudaSetDevice(h->device);
    const long b = h->n_in - 1 - lag;
    if (lag < 0 || lag >= H::XFER_RING) { h->err = "mpasb_wait_uploads_lag: lag out of range"; return 1; }
    if (b >= 0) CUDA_OK(cudaEventSynchronize(h->ev_in_ready[b % H::XFER_RING]));
    return 0;
}
extern "C" int mpasb_wait_uploads(mpasb_handle h) { return mpasb_wait_uploads_lag(h, 0); }
// Device fields -> host arrays, queued behind everything already enqueued on the compute stream (e.g. the step); the host
// arrays are valid after mpasb_wait_downloads.
extern "C" int mpasb_get_fields_async(mpasb_handle h, int n, const char* const* names, const int* levels, mpasb_real* const* dst, const long* counts) {
    cudaSetDevice(h->device);
    std::vector<FieldRec*> fs; size_t bytes = 0;
    if (int rc = batch_fields(h, n, names, levels, counts, fs, bytes)) return rc;
    if (ensure_transfer_state(h, 0, bytes)) return 3;
    if (h->n_out) CUDA_OK(cudaStreamWaitEvent(h->stream, h->ev_out_done[(h->n_out - 1) % H::XFER_RING], 0));     // the previous batch has left the staging area
    size_t off = 0;
    for (int q = 0; q < n; q++) {
        FieldRec* f = fs[q];
        const real* srcp = (const real*)f->d[levels[q] - 1];
        real* st = (real*)((char*)h->stage_out + off);
        if (is_padded(f)) launch_unpad(h, f, st, srcp);
        else CUDA_OK(cudaMemcpyAsync(st, srcp, (size_t)f->host_count * sizeof(real), cudaMemcpyDeviceToDevice, h->stream));
        off += ((size_t)f->host_count * sizeof(real) + 255) / 256 * 256;
    }
    CUDA_OK(cudaEventRecord(h->ev_out_ready, h->stream));
    CUDA_OK(cudaStreamWaitEvent(h->d2h_stream, h->ev_out_ready, 0));
    off = 0;
    for (int q = 0; q < n; q++) {
        CUDA_OK(cudaMemcpyAsync(dst[q], (char*)h->stage_out + off, (size_t)fs[q]->host_count * sizeof(real), cudaMemcpyDeviceToHost, h->d2h_stream));
        off += ((size_t)fs[q]->host_count * sizeof(real) + 255) / 256 * 256;
    }
    CUDA_OK(cudaEventRecord(h->ev_out_done[h->n_out % H::XFER_RING], h->d2h_stream));
    h->n_out++;
    return 0;
}
extern "C" int mpasb_wait_downloads_lag(mpasb_handle h, int lag) {
    cudaSetDevice(h->device);
    const long b = h->n_out - 1 - lag;
    if (lag < 0 || lag >= H::XFER_RING) { h->err = "mpasb_wait_downloads_lag: lag out of range"; return 1; }
    if (b >= 0) CUDA_OK(cudaEventSynchronize(h->ev_out_done[b % H::XFER_RING]));
    return 0;
}
extern "C" int mpasb_wait_downloads(mpasb_handle h) { return mpasb_wait_downloads_lag(h, 0); }

extern "C" int mpasb_set_field(mpasb_handle h, const char* name, int time_level, const mpasb_real* src, long count) {
    cudaSetDevice(h->device);
    FieldRec* f = find_field(h, name); if (!f) return 1;
    if (f->type != T_REAL || time_level < 1 || time_level > f->levels || count != f->host_count) { h->err = std::string("bad set_field ") + name; return 2; }
    real* dst = (real*)f->d[time_level - 1];
    const int LDK = h->D.LDK;
    const size_t o = outer_of(h, f->loc);
    const bool padded = f->inner == IN_NL || f->inner == IN_NL1 || f->inner == IN_NL1_ME || f->inner == IN_S_NL || f->inner == IN_NL_TWO;
    if (f->inner == IN_NL1_ME) h->zb_dirty = true;
    if (!strcmp(name, "weightsOnEdge")) {
        if ((long)h->hc_weightsOnEdge.size() != count || memcmp(h->hc_weightsOnEdge.data(), src, count * sizeof(real))) { h->hc_weightsOnEdge.assign(src, src + count); h->cor_dirty = true; }
    }
    if (!strcmp(name, "adv_coefs") || !strcmp(name, "adv_coefs_3rd")) {
        std::vector<real>& hc = !strcmp(name, "adv_coefs") ? h->hc_adv_coefs : h->hc_adv_coefs_3rd;
        if ((long)hc.size() != count || memcmp(hc.data(), src, count * sizeof(real))) { hc.assign(src, src + count); h->rings_dirty = true; }
    }
    if (!padded) { CUDA_OK(cudaMemcpyAsync(dst, src, count * sizeof(real), cudaMemcpyHostToDevice, h->stream)); }
    else {
        real* st = (real*)h->staging;
        CUDA_OK(cudaMemcpyAsync(st, src, count * sizeof(real), cudaMemcpyHostToDevice, h->stream));
        const int n1 = inner_dense1(h, f->inner);
        if (f->inner == IN_NL || f->inner == IN_NL1) k_pad<<<nblk(o * LDK), 256, 0, h->stream>>>(dst, st, o, n1, LDK);
        else if (f->inner == IN_NL1_ME) k_pad<<<nblk(o * h->dims.maxEdges * LDK), 256, 0, h->stream>>>(dst, st, o * h->dims.maxEdges, n1, LDK);
        else if (f->inner == IN_S_NL) k_pad_planes<<<nblk(f->dev_count), 256, 0, h->stream>>>(dst, st, o, h->dims.nVertLevels, h->dims.num_scalars, LDK);
        else k_pad_midplanes<<<nblk(f->dev_count), 256, 0, h->stream>>>(dst, st, o, h->dims.nVertLevels, 2, LDK);
        h->launches++;
    }
    CUDA_OK(cudaStreamSynchronize(h->stream));      // the host buffer and the staging area may be reused on return
    return 0;
}

extern "C" int mpasb_get_field(mpasb_handle h, const char* name, int time_level, mpasb_real* dstp, long count) {
    cudaSetDevice(h->device);
    FieldRec* f = find_field(h, name); if (!f) return 1;
    if (f->type != T_REAL || time_level < 1 || time_level > f->levels || count != f->host_count) { h->err = std::string("bad get_field ") + name; return 2; }
    const real* src = (const real*)f->d[time_level - 1];
    const int LDK = h->D.LDK;
    const size_t o = outer_of(h, f->loc);
    const bool padded = f->inner == IN_NL || f->inner == IN_NL1 || f->inner == IN_NL1_ME || f->inner == IN_S_NL || f->inner == IN_NL_TWO;
    if (!padded) { CUDA_OK(cudaMemcpyAsync(dstp, src, count * sizeof(real), cudaMemcpyDeviceToHost, h->stream)); }
    else {
        real* st = (real*)h->staging;
        const int n1 = inner_dense1(h, f->inner);
        if (f->inner == IN_NL || f->inner == IN_NL1) k_unpad<<<nblk(count), 256, 0, h->stream>>>(st, src, o, n1, LDK);
        else if (f->inner == IN_NL1_ME) k_unpad<<<nblk(count), 256, 0, h->stream>>>(st, src, o * h->dims.maxEdges, n1, LDK);
        else if (f->inner == IN_S_NL) k_unpad_planes<<<nblk(count), 256, 0, h->stream>>>(st, src, o, h->dims.nVertLevels, h->dims.num_scalars, LDK);
        else k_unpad_midplanes<<<nblk(count), 256, 0, h->stream>>>(st, src, o, h->dims.nVertLevels, 2, LDK);
        h->launches++;
        CUDA_OK(cudaMemcpyAsync(dstp, st, count * sizeof(real), cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int mpasb_set_field_int(mpasb_handle h, const char* name, const int* src, long count) {
    cudaSetDevice(h->device);
    FieldRec* f = find_field(h, name); if (!f) return 1;
    if (f->type != T_INT || count != f->host_count) { h->err = std::string("bad set_field_int ") + name; return 2; }
    int* st = (int*)h->staging;
    if (!strcmp(name, "nEdgesOnCell")) {
        h->max_ne = 0;
        for (long n = 0; n < count; n++) h->max_ne = std::max(h->max_ne, src[n]);
        // regional runs: the column-warp kernels carry the bdyMask / specZoneMask branches except the block-tiled column solve of
        // the strict path, so a strict regional handle (the bit-exact cross-check) stays on the generic family
        const bool regional_ok = !h->cfg.config_apply_lbcs || (h->relaxed && !getenv("MPASB_NO_SCAN"));
        h->colwarp = h->D.LDK <= 64 && h->max_ne <= CW_MAXNE && h->dims.maxEdges >= CW_NE && !getenv("MPASB_GENERIC_KERNELS") && regional_ok;
    }
    if (!strcmp(name, "advCellsForEdge")) { h->hc_advCells.assign(src, src + count); h->tiles_dirty = true; }
    if (!strcmp(name, "nAdvCellsForEdge")) { h->hc_nAdv.assign(src, src + count); h->tiles_dirty = true; }
    if (!strcmp(name, "cellsOnEdge")) { h->hc_cellsOnEdge.assign(src, src + count); h->tiles_dirty = true; }
    {   // host copies for build_flux_rings; an identical re-upload does not invalidate the tables
        std::vector<int>* hc = !strcmp(name, "cellsOnCell") ? &h->hc_cellsOnCell : !strcmp(name, "edgesOnCell") ? &h->hc_edgesOnCell :
                               !strcmp(name, "nEdgesOnCell") ? &h->hc_nEdgesOnCell : nullptr;
        if (!strcmp(name, "edgesOnEdge")) { h->hc_edgesOnEdge.assign(src, src + count); h->cor_dirty = true; }
        if (!strcmp(name, "nEdgesOnEdge")) { h->hc_nEdgesOnEdge.assign(src, src + count); h->cor_dirty = true; }
        if (hc || !strcmp(name, "cellsOnEdge")) h->cor_dirty = true;
        const bool ring_input = hc || !strcmp(name, "cellsOnEdge") || !strcmp(name, "advCellsForEdge") || !strcmp(name, "nAdvCellsForEdge");
        if (hc && ((long)hc->size() != count || memcmp(hc->data(), src, count * sizeof(int)))) { hc->assign(src, src + count); h->rings_dirty = true; }
        else if (ring_input && !hc) h->rings_dirty = true;
    }
    CUDA_OK(cudaMemcpyAsync(st, src, count * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    k_int_to_zero_based<<<nblk(count), 256, 0, h->stream>>>((int*)f->d[0], st, (size_t)count, f->target == TG_NONE ? 0 : 1);
    h->launches++;
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int mpasb_shift_time_levels(mpasb_handle h) {     // mpas_pool_shift_time_levels, shift_time_levs_array.inc:26-33
    // the state pool only (mpas_atm_core.F:808); the two "levels" of the lbc_* fields are tendency and state, not times
    for (FieldRec& f : h->fields) if (f.levels == 2 && strncmp(f.name, "lbc_", 4) != 0) { std::swap(f.d[0], f.d[1]); set_dev_ptr(h, f); }
    h->halo.parity ^= 1;        // halo plans cache field pointers per time-level parity
    return 0;
}

// ------------------------------------------------------------------ launch helpers
struct Scope {      // per-routine CUDA-event timing when profiling is on (timer names follow mpas_timer_start, TI:1045...)
    H* h; const char* name;
    Scope(H* h_, const char* n) : h(h_), name(n) { if (h->profile) cudaEventRecord(h->ev0, h->stream); }
    ~Scope() {
        if (!h->profile) return;
        cudaEventRecord(h->ev1, h->stream); cudaEventSynchronize(h->ev1);
        float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        ProfRec& p = h->prof[name]; p.ms += ms; p.count++;
    }
};
#define GRID(n) dim3((unsigned)(((n) + h->cpb - 1) / h->cpb)), dim3(h->D.LDK, h->cpb)
// In profile mode every launch is bracketed by CUDA events on the launching stream and
// accounted under the kernel's name ("k:<kernel>"); otherwise launches are fully asynchronous.
struct KScope {
    H* h; const char* name;
    KScope(H* h_, const char* n) : h(h_), name(n) { if (h->profile) cudaEventRecord(h->kev0, h->stream); }
    ~KScope() {
        if (!h->profile) return;
        cudaEventRecord(h->kev1, h->stream); cudaEventSynchronize(h->kev1);
        float ms = 0; cudaEventElapsedTime(&ms, h->kev0, h->kev1);
        ProfRec& p = h->prof[name]; p.ms += ms; p.count++;
    }
};
#define LAUNCH(kern, n, smem, ...) do { KScope ks_(h, "k:" #kern); kern<<<GRID(n), smem, h->stream>>>(__VA_ARGS__); h->launches++; } while (0)
// Launch of a kernel that contains pdl_wait() (every column-warp kernel, k_segments): with programmatic dependent launch its
// blocks may become resident while the previous kernel of the stream drains (mpasb_dev.cuh); MPASB_PDL=0 launches them plainly.
template <typename... P, typename... A>
static inline void klaunch(H* h, void (*kern)(P...), dim3 grid, dim3 block, size_t smem, A&&... args) {
    // MPASB_CARVEOUT=<percent>: one shared-memory carve-out preference for every kernel launched through here (experiment: do
    // carve-out changes between consecutive kernels cost anything?  the k9 arguments keep their own 100 %)
    static const int carve = getenv("MPASB_CARVEOUT") ? atoi(getenv("MPASB_CARVEOUT")) : -1;
    if (carve >= 0) {
        static std::map<const void*, bool> seen;
        if (!seen.count((const void*)kern)) { cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve); seen[(const void*)kern] = true; }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization; at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at; cfg.numAttrs = h->pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);
}
#define LAUNCHW(kern, n, ...) do { KScope ks_(h, "k:" #kern); klaunch(h, kern, dim3((unsigned)(((n) + CW_WARPS - 1) / CW_WARPS)), dim3(CW_THREADS), 0, __VA_ARGS__); h->launches++; } while (0)
#define LAUNCHWB(kern, W, n, ...) do { KScope ks_(h, "k:" #kern); klaunch(h, kern, dim3((unsigned)(((n) + (W) - 1) / (W))), dim3((W) * 32), 0, __VA_ARGS__); h->launches++; } while (0)
#define LAUNCHWY(kern, n, ny, ...) do { KScope ks_(h, "k:" #kern); klaunch(h, kern, dim3((unsigned)(((n) + CW_WARPS - 1) / CW_WARPS), (unsigned)(ny)), dim3(CW_THREADS), 0, __VA_ARGS__); h->launches++; } while (0)
#define LAUNCH1D(kern, n, ...) do { KScope ks_(h, "k:" #kern); kern<<<nblk(n), 256, 0, h->stream>>>(__VA_ARGS__); h->launches++; } while (0)

static const real rgas = RGAS, cp = CP_, rv = RV_;


// ------------------------------------------------------------------ routines
struct SegBuilder {       // collects the column ranges of one routine into one k_segments launch
    H* h; SegList L; int n = 0;
    explicit SegBuilder(H* h_) : h(h_) { L.scale = 1.0; }
    void add(real* d, const real* s, size_t ncols, int op = 0) {
        if (n == SEG_MAX) flush();
        L.seg[n].d = d; L.seg[n].s = const_cast<real*>(s); L.seg[n].n2 = (unsigned)(ncols * h->D.LDK / 2); L.seg[n].op = op; n++;
    }
    void flush() {
        if (!n) return;
        KScope ks_(h, "k:k_segments");
        klaunch(h, k_segments, dim3(SEG_BLOCKS, n), dim3(256), 0, L);
        h->launches++; n = 0;
    }
};
static int exchange(H* h, const char* group);
static int exchange_async(H* h, const char* group);
static void comm_wait(H* h);
static void rk_integration_setup(H* h) {       // TI:1930-2039
    Scope sc(h, "atm_rk_integration_setup");
    Dev& D = h->D; const size_t nC = D.nCells, nE = D.nEdges;
    SegBuilder sb(h);
    sb.add(D.ru_save, D.ru, nE); sb.add(D.u_2, D.u, nE);
    sb.add(D.rtheta_p_save, D.rtheta_p, nC); sb.add(D.rho_p_save, D.rho_p, nC);
    sb.add(D.theta_m_2, D.theta_m, nC); sb.add(D.rho_zz_2, D.rho_zz, nC);
    sb.add(D.rho_zz_old_split, D.rho_zz, nC);
    sb.add(D.rw_save, D.rw, nC); sb.add(D.w_2, D.w, nC);
    for (int s = 0; s < D.num_scalars; s++) sb.add(D.scalars_2 + s * D.cellPlane, D.scalars + s * D.cellPlane, nC);
    sb.flush();
    cudaMemsetAsync(D.theta_m_2 + nC * D.LDK, 0, D.LDK * sizeof(real), h->stream);          // TI:1987
}
static void compute_moist_coefficients(H* h) { // TI:2042-2146
    Scope sc(h, "atm_compute_moist_coefficients");
    LAUNCH(k_moist_cell, h->D.nCells, 0, h->D);
    LAUNCH(k_moist_edge, h->D.nEdges, 0, h->D);
}
static void compute_vert_imp_coefs(H* h, real dts) {   // TI:2225-2366
    Scope sc(h, "atm_compute_vert_imp_coefs");
    const real dtseps = .5 * dts * (1. + h->cfg.config_epssm);
    const real rcv = rgas / (cp - rgas);
    const real c2 = cp * rcv;
    if (h->colwarp) {
        const size_t smem3 = (size_t)3 * VIC_COLS * (h->D.LDK | 1) * sizeof(real);
        if (!h->smem_attr_vic) { cudaFuncSetAttribute(k3_vert_imp_coefs, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); h->smem_attr_vic = true; }
        KScope ks_(h, "k:k3_vert_imp_coefs");
        klaunch(h, k3_vert_imp_coefs, dim3((unsigned)((h->D.nCellsSolve + VIC_COLS - 1) / VIC_COLS)), dim3(VIC_WARPS * 32), smem3, h->D, dtseps, c2, rcv);
        h->launches++;
        return;
    }
    const size_t smem = (size_t)9 * h->D.LDK * h->cpb * sizeof(real);
    LAUNCH(k_vert_imp_coefs, h->D.nCellsSolve, smem, h->D, dtseps, c2, rcv);
}
static DynTendArgs dyn_tend_args(H* h, int rk_step, real dt) {
    const mpasb_config& c = h->cfg;
    DynTendArgs A; memset(&A, 0, sizeof(A));
    A.rk_step = rk_step; A.smag = c.config_horiz_mixing == 0;
    A.invDt = 1.0 / dt;
    const real len = c.config_len_disp;
    A.cs_len2 = (c.config_smagorinsky_coef * len) * (c.config_smagorinsky_coef * len);
    A.kdiff_cap = (0.01 * (len * len)) * A.invDt;
    A.fixed_visc2 = c.config_h_theta_eddy_visc2;
    if (A.smag) { A.h_mom_eddy_visc4 = c.config_visc4_2dsmag * (len * len * len); A.h_theta_eddy_visc4 = A.h_mom_eddy_visc4; }
    else { A.h_mom_eddy_visc4 = c.config_h_mom_eddy_visc4; A.h_theta_eddy_visc4 = c.config_h_theta_eddy_visc4; }
    A.del4u_div_factor = c.config_del4u_div_factor;
    A.v_mom_eddy_visc2 = c.config_v_mom_eddy_visc2; A.v_theta_eddy_visc2 = c.config_v_theta_eddy_visc2;
    A.coef_3rd_order = c.config_coef_3rd_order; A.prandtl_inv = 1.0 / PRANDTL; A.mix_full = c.config_mix_full;
    A.cam_coef = c.config_mpas_cam_coef; A.len_disp = len; A.n_cam_levels = c.config_number_cam_damping_levels;
    A.rayleigh_damp_u = c.config_rayleigh_damp_u; A.n_rayleigh_levels = c.config_number_rayleigh_damp_u_levels;
    if (A.rayleigh_damp_u)
        A.rayleigh_coef_inverse = 1.0 / ((real)A.n_rayleigh_levels * (c.config_rayleigh_damp_u_timescale_days * 86400.0));
    return A;
}
// Per block of EF_EB consecutive edges: the distinct stencil cells of its active edges (advCellsForEdge of edges that
// touch an owned cell), covered by runs of consecutive cell indices (gaps of <= EF_GAP unused cells are bridged), and
// for every stencil entry the shared-memory slot of its cell.  Derived from the host's own connectivity at upload;
// tiles that need more than EF_MAXR runs or EF_MAXT staged columns are marked -1 (global gathers).
static void build_flux_tiles(H* h) {
    h->tiles_dirty = false; h->tiles_ok = false;
    const int nE = h->dims.nEdges, nC = h->dims.nCells;
    if ((long)h->hc_advCells.size() != (long)(nE + 1) * 15 || (long)h->hc_nAdv.size() != nE + 1 || (long)h->hc_cellsOnEdge.size() != (long)(nE + 1) * 2) return;
    // Opt-in (MPASB_TMA_FLUX=1): measured on B200 the staged kernel takes 113 us against 109 us for the L1 gathers of
    // k2_dt_edge_flux -- shared-memory reads go through the same 64 B/clk/SM LSU data pipe that bounds the gathers
    // (profiles/r1_ncu_step_metrics_u.csv: 73 % pipe busy), so staging does not lift the bound.
    if ((h->D.LDK * sizeof(real)) % 16 != 0 || !getenv("MPASB_TMA_FLUX")) return;          // bulk copies move 16-byte multiples
    const int nTiles = (nE + EF_EB - 1) / EF_EB;
    std::vector<int4> hdr(nTiles, make_int4(0, 0, 0, 0));
    std::vector<int4> runs((size_t)nTiles * EF_MAXR, make_int4(0, 0, 0, 0));
    std::vector<int> stamp(nC + 1, -1), slotof(nC + 1, 0), tmp;
    std::vector<unsigned char> slot((size_t)nTiles * EF_EB * 15, 0xff);      // 0xff: beyond the stencil / inactive edge
    for (int t = 0; t < nTiles; t++) {
        tmp.clear();
        const int e0 = t * EF_EB, e1 = std::min(nE, (t + 1) * EF_EB);
        auto active = [&](int e) {
            const int c1 = h->hc_cellsOnEdge[2 * (size_t)e] - 1, c2 = h->hc_cellsOnEdge[2 * (size_t)e + 1] - 1;      // host lists are 1-based
            return c1 < h->dims.nCellsSolve || c2 < h->dims.nCellsSolve;
        };
        for (int e = e0; e < e1; e++) {
            if (!active(e)) continue;
            for (int j = 0; j < h->hc_nAdv[e]; j++) {
                const int c = h->hc_advCells[(size_t)e * 15 + j] - 1;
                if (c < 0 || c > nC) return;                                             // malformed list: keep the gather kernel
                if (stamp[c] != t) { stamp[c] = t; tmp.push_back(c); }
            }
        }
        if (tmp.empty()) continue;
        unsigned mask = 0;
        for (int e = e0; e < e1; e++) if (active(e)) mask |= 1u << (e - e0);
        std::sort(tmp.begin(), tmp.end());
        int nr = 0, total = 0; bool ok = true;
        int4* R = &runs[(size_t)t * EF_MAXR];
        for (size_t n = 0; n < tmp.size();) {
            size_t m = n;
            while (m + 1 < tmp.size() && tmp[m + 1] - tmp[m] <= EF_GAP + 1) m++;
            const int ncols = tmp[m] - tmp[n] + 1;
            if (nr == EF_MAXR || total + ncols > EF_MAXT) { ok = false; break; }
            R[nr++] = make_int4(tmp[n], ncols, total, 0);
            for (size_t q = n; q <= m; q++) slotof[tmp[q]] = total + (tmp[q] - tmp[n]);
            total += ncols; n = m + 1;
        }
        hdr[t] = ok ? make_int4(nr, total, (int)mask, 0) : make_int4(-1, 0, (int)mask, 0);
        for (int e = e0; e < e1; e++) {
            if (!active(e)) continue;
            for (int j = 0; j < h->hc_nAdv[e]; j++) slot[(size_t)e * 15 + j] = ok ? (unsigned char)slotof[h->hc_advCells[(size_t)e * 15 + j] - 1] : (unsigned char)0;
        }
    }
    if (h->d_tile_hdr) { cudaFree(h->d_tile_hdr); cudaFree(h->d_tile_runs); cudaFree(h->d_tile_slot); h->d_tile_hdr = nullptr; }
    if (cudaMalloc(&h->d_tile_hdr, hdr.size() * sizeof(int4)) != cudaSuccess || cudaMalloc(&h->d_tile_runs, runs.size() * sizeof(int4)) != cudaSuccess ||
        cudaMalloc(&h->d_tile_slot, slot.size()) != cudaSuccess) return;
    cudaMemcpyAsync(h->d_tile_hdr, hdr.data(), hdr.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(h->d_tile_runs, runs.data(), runs.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(h->d_tile_slot, slot.data(), slot.size(), cudaMemcpyHostToDevice, h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return;
    cudaFuncSetAttribute(k4_dt_edge_flux, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * EF_MAXT * h->D.LDK * sizeof(real)));
    h->tiles_ok = true;
}
// Canonical two-ring neighbourhoods for the cell-centred flux sweep (k5_flux_cell).  For an owned hexagon c whose six
// neighbours n_0..n_5 (= cellsOnCell order) are hexagons too, the twelve cells of the second ring are named by walking each
// neighbour's own cellsOnCell list away from c: m_{2i} lies straight beyond n_i, m_{2i-1} and m_{2i+1} to its sides, shared
// with n_{i-1} / n_{i+1}.  The 10-cell stencil of edge i (advCellsForEdge, mpas_atm_core.F:1285-1430) is then exactly
// {c, n_0..n_5, m_{2i-1}, m_{2i}, m_{2i+1}}; its weights adv_coefs / adv_coefs_3rd are re-ordered into that slot order.
// Every identity is CHECKED against the host's lists; a cell for which any of them fails (pentagons and heptagons, their
// neighbours, lists that are not consistently oriented, stencils of another size) is flagged irregular and the kernel
// walks its advCellsForEdge lists as the reference does.  Built once per mesh upload.
static void build_flux_rings(H* h) {
    h->rings_dirty = false; h->rings_ok = false; h->n_regular = 0;
    const int nC = h->dims.nCells, nE = h->dims.nEdges, nCS = h->dims.nCellsSolve, mx = h->dims.maxEdges;
    if (!h->relaxed || mx < 6 || nCS <= 0) return;
    if ((long)h->hc_cellsOnCell.size() != (long)(nC + 1) * mx || (long)h->hc_edgesOnCell.size() != (long)(nC + 1) * mx ||
        (long)h->hc_nEdgesOnCell.size() != nC + 1 || (long)h->hc_cellsOnEdge.size() != (long)(nE + 1) * 2 ||
        (long)h->hc_nAdv.size() != nE + 1 || (long)h->hc_advCells.size() != (long)(nE + 1) * 15 ||
        (long)h->hc_adv_coefs.size() != (long)(nE + 1) * 15 || (long)h->hc_adv_coefs_3rd.size() != (long)(nE + 1) * 15) return;
    std::vector<int> ring((size_t)nCS * FX_RING, 0);
    std::vector<real> wts((size_t)nCS * FX_WTS, (real)0);
    auto coc = [&](int c, int i) { return h->hc_cellsOnCell[(size_t)c * mx + i] - 1; };        // host lists are 1-based
    auto ne = [&](int c) { return h->hc_nEdgesOnCell[c]; };
    for (int c = 0; c < nCS; c++) {
        int* R = &ring[(size_t)c * FX_RING];
        real* W = &wts[(size_t)c * FX_WTS];
        if (ne(c) != 6) continue;
        int n[6], m[12]; bool ok = true;
        for (int i = 0; i < 6 && ok; i++) { n[i] = coc(c, i); ok = n[i] >= 0 && n[i] < nC && ne(n[i]) == 6; }
        for (int q = 0; q < 12; q++) m[q] = -1;
        int orient = 0;                                     // +1: neighbour lists run the same way round as c's, -1: the other way
        for (int i = 0; i < 6 && ok; i++) {
            int p = -1;
            for (int q = 0; q < 6; q++) if (coc(n[i], q) == c) p = q;
            if (p < 0) { ok = false; break; }
            int L[5];
            for (int q = 0; q < 5; q++) L[q] = coc(n[i], (p + 1 + q) % 6);
            const int prev = n[(i + 5) % 6], next = n[(i + 1) % 6];
            int o = 0;
            if (L[0] == prev && L[4] == next) o = 1; else if (L[0] == next && L[4] == prev) o = -1; else { ok = false; break; }
            if (orient == 0) orient = o; else if (orient != o) { ok = false; break; }
            const int side_a = o == 1 ? L[1] : L[3], side_b = o == 1 ? L[3] : L[1];      // m_{2i-1}, m_{2i+1}
            const int ia = (2 * i + 11) % 12, ib = (2 * i + 1) % 12;
            if (m[ia] >= 0 && m[ia] != side_a) ok = false;
            if (m[ib] >= 0 && m[ib] != side_b) ok = false;
            m[ia] = side_a; m[2 * i] = L[2]; m[ib] = side_b;
        }
        for (int q = 0; q < 12 && ok; q++) ok = m[q] >= 0 && m[q] < nC;
        for (int i = 0; i < 6 && ok; i++) {
            const int e = h->hc_edgesOnCell[(size_t)c * mx + i] - 1;
            if (e < 0 || e >= nE || h->hc_nAdv[e] != 10) { ok = false; break; }
            const int c1 = h->hc_cellsOnEdge[2 * (size_t)e] - 1, c2 = h->hc_cellsOnEdge[2 * (size_t)e + 1] - 1;
            if (!((c1 == c && c2 == n[i]) || (c2 == c && c1 == n[i]))) { ok = false; break; }
            int slot_cell[10] = {c, n[0], n[1], n[2], n[3], n[4], n[5], m[(2 * i + 11) % 12], m[2 * i], m[(2 * i + 1) % 12]};
            bool used[10] = {false};
            for (int j = 0; j < 10 && ok; j++) {
                const int cj = h->hc_advCells[(size_t)e * 15 + j] - 1;
                int sl = -1;
                for (int q = 0; q < 10; q++) if (slot_cell[q] == cj && !used[q]) sl = q;
                if (sl < 0) { ok = false; break; }
                used[sl] = true;
                W[i * 10 + sl] = h->hc_adv_coefs[(size_t)e * 15 + j];
                W[FX_WTS / 2 + i * 10 + sl] = h->hc_adv_coefs_3rd[(size_t)e * 15 + j];
            }
        }
        if (!ok) { for (int q = 0; q < FX_WTS; q++) W[q] = 0; continue; }
        for (int i = 0; i < 6; i++) R[i] = n[i];
        for (int q = 0; q < 12; q++) R[6 + q] = m[q];
        R[18] = 1;
        h->n_regular++;
    }
    Dev& D = h->D;
    for (void* p : {(void*)D.fx_ring, (void*)D.fx_w, (void*)D.hdiv_w, (void*)D.hdiv_theta}) if (p) cudaFree(p);
    D.fx_ring = nullptr; D.fx_w = nullptr; D.hdiv_w = nullptr; D.hdiv_theta = nullptr;
    if (cudaMalloc(&D.fx_ring, ring.size() * sizeof(int)) != cudaSuccess || cudaMalloc(&D.fx_w, wts.size() * sizeof(real)) != cudaSuccess ||
        cudaMalloc(&D.hdiv_w, D.cellPlane * sizeof(real)) != cudaSuccess || cudaMalloc(&D.hdiv_theta, D.cellPlane * sizeof(real)) != cudaSuccess) return;
    // on the compute stream, which is a non-blocking stream: a plain cudaMemcpy from pageable memory may return before its
    // DMA has landed and would not be ordered with the kernels that follow on h->stream
    cudaMemcpyAsync(D.fx_ring, ring.data(), ring.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(D.fx_w, wts.data(), wts.size() * sizeof(real), cudaMemcpyHostToDevice, h->stream);
    cudaMemsetAsync(D.hdiv_w, 0, D.cellPlane * sizeof(real), h->stream); cudaMemsetAsync(D.hdiv_theta, 0, D.cellPlane * sizeof(real), h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return;          // the host vectors go out of scope
    h->rings_ok = true;
}
// Tables for the cell-centred evaluation of the nonlinear Coriolis sum (TI:5418-5428).  The edgesOnEdge list of an edge e is the
// other edges of its two cells, so sum_j w_j u_j (pv_e + pv_j)/2 splits into one partial sum per adjacent cell; a cell that
// holds u and pv_edge of its ne edges in registers produces the partial sums of all ne of them from 2 ne gathered columns
// (k8_coriolis_cell), against 2 (ne1 + ne2 - 2) per edge in the edge-centred loop.  Every list is checked: if any edge with two
// cells inside the block has an edgesOnEdge entry that is not an edge of one of its cells, or entries missing, the edge-centred
// kernel stays in charge for the whole block.
static void build_coriolis_tables(H* h) {
    h->cor_dirty = false; h->cor_ok = false;
    const int nC = h->dims.nCells, nE = h->dims.nEdges, mx = h->dims.maxEdges, mx2 = h->dims.maxEdges2;
    // opt-in (MPASB_COR=1): measured on B200 (x1.40962 x 55) the edge kernel drops from 1.68 to 1.25 ms/step, but the cell kernel
    // that writes the partial sums costs 0.63 ms/step (12 C of extra HBM traffic per call): 12.34 vs 12.12 ms per step
    if (!h->relaxed || !getenv("MPASB_COR") || mx > CW_MAXNE) return;
    if ((long)h->hc_edgesOnCell.size() != (long)(nC + 1) * mx || (long)h->hc_nEdgesOnCell.size() != nC + 1 ||
        (long)h->hc_cellsOnEdge.size() != (long)(nE + 1) * 2 || (long)h->hc_edgesOnEdge.size() != (long)(nE + 1) * mx2 ||
        (long)h->hc_nEdgesOnEdge.size() != nE + 1 || (long)h->hc_weightsOnEdge.size() != (long)(nE + 1) * mx2) return;
    std::vector<real> W((size_t)(nC + 1) * 64, (real)0);
    std::vector<int> slot((size_t)nE + 1, 0);
    auto eoc = [&](int c, int i) { return h->hc_edgesOnCell[(size_t)c * mx + i] - 1; };
    auto slot_of = [&](int c, int e) { for (int i = 0; i < h->hc_nEdgesOnCell[c]; i++) if (eoc(c, i) == e) return i; return -1; };
    for (int e = 0; e < nE; e++) {
        const int c1 = h->hc_cellsOnEdge[2 * (size_t)e] - 1, c2 = h->hc_cellsOnEdge[2 * (size_t)e + 1] - 1;
        if (c1 < 0 || c1 >= nC || c2 < 0 || c2 >= nC) continue;            // an edge on the rim of the block: never a "solve" edge
        const int s1 = slot_of(c1, e), s2 = slot_of(c2, e);
        if (s1 < 0 || s2 < 0) { if (e < h->dims.nEdgesSolve) return; continue; }
        slot[e] = s1 | (s2 << 8);
        const int n = h->hc_nEdgesOnEdge[e];
        int found = 0;
        for (int j = 0; j < n; j++) {
            const int eoe = h->hc_edgesOnEdge[(size_t)e * mx2 + j] - 1;
            const real w = h->hc_weightsOnEdge[(size_t)e * mx2 + j];
            const int j1 = slot_of(c1, eoe), j2 = slot_of(c2, eoe);
            if (eoe == e || (j1 < 0) == (j2 < 0)) { if (e < h->dims.nEdgesSolve) return; found = -1000; break; }    // not exactly one owner cell
            if (j1 >= 0) W[(size_t)c1 * 64 + s1 * 8 + j1] = w; else W[(size_t)c2 * 64 + s2 * 8 + j2] = w;
            found++;
        }
        if (found != h->hc_nEdgesOnCell[c1] + h->hc_nEdgesOnCell[c2] - 2 && e < h->dims.nEdgesSolve) return;
    }
    Dev& D = h->D;
    for (void* p : {(void*)D.cor_w, (void*)D.cor_slot, (void*)D.cor_part}) if (p) cudaFree(p);
    D.cor_w = nullptr; D.cor_slot = nullptr; D.cor_part = nullptr;
    if (cudaMalloc(&D.cor_w, W.size() * sizeof(real)) != cudaSuccess || cudaMalloc(&D.cor_slot, slot.size() * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&D.cor_part, (size_t)(nC + 1) * mx * D.LDK * sizeof(real)) != cudaSuccess) return;
    cudaMemcpyAsync(D.cor_w, W.data(), W.size() * sizeof(real), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(D.cor_slot, slot.data(), slot.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    cudaMemsetAsync(D.cor_part, 0, (size_t)(nC + 1) * mx * D.LDK * sizeof(real), h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return;
    h->cor_ok = true;
}
// in_step: called from srk3, where (a) the exchange of w, pv_edge, rho_edge of the previous stage may still be in flight
// while the first kernel (which reads none of them) runs, and (b) tend_u is final before the w/theta tendencies are
// computed, so its exchange (TI:1228) is started here and overlaps them
static int compute_dyn_tend(H* h, int rk_step, real dt, bool in_step = false) {    // TI:4982-6240
    Scope sc(h, "atm_compute_dyn_tend");
    const Dev& D = h->D;
    const DynTendArgs A = dyn_tend_args(h, rk_step, dt);
    if (h->colwarp && !(A.cam_coef > 0.0)) LAUNCHW(k2_dt_cell_a, D.nCells, D, A);
    else LAUNCH(k_dt_cell_a, D.nCells, 0, D, A);
    comm_wait(h);
    if (h->colwarp && !A.rayleigh_damp_u) {
        if (h->cor_dirty) build_coriolis_tables(h);
        if (h->cor_ok) { LAUNCHW(k8_coriolis_cell, D.nCells, D); LAUNCHWB(k2_dt_edge_b<true>, EB_WARPS, D.nEdges, D, A); }
        else LAUNCHWB(k2_dt_edge_b<false>, EB_WARPS, D.nEdges, D, A);
    }
    else LAUNCH(k_dt_edge_b, D.nEdges, 0, D, A);
    if (rk_step == 1) {
        if (A.h_mom_eddy_visc4 > 0.0) {
            if (h->colwarp) { LAUNCHW(k2_dt_delsq_vertex, D.nVertices, D); LAUNCHW(k2_dt_delsq_cell, D.nCells, D); }
            else { LAUNCH(k_dt_delsq_vertex, D.nVertices, 0, D); LAUNCH(k_dt_delsq_cell, D.nCells, 0, D); }
        }
        if (h->colwarp && !(A.v_mom_eddy_visc2 > 0.0) && !A.rayleigh_damp_u) LAUNCHW(k2_dt_edge_d, D.nEdgesSolve, D, A);
        else LAUNCH(k_dt_edge_d, D.nEdgesSolve, 0, D, A);
    }
    if (in_step && exchange_async(h, "dynamics:tend_u")) return 1;
    if (rk_step == 1) {
        if (h->colwarp) LAUNCHW(k2_dt_cell_e, D.nCells, D, A);
        else LAUNCH(k_dt_cell_e, D.nCells, 0, D, A);
    }
    if (h->colwarp && !(A.v_mom_eddy_visc2 > 0.0) && !(A.v_theta_eddy_visc2 > 0.0)) {
        if (h->rings_dirty) build_flux_rings(h);
        if (h->rings_ok) {
            // relaxed arithmetic: one cell-centred sweep computes the horizontal flux divergence of w and theta_m with the
            // two-ring neighbourhood in registers; no per-edge flux arrays, 38 instead of 60 gathered columns per cell
            static const bool one_field_per_warp = getenv("MPASB_FLUX_SPLIT") != nullptr;      // measured slower: 1.40 vs 0.96 ms/step
            if (one_field_per_warp) {
                KScope ks_(h, "k:k5s_flux_cell");
                klaunch(h, k5s_flux_cell, dim3((unsigned)((2 * (size_t)D.nCellsSolve + CW_WARPS - 1) / CW_WARPS)), dim3(CW_THREADS), 0, D);
                h->launches++;
            } else {
                KScope ks_(h, "k:k5_flux_cell");
                static int sm_count = 0;
                if (!sm_count) cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
                const unsigned need = (unsigned)((D.nCellsSolve + FX_WARPS - 1) / FX_WARPS), resident = (unsigned)(sm_count * FX_MINB);
                klaunch(h, k5_flux_cell, dim3(std::min(need, resident)), dim3(FX_WARPS * 32), 0, D);       // persistent warps
                h->launches++;
            }
            static const bool old_f = getenv("MPASB_OLD_CELL_F") != nullptr;
            if (old_f) LAUNCHW(k2_dt_cell_f<true>, D.nCellsSolve, D, A);
            else {
                static int sm_count = 0;
                if (!sm_count) cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
                static const int variant = getenv("MPASB_CF7") ? atoi(getenv("MPASB_CF7")) : 0;
                KScope ks_(h, "k:k7_dt_cell_f");
#define CF7_LAUNCH(W, MB) do { const unsigned need = (unsigned)((D.nCellsSolve + (W) - 1) / (W)), resident = (unsigned)(sm_count * (MB)); \
                    klaunch(h, k7_dt_cell_f<W, MB>, dim3(std::min(need, resident)), dim3((W) * 32), 0, D, A); } while (0)
                if (variant == 1) CF7_LAUNCH(4, 3);            // 168 registers, 12 warps per SM
                else CF7_LAUNCH(8, 2);                         // 128 registers, 16 warps per SM
#undef CF7_LAUNCH
                h->launches++;
            }
        } else {
            if (h->tiles_dirty) build_flux_tiles(h);
            if (h->tiles_ok) {
                KScope ks_(h, "k:k4_dt_edge_flux");
                klaunch(h, k4_dt_edge_flux, dim3((unsigned)((D.nEdges + EF_EB - 1) / EF_EB)), dim3(CW_THREADS), 2 * EF_MAXT * D.LDK * sizeof(real),
                        D, (const int4*)h->d_tile_hdr, (const int4*)h->d_tile_runs, (const unsigned char*)h->d_tile_slot);
                h->launches++;
            }
            else LAUNCHW(k2_dt_edge_flux, D.nEdges, D);
            static const bool split_f = getenv("MPASB_SPLIT_CELL_F") != nullptr;
            if (split_f) { LAUNCHW(k2_dt_cell_fw, D.nCellsSolve, D, A); LAUNCHW(k2_dt_cell_ft, D.nCellsSolve, D, A); }
            else LAUNCHW(k2_dt_cell_f<false>, D.nCellsSolve, D, A);
        }
    }
    else LAUNCH(k_dt_cell_f, D.nCellsSolve, 0, D, A);
    return 0;
}
static void refresh_zb_flags(H* h) {
    if (!h->zb_dirty) return;
    k_zb_flags<<<nblk((size_t)h->D.nCells + 1), 256, 0, h->stream>>>(h->D);
    h->launches++; h->zb_dirty = false;
}
static void set_smlstep_pert_variables(H* h) {               // TI:2427-2508
    Scope sc(h, "small_step_prep");
    refresh_zb_flags(h);
    if (h->colwarp) LAUNCHW(k2_smlstep_pert, h->D.nCellsSolve, h->D);
    else LAUNCH(k_smlstep_pert, h->D.nCellsSolve, 0, h->D);
}
// Columns of the cell solve that an exchange of cell fields touches -- every owned cell on a send list, and the halo cells,
// whose old rtheta_pp must be saved before the exchange overwrites it (TI:2827-2842) -- and all the others.
static void build_acoustic_lists(H* h) {
    const HaloKind& K = h->halo.kind[0];
    const int nC = h->D.nCells, nS = h->D.nCellsSolve;
    std::vector<char> mark(nC, 0);
    for (int c : K.h_send) if (c >= 0 && c < nC) mark[c] = 1;
    for (int c = nS; c < nC; c++) mark[c] = 1;
    std::vector<int> bnd, inter;
    for (int c = 0; c < nC; c++) (mark[c] ? bnd : inter).push_back(c);
    if (h->d_ac_bnd) { cudaFree(h->d_ac_bnd); cudaFree(h->d_ac_int); }
    cudaMalloc(&h->d_ac_bnd, std::max<size_t>(1, bnd.size()) * sizeof(int)); cudaMalloc(&h->d_ac_int, std::max<size_t>(1, inter.size()) * sizeof(int));
    cudaMemcpyAsync(h->d_ac_bnd, bnd.data(), bnd.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(h->d_ac_int, inter.data(), inter.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    cudaStreamSynchronize(h->stream);
    h->n_ac_bnd = (int)bnd.size(); h->n_ac_int = (int)inter.size(); h->ac_lists_ok = true;
}
// group != nullptr: the exchange of the cell fields this step produces (TI:1279/1302) is issued here as well -- between the
// boundary and the interior columns of the cell solve where that is possible, after the routine otherwise
static int advance_acoustic_step(H* h, real dts, int small_step, const char* group = nullptr) {    // TI:2646-2984
    comm_wait(h);
    Scope sc(h, "atm_advance_acoustic_step");
    const real epssm = h->cfg.config_epssm;
    const real rcv = rgas / (cp - rgas);
    const real c2 = cp * rcv;
    const real resm = (1.0 - epssm) / (1.0 + epssm);
    if (h->colwarp) {
        // first small step: ru_p = dts * tend_u is evaluated on the fly by the cell kernel and written by the
        // following divergence-damping kernel (h->ru_p_pending), saving one pass over three edge arrays
        if (small_step == 1) h->ru_p_pending = true;
        else {
            const int dd_mode = h->dd_deferred ? (h->ru_p_pending ? 2 : 1) : 0;
            LAUNCHW(k2_acoustic_edge, h->D.nEdges, h->D, dts, c2, dd_mode, h->dd_coef);
            h->dd_deferred = false; h->ru_p_pending = false;
        }
        static const bool no_scan = getenv("MPASB_NO_SCAN") != nullptr;
        if (h->relaxed && !no_scan) {       // the column solve as a warp-level prefix of affine maps: one warp per column, registers only
            static int sm_count = 0;
            if (!sm_count) cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
            static const int variant = getenv("MPASB_AC6") ? atoi(getenv("MPASB_AC6")) : 1;
            KScope ks_(h, "k:k6_acoustic_cell");
#define AC6_LAUNCH(W, MB, LISTED, REG, LIST, N) do { const unsigned need = (unsigned)(((N) + (W) - 1) / (W)), resident = (unsigned)(sm_count * (MB)); \
                klaunch(h, k6_acoustic_cell<W, MB, LISTED, REG>, dim3(std::max(1u, std::min(need, resident))), dim3((W) * 32), 0, h->D, dts, small_step, epssm, resm, (const int*)(LIST), (int)(N)); \
                h->launches++; } while (0)
            const bool regional = h->D.apply_lbcs != 0;
            // the plain kernel in three register / occupancy variants; the listed and the regional forms as (4, 3) only
#define AC6_RUN(LIST, N) do { const bool listed = (LIST) != nullptr; \
                              if (listed && regional) AC6_LAUNCH(4, 3, true, true, LIST, N); \
                              else if (listed) AC6_LAUNCH(4, 3, true, false, LIST, N); \
                              else if (regional) AC6_LAUNCH(4, 3, false, true, LIST, N); \
                              else if (variant == 0) AC6_LAUNCH(8, 2, false, false, LIST, N);      /* 128 registers, 16 warps per SM */ \
                              else if (variant == 2) AC6_LAUNCH(4, 4, false, false, LIST, N); /* 128 registers, 16 warps per SM in smaller blocks */ \
                              else AC6_LAUNCH(4, 3, false, false, LIST, N); } while (0)       /* 168 registers, 12 warps per SM */
            // opt-in (MPASB_SPLIT=1): measured on 4 B200 it does not pay -- 13.73 / 13.85 ms per step against 13.60 ms with the
            // exchange simply following the kernel: the two launches lose more (a scattered boundary launch, two tails) than
            // the ~20 us of exchange they hide (profiles/r2_bench_t_4gpu_*)
            static const bool split = getenv("MPASB_SPLIT") != nullptr && atoi(getenv("MPASB_SPLIT")) != 0;
            if (group && h->halo.active && h->overlap && !h->profile && split) {
                // boundary columns first; their exchange travels while the interior columns are solved
                if (!h->ac_lists_ok) build_acoustic_lists(h);
                AC6_RUN(h->d_ac_bnd, h->n_ac_bnd);
                if (exchange_async(h, group)) return 1;
                AC6_RUN(h->d_ac_int, h->n_ac_int);
                return 0;
            }
            // TMA-pipelined form (own-column operands by cp.async.bulk, one trip ahead): plain runs only
            static const int tma = getenv("MPASB_AC9") ? atoi(getenv("MPASB_AC9")) : AC9_DEFAULT;
            if (tma && !regional) {
                const size_t smem9 = (size_t)2 * AC9_NF * AC9_WARPS * h->D.LDK * sizeof(real);
                if (!h->smem_attr_ac9) {
                    cudaFuncSetAttribute(k9_acoustic_cell<AC9_WARPS, AC9_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::min<size_t>(smem9, 220 * 1024 / AC9_MINB));
                    cudaFuncSetAttribute(k9_acoustic_cell<AC9_WARPS, AC9_MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
                    h->smem_attr_ac9 = true;
                }
                if (smem9 <= (size_t)220 * 1024 / AC9_MINB) {
                    const unsigned need = (unsigned)((h->D.nCells + AC9_WARPS - 1) / AC9_WARPS), resident = (unsigned)(sm_count * AC9_MINB);
                    KScope ks9_(h, "k:k9_acoustic_cell");
                    klaunch(h, k9_acoustic_cell<AC9_WARPS, AC9_MINB>, dim3(std::max(1u, std::min(need, resident))), dim3(AC9_WARPS * 32), smem9,
                            h->D, dts, small_step, epssm, resm);
                    h->launches++;
                    return group ? exchange(h, group) : 0;
                }
            }
            AC6_RUN((const int*)nullptr, h->D.nCells);
#undef AC6_RUN
#undef AC6_LAUNCH
            return group ? exchange(h, group) : 0;
        }
        const size_t smem3 = (size_t)AC3_ARRAYS * AC3_COLS * (h->D.LDK | 1) * sizeof(real);
        if (!h->smem_attr_ac) { cudaFuncSetAttribute(k3_acoustic_cell, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); h->smem_attr_ac = true; }
        KScope ks_(h, "k:k3_acoustic_cell");
        klaunch(h, k3_acoustic_cell, dim3((unsigned)((h->D.nCells + AC3_COLS - 1) / AC3_COLS)), dim3(AC3_WARPS * 32), smem3, h->D, dts, small_step, epssm, resm);
        h->launches++;
        return group ? exchange(h, group) : 0;
    }
    LAUNCH(k_acoustic_edge, h->D.nEdges, 0, h->D, dts, small_step, c2);
    const size_t smem = (size_t)6 * h->D.LDK * h->cpb * sizeof(real);
    LAUNCH(k_acoustic_cell, h->D.nCells, smem, h->D, dts, small_step, epssm, resm);
    if (h->D.apply_lbcs) LAUNCH(k_lbc_acoustic_spec, h->D.nCellsSolve, 0, h->D, dts, small_step, epssm);      // TI:2962-2971
    return group ? exchange(h, group) : 0;
}
// defer: the next kernel that reads ru_p (the edge update of the next small step, or -- single block -- the edge part of
// recover_large_step_variables) applies the damping in registers; same arithmetic, one kernel and one ru_p round trip less
// Owned edges on a send list of the edge kind (any layer) and their flags: see k2_recover_edge
static void build_dd_lists(H* h) {
    const HaloKind& K = h->halo.kind[1];
    const int nE = h->D.nEdges;
    std::vector<unsigned char> done(nE + 1, 0);
    for (int e : K.h_send) if (e >= 0 && e < nE) done[e] = 1;
    std::vector<int> list;
    for (int e = 0; e < nE; e++) if (done[e]) list.push_back(e);
    if (h->d_dd_edges) { cudaFree(h->d_dd_edges); cudaFree(h->d_dd_done); }
    cudaMalloc(&h->d_dd_edges, std::max<size_t>(1, list.size()) * sizeof(int)); cudaMalloc(&h->d_dd_done, done.size());
    cudaMemcpyAsync(h->d_dd_edges, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(h->d_dd_done, done.data(), done.size(), cudaMemcpyHostToDevice, h->stream);
    cudaStreamSynchronize(h->stream);
    h->n_dd_edges = (int)list.size(); h->dd_lists_ok = true;
}
// partial: decomposed block, the exchange of ru_p follows (TI:1322)
static void divergence_damping_3d(H* h, real dts, bool defer = false, bool partial = false) {           // TI:2987-3075
    Scope sc(h, "atm_divergence_damping_3d");
    const real rdts = 1.0 / dts;
    const real coef_divdamp = 2.0 * h->cfg.config_smdiv * h->cfg.config_len_disp * rdts;
    if (h->colwarp && defer && h->fuse_dd) { h->dd_deferred = true; h->dd_coef = coef_divdamp; h->dd_dts = dts; return; }
    comm_wait(h);                      // the rtheta_pp halos of an exchange started inside advance_acoustic_step
    if (h->colwarp && partial && h->fuse_dd) {
        // decomposed block, last small step of a stage: damp (and, on a first small step, materialise) only the edges a neighbour
        // is about to receive (TI:1322); k2_recover_edge applies the same arithmetic to every other edge in registers
        if (!h->dd_lists_ok) build_dd_lists(h);
        if (h->n_dd_edges) LAUNCHW(k2_divergence_damping, h->n_dd_edges, h->D, coef_divdamp, h->ru_p_pending ? 1 : 0, dts, (const int*)h->d_dd_edges, h->n_dd_edges);
        h->dd_deferred = true; h->dd_partial = true; h->dd_coef = coef_divdamp; h->dd_dts = dts;      // ru_p_pending stays as it is: mode 2 for the folded edges
        return;
    }
    if (h->colwarp) {
        LAUNCHW(k2_divergence_damping, h->D.nEdges, h->D, coef_divdamp, h->ru_p_pending ? 1 : 0, dts, (const int*)nullptr, 0);
        h->ru_p_pending = false;
        return;
    }
    LAUNCH(k_divergence_damping, h->D.nEdges, 0, h->D, coef_divdamp);
}
// in_step: u (time level 2) is final after the edge kernel, so its layer-3 exchange (TI:1371) is started before the
// kernel that finishes w and overlaps it
static int recover_large_step_variables(H* h, real dt, int ns, int rk_step, bool in_step = false) {    // TI:3189-3431
    Scope sc(h, "atm_recover_large_step_variables");
    const real rcv = rgas / (cp - rgas);
    const real p0 = 1.0e+05;
    const real invNs = 1 / (real)ns;
    refresh_zb_flags(h);
    if (h->colwarp) {
        LAUNCHW(k2_recover_cell1, h->D.nCells + 1, h->D, dt, invNs, rk_step, rcv, rgas / p0);
        {
            const int dd_mode = h->dd_deferred ? (h->ru_p_pending ? 2 : 1) : 0;
            LAUNCHW(k2_recover_edge, h->D.nEdges, h->D, invNs, dd_mode, h->dd_coef, h->dd_dts, (const unsigned char*)(h->dd_partial ? h->d_dd_done : nullptr));
            h->dd_deferred = false; h->ru_p_pending = false; h->dd_partial = false;
        }
    } else {
        LAUNCH(k_recover_cell1, h->D.nCells + 1, 0, h->D, dt, invNs, rk_step, rcv, rgas / p0);
        LAUNCH(k_recover_edge, h->D.nEdges, 0, h->D, invNs);
    }
    if (in_step && exchange_async(h, "dynamics:u_3")) return 1;
    if (h->colwarp) LAUNCHW(k2_recover_cell2, h->D.nCells, h->D, h->cfg.cf1, h->cfg.cf2, h->cfg.cf3);
    else LAUNCH(k_recover_cell2, h->D.nCells, 0, h->D, h->cfg.cf1, h->cfg.cf2, h->cfg.cf3);
    return 0;
}
static void compute_solve_diagnostics(H* h, real dt, int time_lev, int rk_step) {  // TI:6337-6773
    Scope sc(h, "atm_compute_solve_diagnostics");
    const Dev& D = h->D;
    const real* u = time_lev == 2 ? D.u_2 : D.u;
    const real* hh = time_lev == 2 ? D.rho_zz_2 : D.rho_zz;
    const int apvm = h->cfg.config_apvm_upwinding > 0.0;
    const int reconstruct_v = (rk_step == 0 || rk_step == 3);
    if (h->colwarp) {
        LAUNCHW(k2_diag_vertex, D.nVertices, D, u);
        LAUNCHW(k2_diag_cell, D.nCells, D, u, apvm);
        LAUNCHW(k2_diag_edge, D.nEdges, D, u, hh, reconstruct_v, apvm, h->cfg.config_apvm_upwinding * dt);
        return;
    }
    LAUNCH(k_diag_vertex, D.nVertices, 0, D, u);
    LAUNCH(k_diag_cell, D.nCells, 0, D, u, apvm);
    LAUNCH(k_diag_edge, D.nEdges, 0, D, u, hh, reconstruct_v, apvm, h->cfg.config_apvm_upwinding * dt);
}
static void rk_dynamics_substep_finish(H* h, int dynamics_substep, int dynamics_split) {   // TI:7013-7191
    Scope sc(h, "atm_rk_dynamics_substep_finish");
    Dev& D = h->D; const size_t nC = D.nCells, nE = D.nEdges, L = D.LDK;
    cudaMemsetAsync(D.theta_m + nC * L, 0, L * sizeof(real), h->stream);                    // TI:7082
    SegBuilder sb(h);
    sb.L.scale = 1.0 / (real)dynamics_split;                                                // inv_dynamics_split
    if (dynamics_substep < dynamics_split) {
        sb.add(D.ru_save, D.ru, nE); sb.add(D.u, D.u_2, nE);
        sb.add(D.rtheta_p_save, D.rtheta_p, nC); sb.add(D.rho_p_save, D.rho_p, nC);
        sb.add(D.theta_m, D.theta_m_2, nC); sb.add(D.rho_zz, D.rho_zz_2, nC);
        sb.add(D.rw_save, D.rw, nC); sb.add(D.w, D.w_2, nC);
    }
    // *_split = (first substep ? Avg : *_split + Avg); on the last substep also Avg = *_split * inv_dynamics_split
    const int op = (dynamics_substep == 1 ? 0 : 1) + (dynamics_substep == dynamics_split ? 2 : 0);
    sb.add(D.ruAvg_split, D.ruAvg, nE, op); sb.add(D.wwAvg_split, D.wwAvg, nC, op);
    if (dynamics_substep == dynamics_split) sb.add(D.rho_zz, D.rho_zz_old_split, nC);
    sb.flush();
}
static void advance_scalars(H* h, real dt, int rk_step) {     // TI:3575-3855
    Scope sc(h, "atm_advance_scalars");
    const mpasb_config& c = h->cfg;
    real weight_time_new = 1.;
    if (c.config_split_dynamics_transport) {
        if ((rk_step == 1) && c.config_time_integration_order == 3) weight_time_new = 1. / 3;
        if ((rk_step == 1) && c.config_time_integration_order == 2) weight_time_new = 1. / 2;
        if (rk_step == 2) weight_time_new = 1. / 2;
        if (rk_step == 3) weight_time_new = 1.;
    }
    const real weight_time_old = 1. - weight_time_new;
    if (h->colwarp) {
        LAUNCHW(k2_scalars_edge, h->D.nEdges, h->D);
        LAUNCHW(k2_scalars_cell, h->D.nCellsSolve, h->D, dt, weight_time_old, weight_time_new, c.config_coef_3rd_order);
        return;
    }
    LAUNCH(k_scalars_edge, h->D.nEdges, 0, h->D);
    LAUNCH(k_scalars_cell, h->D.nCellsSolve, 0, h->D, dt, weight_time_old, weight_time_new, c.config_coef_3rd_order);
}
static int advance_scalars_mono(H* h, real dt) {              // TI:4012-4734
    Scope sc(h, "atm_advance_scalars_mono");
    const Dev& D = h->D;
    const bool adv_density = h->cfg.config_split_dynamics_transport != 0;
    const int S = D.num_scalars, last = D.mb_planes - 1;
    LAUNCH(k_mono_pre, D.nCellsSolve, 0, D, dt);
    if (exchange(h, "dynamics:scalars_old")) return 1;
    if (adv_density) { if (h->colwarp) LAUNCHW(k2_mono_rho_int, D.nCellsSolve, D, dt); else LAUNCH(k_mono_rho_int, D.nCellsSolve, 0, D, dt); }
    const real* rho = adv_density ? D.rho_zz_int : D.rho_zz_2;
    if (h->colwarp && D.mb_planes == S && S > 1) {
        // batched over scalars: the scalar loop of TI:4220 runs INSIDE each kernel launch (gridDim.y = scalar, work arrays
        // one plane per scalar), and the scale factors of all scalars travel in one exchange instead of S (TI:4568)
        LAUNCHWY(k2_mono_cell1, D.nCellsSolve, S, D, 0, -1, dt, h->cfg.config_coef_3rd_order);
        LAUNCHWY(k2_mono_edge2, D.nEdges, S, D, 0, -1, dt);
        LAUNCHWY(k2_mono_cell3, D.nCellsSolve, S, D, -1, rho);
        if (exchange(h, "dynamics:scale_all")) return 1;
        LAUNCHWY(k2_mono_edge4, D.nEdges, S, D, -1);
        LAUNCHWY(k2_mono_cell5, D.nCells, S, D, 0, -1, rho);
        return 0;
    }
    for (int s = 0; s < S; s++) {
        if (h->colwarp) LAUNCHW(k2_mono_cell1, D.nCellsSolve, D, s, last, dt, h->cfg.config_coef_3rd_order); else LAUNCH(k_mono_cell1, D.nCellsSolve, 0, D, s, dt, h->cfg.config_coef_3rd_order);
        if (h->colwarp) LAUNCHW(k2_mono_edge2, D.nEdges, D, s, last, dt); else LAUNCH(k_mono_edge2, D.nEdges, 0, D, s, dt);
        if (h->colwarp) LAUNCHW(k2_mono_cell3, D.nCellsSolve, D, last, rho); else LAUNCH(k_mono_cell3, D.nCellsSolve, 0, D, rho);
        if (exchange(h, "dynamics:scale")) return 1;
        if (h->colwarp) LAUNCHW(k2_mono_edge4, D.nEdges, D, last); else LAUNCH(k_mono_edge4, D.nEdges, 0, D);
        if (h->colwarp) LAUNCHW(k2_mono_cell5, D.nCells, D, s, last, rho); else LAUNCH(k_mono_cell5, D.nCells, 0, D, s, rho);
    }
    return 0;
}
// the same routine split at its two exchange points, for hosts that exchange halos themselves (one scalar at a time; the
// work arrays are the ones of the field table, i.e. the last plane)
static void mono_pre(H* h, real dt) { LAUNCH(k_mono_pre, h->D.nCellsSolve, 0, h->D, dt); }
static void mono_a(H* h, real dt, int s) {
    const Dev& D = h->D;
    const bool adv_density = h->cfg.config_split_dynamics_transport != 0;
    const int last = D.mb_planes - 1;
    if (s == 0 && adv_density) { if (h->colwarp) LAUNCHW(k2_mono_rho_int, D.nCellsSolve, D, dt); else LAUNCH(k_mono_rho_int, D.nCellsSolve, 0, D, dt); }
    const real* rho = adv_density ? D.rho_zz_int : D.rho_zz_2;
    if (h->colwarp) LAUNCHW(k2_mono_cell1, D.nCellsSolve, D, s, last, dt, h->cfg.config_coef_3rd_order); else LAUNCH(k_mono_cell1, D.nCellsSolve, 0, D, s, dt, h->cfg.config_coef_3rd_order);
    if (h->colwarp) LAUNCHW(k2_mono_edge2, D.nEdges, D, s, last, dt); else LAUNCH(k_mono_edge2, D.nEdges, 0, D, s, dt);
    if (h->colwarp) LAUNCHW(k2_mono_cell3, D.nCellsSolve, D, last, rho); else LAUNCH(k_mono_cell3, D.nCellsSolve, 0, D, rho);
}
static void mono_b(H* h, int s) {
    const Dev& D = h->D;
    const real* rho = h->cfg.config_split_dynamics_transport ? D.rho_zz_int : D.rho_zz_2;
    const int last = D.mb_planes - 1;
    if (h->colwarp) LAUNCHW(k2_mono_edge4, D.nEdges, D, last); else LAUNCH(k_mono_edge4, D.nEdges, 0, D);
    if (h->colwarp) LAUNCHW(k2_mono_cell5, D.nCells, D, s, last, rho); else LAUNCH(k_mono_cell5, D.nCells, 0, D, s, rho);
}
static void init_coupled_diagnostics(H* h) {                  // TI:6776-7010
    const real rcv = rgas / (cp - rgas);
    const real p0 = 1.e5;
    LAUNCH(k_initcd_cell1, h->D.nCells, 0, h->D, rv / rgas, rcv, rgas / p0);
    LAUNCH(k_initcd_edge, h->D.nEdges, 0, h->D);
    LAUNCH(k_initcd_cell2, h->D.nCells, 0, h->D);
}

// mpas_reconstruct_2d (mpas_vector_reconstruction.F:205-330): cell-centre velocity from the edge-normal component
static void reconstruct(H* h, int time_lev, int include_halos) {
    Scope sc(h, "mpas_reconstruct");
    const Dev& D = h->D;
    const real* u = time_lev == 2 ? D.u_2 : D.u;
    const int n = include_halos ? D.nCells : D.nCellsSolve;
    if (h->colwarp) LAUNCHW(k2_reconstruct, n, D, u, n, h->cfg.on_a_sphere);
    else LAUNCH(k_reconstruct, n, 0, D, u, n, h->cfg.on_a_sphere);
}
// atm_compute_output_diagnostics (mpas_atm_core.F:901-950): theta, rho, pressure for the history stream
static void compute_output_diagnostics(H* h, int time_lev) {
    Scope sc(h, "atm_compute_output_diagnostics");
    const Dev& D = h->D;
    const real* qv = (time_lev == 2 ? D.scalars_2 : D.scalars) + (size_t)D.index_qv * D.cellPlane;
    LAUNCH(k_output_diagnostics, D.nCells, 0, D, time_lev == 2 ? D.theta_m_2 : D.theta_m, time_lev == 2 ? D.rho_zz_2 : D.rho_zz, qv, rv / rgas);
}

// ------------------------------------------------------------------ regional runs (config_apply_lbcs), kernels_lbc.cuh
// delta_t as in mpas_atm_get_bdy_state (mpas_atm_boundaries.F:497-503): dt = (LBC_intv_end - currTime) - delta_t, in RKIND
static real lbc_dtl(H* h, real delta_t) { real dt = h->lbc_dt_end; dt = dt - delta_t; return dt; }
static void lbc_speczone_tend(H* h) {                              // TI:1218-1239
    Scope sc(h, "atm_bdy_adjust_dynamics_speczone_tend");
    LAUNCH(k_lbc_speczone_cell, h->D.nCellsSolve, 0, h->D);
    LAUNCH(k_lbc_speczone_edge, h->D.nEdgesSolve, 0, h->D);
}
static void lbc_relaxzone_tend(H* h, real time_dyn_step, real dt) {    // TI:1243-1267
    Scope sc(h, "atm_bdy_adjust_dynamics_relaxzone_tend");
    const real dtl = lbc_dtl(h, time_dyn_step);
    LAUNCH(k_lbc_relax_cell, h->D.nCellsSolve, 0, h->D, dt, dtl);
    LAUNCH(k_lbc_relax_edge, h->D.nEdges, 0, h->D, dt, dtl, (real)h->cfg.config_relax_zone_divdamp_coef);
}
static void lbc_reset_u_ru(H* h, real time_dyn_step) { LAUNCH(k_lbc_reset_u_ru, h->D.nEdges, 0, h->D, lbc_dtl(h, time_dyn_step)); }    // TI:1343-1388
static void lbc_adjust_scalars(H* h, real dt, real dt_rk) {        // TI:1413-1428, 1565-1580
    Scope sc(h, "atm_bdy_adjust_scalars");
    const real dtl = lbc_dtl(h, dt_rk);
    for (int s = 0; s < h->D.num_scalars; s++) {
        LAUNCH(k_lbc_adjust_scalars_a, h->D.nCellsSolve, 0, h->D, s, dt, dt_rk, dtl, h->D.scalar_new);
        LAUNCH(k_lbc_adjust_scalars_b, h->D.nCellsSolve, 0, h->D, s, h->D.scalar_new);
    }
}
static void lbc_zero_gradient_w(H* h) { LAUNCH(k_lbc_zero_w, h->D.nCellsSolve, 0, h->D); }                  // TI:1477-1484
static void lbc_reset_speczone_values(H* h, real dt) { LAUNCH(k_lbc_reset_speczone, h->D.nCellsSolve, 0, h->D, lbc_dtl(h, dt)); }   // TI:1676-1695
static void lbc_set_scalars(H* h, real dt) {                       // TI:1700-1719
    const real dtl = lbc_dtl(h, dt);
    for (int s = 0; s < h->D.num_scalars; s++) LAUNCH(k_lbc_set_scalars, h->D.nCellsSolve, 0, h->D, s, dtl);
}

// ------------------------------------------------------------------ halo exchange (mpas_halo.F:498-846)
#include "halo_host.inl"

// the compute stream waits for an exchange started with exchange_async
static void comm_wait(H* h) {
    if (!h->comm_pending) return;
    cudaStreamWaitEvent(h->stream, h->ev_done, 0);
    h->comm_pending = false;
}
static int exchange(H* h, const char* group) {
    if (!h->halo.active) return 0;                // single block: exchange lists are empty (mpas_dmpar.F:2074-2153)
    comm_wait(h);
    Scope sc(h, "exchange_halo_group");
    return halo_exchange(h, group, h->stream);
}
// Start an exchange whose fields are final but that the NEXT kernels neither read (halo) nor write: pack, send/recv and
// unpack run on comm_stream while the compute stream continues; comm_wait() is called before the first consumer.
static int exchange_async(H* h, const char* group) {
    if (!h->halo.active) return 0;
    if (!h->overlap || h->profile) return exchange(h, group);
    comm_wait(h);
    cudaEventRecord(h->ev_ready, h->stream);
    cudaStreamWaitEvent(h->comm_stream, h->ev_ready, 0);
    if (halo_exchange(h, group, h->comm_stream)) return 1;
    cudaEventRecord(h->ev_done, h->comm_stream);
    h->comm_pending = true;
    return 0;
}

// ------------------------------------------------------------------ atm_srk3  TI:803-1725
static int srk3(H* h, real dt) {
    const mpasb_config& c = h->cfg;
    Dev& D = h->D;
    cudaSetDevice(h->device);
    // TI:967-991, 1091-1093: the physics tendencies tend_ru_physics, tend_rtheta_physics, tend_rho_physics are zero from
    // mpasb_create on and hold whatever a host with physics uploads with mpasb_set_field (physics_get_tend fills them inside
    // atm_srk3 in the reference); mpasb_zero_physics_tendencies resets them
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
    } else if (c.config_time_integration_order == 2) {
        rk_timestep[1] = dt_dynamics / 2.; rk_timestep[2] = dt_dynamics / 2.; rk_timestep[3] = dt_dynamics;
        rk_sub_timestep[1] = rk_sub_timestep[2] = rk_sub_timestep[3] = dt_dynamics / (real)number_of_sub_steps;
        number_sub_steps[1] = std::max(1, number_of_sub_steps / 2); number_sub_steps[2] = std::max(1, number_of_sub_steps / 2); number_sub_steps[3] = number_of_sub_steps;
    } else { h->err = "config_time_integration_order must be 2 or 3"; return 1; }
    // config_split_dynamics_transport = false: the scalars are advanced inside the dynamics RK loop (TI:1404-1407)
    const bool coupled_transport = c.config_scalar_advection && !c.config_split_dynamics_transport;
    const bool lbcs = c.config_apply_lbcs != 0;               // regional run: TI:1218-1268, 1343-1430, 1477-1484, 1560-1582, 1676-1720
    auto advance_scalars_stage = [&](int rk_step, real dt_rk) -> int {        // advance_scalars, TI:1730-1927
        if (rk_step < 3 || (!c.config_monotonic && !c.config_positive_definite)) { advance_scalars(h, dt_rk, rk_step); return 0; }
        return advance_scalars_mono(h, dt_rk);
    };

    if (exchange(h, "dynamics:theta_m,scalars,pressure_p,rtheta_p")) return 1;
    rk_integration_setup(h);
    compute_moist_coefficients(h);
    for (int dynamics_substep = 1; dynamics_substep <= dynamics_split; dynamics_substep++) {
        compute_vert_imp_coefs(h, rk_sub_timestep[1]);
        if (exchange(h, "dynamics:exner")) return 1;
        for (int rk_step = 1; rk_step <= 3; rk_step++) {
            if (c.config_time_integration_order == 3 && rk_step == 2) compute_vert_imp_coefs(h, rk_sub_timestep[rk_step]);
            if (compute_dyn_tend(h, rk_step, dt, true)) return 1;        // starts the exchange of tend_u (TI:1228)
            comm_wait(h);
            const real time_dyn_step = dt_dynamics * (real)(dynamics_substep - 1) + rk_timestep[rk_step];        // TI:1246
            if (lbcs) { lbc_speczone_tend(h); lbc_relaxzone_tend(h, time_dyn_step, dt); }                          // TI:1218-1268
            set_smlstep_pert_variables(h);
            for (int small_step = 1; small_step <= number_sub_steps[rk_step]; small_step++) {
                // TI:1279 exchanges rho_pp before every acoustic step.  On the first small step nothing reads the
                // halo of rho_pp (the edge update is ru_p = dts * tend_u, TI:2798-2806), so that exchange is skipped;
                // on later ones it is merged into the rtheta_pp exchange (TI:1302) that directly precedes it.
                const bool more = small_step < number_sub_steps[rk_step];
                if (advance_acoustic_step(h, rk_sub_timestep[rk_step], small_step, more ? "dynamics:rtheta_pp,rho_pp" : "dynamics:rtheta_pp")) return 1;
                // (more small steps follow, or nothing exchanges ru_p: the damping is folded into the next edge kernel)
                static const bool dd_partial_ok = !getenv("MPASB_NO_DD_PARTIAL");
                divergence_damping_3d(h, rk_sub_timestep[rk_step], more || !h->halo.active, !more && h->halo.active && dd_partial_ok);
            }
            if (exchange(h, "dynamics:rw_p,ru_p,rho_pp,rtheta_pp")) return 1;
            if (recover_large_step_variables(h, rk_timestep[rk_step], number_sub_steps[rk_step], rk_step, !lbcs)) return 1;   // starts u_3 (TI:1371)
            comm_wait(h);
            if (lbcs) {                                           // TI:1343-1395: driving u, ru in the specified zone, then all three edge layers
                lbc_reset_u_ru(h, time_dyn_step);
                if (exchange(h, "dynamics:u_123")) return 1;
            }
            if (coupled_transport) {
                if (advance_scalars_stage(rk_step, rk_timestep[rk_step])) return 1;
                if (lbcs) {                                       // TI:1409-1430
                    if (exchange(h, "dynamics:scalars")) return 1;
                    lbc_adjust_scalars(h, dt, rk_timestep[rk_step]);
                }
            }
            compute_solve_diagnostics(h, dt, 2, rk_step);
            // TI:1424.  Stages 1 and 2: the next kernels (vertical coefficients, first cell kernel of the tendencies) read none
            // of the three fields, so the exchange overlaps them; after stage 3 the substep roll copies w and must wait
            const char* grp = coupled_transport ? "dynamics:w,pv_edge,rho_edge,scalars" : "dynamics:w,pv_edge,rho_edge";      // TI:1463-1473
            if (rk_step < 3 && !lbcs ? exchange_async(h, grp) : exchange(h, grp)) return 1;
            if (lbcs) { lbc_zero_gradient_w(h); if (exchange(h, "dynamics:w")) return 1; }       // TI:1477-1484
        }
        if (dynamics_substep < dynamics_split)
            if (exchange(h, "dynamics:theta_m,pressure_p,rtheta_p")) return 1;
        rk_dynamics_substep_finish(h, dynamics_substep, dynamics_split);
    }
    if (c.config_scalar_advection && c.config_split_dynamics_transport) {
        rk_timestep[1] = dt / 3.; rk_timestep[2] = dt / 2.; rk_timestep[3] = dt;
        if (c.config_time_integration_order == 2) rk_timestep[1] = dt / 2.;
        for (int rk_step = 1; rk_step <= 3; rk_step++) {
            if (advance_scalars_stage(rk_step, rk_timestep[rk_step])) return 1;
            if (lbcs) {                                           // TI:1560-1582
                if (exchange(h, "dynamics:scalars")) return 1;
                lbc_adjust_scalars(h, dt, rk_timestep[rk_step]);
            }
            if (rk_step < 3) if (exchange(h, "dynamics:scalars")) return 1;
        }
    }
    if (lbcs) {                                                   // TI:1676-1720
        lbc_reset_speczone_values(h, dt);
        if (exchange(h, "dynamics:scalars")) return 1;
        lbc_set_scalars(h, dt);
    }
    comm_wait(h);
    reconstruct(h, 2, 0);                         // TI:1596-1611: uReconstruct* from u (time level 2), owned cells
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { h->err = std::string("kernel launch: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

extern "C" int mpasb_zero_physics_tendencies(mpasb_handle h) {
    cudaSetDevice(h->device);
    const Dev& D = h->D;
    CUDA_OK(cudaMemsetAsync(D.tend_ru_physics, 0, D.edgePlane * sizeof(real), h->stream));
    CUDA_OK(cudaMemsetAsync(D.tend_rtheta_physics, 0, D.cellPlane * sizeof(real), h->stream));
    CUDA_OK(cudaMemsetAsync(D.tend_rho_physics, 0, D.cellPlane * sizeof(real), h->stream));
    return 0;
}

extern "C" int mpasb_step(mpasb_handle h, mpasb_real dt, int itimestep) { (void)itimestep; return srk3(h, dt); }

extern "C" int mpasb_minmax(mpasb_handle h, mpasb_real out[4]) {
    cudaSetDevice(h->device);
    const Dev& D = h->D;
    double tmp[4];
    CUDA_OK(cudaMemsetAsync(h->d_minmax, 0, 4 * sizeof(double), h->stream));
    k_minmax<<<296, 256, 0, h->stream>>>(D.w_2, D.nCellsSolve, D.nl, D.LDK, h->d_minmax);
    k_minmax<<<296, 256, 0, h->stream>>>(D.u_2, D.nEdgesSolve, D.nl, D.LDK, h->d_minmax + 2);
    h->launches += 2;
    CUDA_OK(cudaMemcpyAsync(tmp, h->d_minmax, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    for (int n = 0; n < 4; n++) out[n] = (mpasb_real)tmp[n];
    return 0;
}

// summarize_timestep without a host stall.  _async enqueues, behind the step on the compute stream,
// the min/max reductions of w and u (TI:8286-8319), of every scalar (config_print_global_minmax_sca, TI:8322-8342) and
// the NaN tests of w and u (TI:8258-8281), followed by one copy into pinned host memory; the host keeps going (e.g. it
// launches the next step) and _fetch waits for that copy only.  Layout of `minmax`: {min w, max w, min u, max u,
// min s1, max s1, ...}; nan_count = {NaNs in w, NaNs in u} over owned elements of time level 2.
extern "C" int mpasb_summarize_timestep_async(mpasb_handle h) {
    cudaSetDevice(h->device);
    const Dev& D = h->D;
    const int S = D.num_scalars, nval = 2 * (2 + S) + 2;
    if (h->summary_head - h->summary_tail >= H::SUMMARY_RING) { h->err = "mpasb_summarize_timestep_async: four summaries are already pending, fetch one first"; return 1; }
    const int q = (int)(h->summary_head % H::SUMMARY_RING);
    if (!h->d_summary[q]) {
        CUDA_OK(cudaMalloc(&h->d_summary[q], nval * sizeof(double)));
        CUDA_OK(cudaMallocHost(&h->h_summary[q], nval * sizeof(double)));
        CUDA_OK(cudaEventCreateWithFlags(&h->ev_summary[q], cudaEventDisableTiming));
    }
    double* ds = h->d_summary[q];
    CUDA_OK(cudaMemsetAsync(ds, 0, nval * sizeof(double), h->stream));
    unsigned long long* nan = reinterpret_cast<unsigned long long*>(ds + 2 * (2 + S));
    k_minmax_nan<<<296, 256, 0, h->stream>>>(D.w_2, D.nCellsSolve, D.nl, D.LDK, ds, nan);
    k_minmax_nan<<<296, 256, 0, h->stream>>>(D.u_2, D.nEdgesSolve, D.nl, D.LDK, ds + 2, nan + 1);
    for (int sc = 0; sc < S; sc++)
        k_minmax<<<296, 256, 0, h->stream>>>(D.scalars_2 + (size_t)sc * D.cellPlane, D.nCellsSolve, D.nl, D.LDK, ds + 4 + 2 * sc);
    h->launches += 2 + S;
    // The result goes to the pinned host buffer by stores from a kernel (pinned memory is device-accessible under unified
    // addressing), not by cudaMemcpyAsync: a copy-engine transfer on the compute stream would queue behind a bulk download
    // already running on the same engine (mpasb_get_fields_async) and hold the next step back for its whole duration --
    // measured: 17.6 instead of 13.1 ms per pipelined request.
    k_copy_to_host<<<1, 64, 0, h->stream>>>(h->h_summary[q], ds, nval);
    CUDA_OK(cudaEventRecord(h->ev_summary[q], h->stream));
    h->summary_head++;
    return 0;
}
extern "C" int mpasb_summarize_timestep_fetch(mpasb_handle h, mpasb_real* minmax, long n_minmax, long nan_count[2]) {
    cudaSetDevice(h->device);
    if (h->summary_head == h->summary_tail) { h->err = "mpasb_summarize_timestep_fetch without a pending mpasb_summarize_timestep_async"; return 1; }
    const int q = (int)(h->summary_tail % H::SUMMARY_RING);            // oldest pending summary
    const int S = h->D.num_scalars;
    if (n_minmax < 4 || n_minmax > 2 * (2 + S)) { h->err = "mpasb_summarize_timestep_fetch: n_minmax out of range"; return 2; }
    CUDA_OK(cudaEventSynchronize(h->ev_summary[q]));
    h->summary_tail++;
    for (long n = 0; n < n_minmax; n++) minmax[n] = (mpasb_real)h->h_summary[q][n];
    const unsigned long long* nan = reinterpret_cast<const unsigned long long*>(h->h_summary[q] + 2 * (2 + S));
    if (nan_count) { nan_count[0] = (long)nan[0]; nan_count[1] = (long)nan[1]; }
    return 0;
}

// CUDA-event timer on the stream the kernels are launched on (bench.py's timed region)
extern "C" int mpasb_timer_start(mpasb_handle h) { cudaSetDevice(h->device); CUDA_OK(cudaEventRecord(h->tev0, h->stream)); return 0; }
extern "C" int mpasb_timer_stop(mpasb_handle h, double* ms) {
    cudaSetDevice(h->device);
    CUDA_OK(cudaEventRecord(h->tev1, h->stream)); CUDA_OK(cudaEventSynchronize(h->tev1));
    float f = 0; CUDA_OK(cudaEventElapsedTime(&f, h->tev0, h->tev1)); *ms = f; return 0;
}
extern "C" int mpasb_set_profile(mpasb_handle h, int on) { h->profile = on != 0; if (on) h->prof.clear(); return 0; }
extern "C" int mpasb_get_profile(mpasb_handle h, char* buf, long buflen) {
    std::string s;
    for (auto& kv : h->prof) { char line[256]; snprintf(line, sizeof line, "%s %.6f %ld\n", kv.first.c_str(), kv.second.ms, kv.second.count); s += line; }
    strncpy(buf, s.c_str(), buflen - 1); buf[buflen - 1] = 0;
    return 0;
}

// ------------------------------------------------------------------ init-time and per-routine entry points
// ------------------------------------------------------------------ atm_mpas_init_block, mesh part (CORE:456-470, 1091-1452)
#include "init_block_host.inl"
static const char* const INITBLK_REAL_OUT[] = {"edgesOnVertex_sign", "edgesOnCell_sign", "zb_cell", "zb3_cell", "invAreaCell", "invDvEdge", "invDcEdge",
    "invAreaTriangle", "adv_coefs", "adv_coefs_3rd", "meshScalingDel2", "meshScalingDel4", "meshScalingRegionalCell", "meshScalingRegionalEdge", "dss",
    "coeffs_reconstruct"};          // the last one only when the six coordinate arrays were given (and the mesh is on a sphere)
static const char* const INITBLK_INT_OUT[] = {"kiteForCell", "nAdvCellsForEdge", "advCellsForEdge"};
static std::vector<real>* initblk_real(initblk::Out& o, const char* name) {
    std::vector<real>* v[] = {&o.edgesOnVertex_sign, &o.edgesOnCell_sign, &o.zb_cell, &o.zb3_cell, &o.invAreaCell, &o.invDvEdge, &o.invDcEdge,
        &o.invAreaTriangle, &o.adv_coefs, &o.adv_coefs_3rd, &o.meshScalingDel2, &o.meshScalingDel4, &o.meshScalingRegionalCell, &o.meshScalingRegionalEdge, &o.dss,
        &o.coeffs_reconstruct};
    for (size_t q = 0; q < sizeof(INITBLK_REAL_OUT) / sizeof(*INITBLK_REAL_OUT); q++) if (!strcmp(name, INITBLK_REAL_OUT[q])) return v[q];
    return nullptr;
}
static std::vector<int>* initblk_int(initblk::Out& o, const char* name) {
    std::vector<int>* v[] = {&o.kiteForCell, &o.nAdvCellsForEdge, &o.advCellsForEdge};
    for (size_t q = 0; q < sizeof(INITBLK_INT_OUT) / sizeof(*INITBLK_INT_OUT); q++) if (!strcmp(name, INITBLK_INT_OUT[q])) return v[q];
    return nullptr;
}
static bool initblk_dims_ok(const mpasb_dims* d) {
    return d && d->nCells > 0 && d->nEdges > 0 && d->nVertices > 0 && d->nVertLevels > 0 && d->maxEdges > 0 && d->maxEdges <= 14 && d->vertexDegree > 0;
}
extern "C" int mpasb_init_block_host(const mpasb_dims* dims, const mpasb_config* cfg, int config_h_ScaleWithMesh, double config_zd, double config_xnutr,
                                     int n_in, const char* const* in_names, const void* const* in_arrays,
                                     int n_out, const char* const* out_names, void* const* out_arrays) {
    if (!initblk_dims_ok(dims) || !cfg || n_in < 0 || n_out < 0 || (n_in && (!in_names || !in_arrays)) || (n_out && (!out_names || !out_arrays))) return 2;
    initblk::In in;
    if (initblk::bind(in, n_in, in_names, in_arrays)) return 1;                 // an input is missing
    initblk::Out o;
    initblk::compute(*dims, *cfg, config_h_ScaleWithMesh, config_zd, config_xnutr, in, o);
    if (in.coords() && cfg->on_a_sphere) initblk::reconstruct_coeffs(*dims, in, o);
    for (int k = 0; k < n_out; k++) {
        if (!out_names[k] || !out_arrays[k]) return 2;
        if (std::vector<real>* v = initblk_real(o, out_names[k])) { if (v->empty()) return 1; memcpy(out_arrays[k], v->data(), v->size() * sizeof(real)); }
        else if (std::vector<int>* w = initblk_int(o, out_names[k])) memcpy(out_arrays[k], w->data(), w->size() * sizeof(int));
        else return 1;                                                            // not a field this routine derives
    }
    return 0;
}
extern "C" int mpasb_init_block(mpasb_handle h, int config_h_ScaleWithMesh, double config_zd, double config_xnutr,
                                int n_in, const char* const* in_names, const void* const* in_arrays) {
    if (!h || n_in < 0 || (n_in && (!in_names || !in_arrays))) return 2;
    initblk::In in;
    if (const char* missing = initblk::bind(in, n_in, in_names, in_arrays)) { h->err = std::string("mpasb_init_block: input missing: ") + missing; return 1; }
    initblk::Out o;
    initblk::compute(h->dims, h->cfg, config_h_ScaleWithMesh, config_zd, config_xnutr, in, o);
    if (in.coords() && h->cfg.on_a_sphere) initblk::reconstruct_coeffs(h->dims, in, o);
    for (const char* name : INITBLK_REAL_OUT) {
        std::vector<real>* v = initblk_real(o, name);
        if (v->empty()) continue;                              // coeffs_reconstruct without coordinates: stays what the host uploaded
        if (int rc = mpasb_set_field(h, name, 1, v->data(), (long)v->size())) return rc;
    }
    for (const char* name : INITBLK_INT_OUT) { std::vector<int>* v = initblk_int(o, name); if (int rc = mpasb_set_field_int(h, name, v->data(), (long)v->size())) return rc; }
    return 0;
}

#define ENTRY(body) { cudaSetDevice(h->device); body; CUDA_OK(cudaStreamSynchronize(h->stream)); CUDA_OK(cudaGetLastError()); return 0; }
extern "C" int mpasb_init_coupled_diagnostics(mpasb_handle h) ENTRY(init_coupled_diagnostics(h))
extern "C" int mpasb_init_solve_diagnostics(mpasb_handle h, mpasb_real dt) ENTRY(compute_solve_diagnostics(h, dt, 1, 0))
// the same without the host synchronisation: for requests queued back to back (mpasb_set_fields_async ... mpasb_get_fields_async)
extern "C" int mpasb_init_solve_diagnostics_async(mpasb_handle h, mpasb_real dt) { cudaSetDevice(h->device); compute_solve_diagnostics(h, dt, 1, 0); return 0; }
extern "C" int mpasb_reconstruct(mpasb_handle h, int time_level, int include_halos) ENTRY(reconstruct(h, time_level, include_halos))
extern "C" int mpasb_compute_output_diagnostics(mpasb_handle h, int time_level) ENTRY(compute_output_diagnostics(h, time_level))
extern "C" int mpasb_set_lbc_time(mpasb_handle h, mpasb_real seconds_to_interval_end) { h->lbc_dt_end = seconds_to_interval_end; return 0; }
// the regional-path routines one at a time (parity tests); a, b: the real arguments of the reference call (see kernels_lbc.cuh)
extern "C" int mpasb_k_lbc(mpasb_handle h, const char* routine, mpasb_real a, mpasb_real b) {
    cudaSetDevice(h->device);
    const std::string r = routine;
    if (r == "speczone_tend") lbc_speczone_tend(h);
    else if (r == "relaxzone_tend") lbc_relaxzone_tend(h, a, b);          // a = time_dyn_step, b = dt
    else if (r == "reset_u_ru") lbc_reset_u_ru(h, a);                     // a = time_dyn_step
    else if (r == "adjust_scalars") lbc_adjust_scalars(h, a, b);          // a = dt, b = rk_timestep(rk_step)
    else if (r == "zero_gradient_w") lbc_zero_gradient_w(h);
    else if (r == "reset_speczone_values") lbc_reset_speczone_values(h, a);    // a = dt
    else if (r == "set_scalars") lbc_set_scalars(h, a);                   // a = dt
    else { h->err = "mpasb_k_lbc: unknown routine " + r; return 1; }
    CUDA_OK(cudaStreamSynchronize(h->stream)); CUDA_OK(cudaGetLastError());
    return 0;
}
extern "C" int mpasb_k_rk_integration_setup(mpasb_handle h) ENTRY(rk_integration_setup(h))
extern "C" int mpasb_k_compute_moist_coefficients(mpasb_handle h) ENTRY(compute_moist_coefficients(h))
extern "C" int mpasb_k_compute_vert_imp_coefs(mpasb_handle h, mpasb_real dts) ENTRY(compute_vert_imp_coefs(h, dts))
extern "C" int mpasb_k_compute_dyn_tend(mpasb_handle h, int rk_step, mpasb_real dt) ENTRY(if (compute_dyn_tend(h, rk_step, dt)) return 1)
extern "C" int mpasb_k_set_smlstep_pert_variables(mpasb_handle h) ENTRY(set_smlstep_pert_variables(h))
// called on its own the routine must leave ru_p/ruAvg as the reference does: materialise the deferred edge update
static void advance_acoustic_step_standalone(H* h, real dts, int small_step) {
    advance_acoustic_step(h, dts, small_step);
    if (h->ru_p_pending) { LAUNCH(k_acoustic_edge, h->D.nEdges, 0, h->D, dts, 1, (real)0); h->ru_p_pending = false; }
}
extern "C" int mpasb_k_advance_acoustic_step(mpasb_handle h, mpasb_real dts, int small_step) ENTRY(advance_acoustic_step_standalone(h, dts, small_step))
extern "C" int mpasb_k_divergence_damping_3d(mpasb_handle h, mpasb_real dts) ENTRY(divergence_damping_3d(h, dts))
extern "C" int mpasb_k_recover_large_step_variables(mpasb_handle h, mpasb_real dt, int ns, int rk_step) ENTRY(if (recover_large_step_variables(h, dt, ns, rk_step)) return 1)
extern "C" int mpasb_k_compute_solve_diagnostics(mpasb_handle h, mpasb_real dt, int rk_step) ENTRY(compute_solve_diagnostics(h, dt, 2, rk_step))
extern "C" int mpasb_k_rk_dynamics_substep_finish(mpasb_handle h, int s, int n) ENTRY(rk_dynamics_substep_finish(h, s, n))
extern "C" int mpasb_k_advance_scalars(mpasb_handle h, mpasb_real dt, int rk_step) ENTRY(advance_scalars(h, dt, rk_step))
extern "C" int mpasb_k_advance_scalars_mono(mpasb_handle h, mpasb_real dt) ENTRY(if (advance_scalars_mono(h, dt)) return 1)
extern "C" int mpasb_k_advance_scalars_mono_pre(mpasb_handle h, mpasb_real dt) ENTRY(mono_pre(h, dt))
extern "C" int mpasb_k_advance_scalars_mono_a(mpasb_handle h, mpasb_real dt, int s) ENTRY(mono_a(h, dt, s))
extern "C" int mpasb_k_advance_scalars_mono_b(mpasb_handle h, mpasb_real dt, int s) ENTRY(mono_b(h, s); (void)dt)
