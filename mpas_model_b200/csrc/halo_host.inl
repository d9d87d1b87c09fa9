// halo_host.inl -- GPU pack -> NCCL send/recv -> GPU unpack, replacing
// mpas_halo_exch_group_full_halo_exch (src/framework/mpas_halo.F:498-846).
// Message layout per neighbour follows the reference (MH:671,695): field-major, then halo
// layer, then list element, with the vertical index fastest.  NCCL is bound at run time
// with dlopen so that single-GPU use has no NCCL dependency.
#include <nccl.h>

// Group table: src/core_atmosphere/mpas_atm_halos.F:211-298 (mpas_halo back-end; identical to :90-167)
struct GroupField { const char* name; int lev; int kind; int layers; };     // layers: bit l-1 set = halo layer l
struct GroupDef { const char* name; std::vector<GroupField> fields; };
static const std::vector<GroupDef>& group_table() {
    static const std::vector<GroupDef> g = {
        {"dynamics:theta_m,scalars,pressure_p,rtheta_p", {{"theta_m", 1, 0, 3}, {"scalars", 1, 0, 3}, {"pressure_p", 1, 0, 3}, {"rtheta_p", 1, 0, 3}}},
        {"dynamics:rw_p,ru_p,rho_pp,rtheta_pp", {{"rw_p", 1, 0, 1}, {"ru_p", 1, 1, 2}, {"rho_pp", 1, 0, 3}, {"rtheta_pp", 1, 0, 2}}},
        {"dynamics:w,pv_edge,rho_edge", {{"w", 2, 0, 3}, {"pv_edge", 1, 1, 3}, {"rho_edge", 1, 1, 3}}},
        {"dynamics:w,pv_edge,rho_edge,scalars", {{"w", 2, 0, 3}, {"pv_edge", 1, 1, 3}, {"rho_edge", 1, 1, 3}, {"scalars", 2, 0, 3}}},
        {"dynamics:theta_m,pressure_p,rtheta_p", {{"theta_m", 2, 0, 3}, {"pressure_p", 1, 0, 3}, {"rtheta_p", 1, 0, 3}}},
        {"dynamics:exner", {{"exner", 1, 0, 3}}},
        {"dynamics:tend_u", {{"tend_u", 1, 1, 1}}},
        {"dynamics:rho_pp", {{"rho_pp", 1, 0, 1}}},
        {"dynamics:rtheta_pp", {{"rtheta_pp", 1, 0, 1}}},
        // not a reference group: "dynamics:rtheta_pp" of small step s merged with "dynamics:rho_pp" of small step s+1
        // (TI:1302 + TI:1279), which the reference issues back to back around atm_divergence_damping_3d
        {"dynamics:rtheta_pp,rho_pp", {{"rtheta_pp", 1, 0, 1}, {"rho_pp", 1, 0, 1}}},
        {"dynamics:u_123", {{"u", 2, 1, 7}}},
        {"dynamics:u_3", {{"u", 2, 1, 4}}},
        {"dynamics:scalars", {{"scalars", 2, 0, 3}}},
        {"dynamics:scalars_old", {{"scalars", 1, 0, 3}}},
        {"dynamics:w", {{"w", 2, 0, 3}}},
        {"dynamics:scale", {{"scale_arr", 1, 0, 3}}},
        // not a reference group: "dynamics:scale" of every scalar in one message (the batched monotonic transport, where the
        // scalar loop of TI:4220 runs inside each kernel; lev 0 = all Dev::mb_planes pairs of the work array)
        {"dynamics:scale_all", {{"scale_arr", 0, 0, 3}}},
        {"initialization:u", {{"u", 1, 1, 7}}},
        {"initialization:pv_edge,ru,rw", {{"pv_edge", 1, 1, 7}, {"ru", 1, 1, 7}, {"rw", 1, 0, 3}}},
    };
    return g;
}

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
    void* lib = nullptr;
};
static NcclApi* nccl_api() {
    static NcclApi api; static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("/usr/lib/x86_64-linux-gnu/libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return nullptr;
#define SYM(n) *(void**)(&api.n) = dlsym(lib, "nccl" #n); if (!api.n) return nullptr;
    SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(GroupStart) SYM(GroupEnd) SYM(Send) SYM(Recv) SYM(GetErrorString)
#undef SYM
    api.lib = lib;
    return &api;
}

static void halo_free_plans(HaloState& hs) {
    for (int p = 0; p < 2; p++) {
        for (auto& kv : hs.plans[p]) {
            HaloGroupPlan& g = kv.second;
            if (g.d_sendbuf) cudaFree(g.d_sendbuf); if (g.d_recvbuf) cudaFree(g.d_recvbuf);
            if (g.d_pack) cudaFree(g.d_pack); if (g.d_unpack) cudaFree(g.d_unpack);
            if (g.d_idx_send) cudaFree(g.d_idx_send); if (g.d_idx_recv) cudaFree(g.d_idx_recv);
        }
        hs.plans[p].clear();
    }
}
static void halo_destroy(HaloState& hs) {
    halo_free_plans(hs);
    for (size_t r = 0; r < hs.peer_mbox.size(); r++) {
        if (hs.peer_mbox[r] && (int)r != hs.rank) cudaIpcCloseMemHandle(hs.peer_mbox[r]);
        if (hs.peer_flags[r] && (int)r != hs.rank) cudaIpcCloseMemHandle(hs.peer_flags[r]);
    }
    if (hs.mbox) cudaFree(hs.mbox);
    if (hs.flags) cudaFree(hs.flags);
    if (hs.done) cudaFree(hs.done);
    if (hs.comm) { NcclApi* a = nccl_api(); if (a) a->CommDestroy((ncclComm_t)hs.comm); hs.comm = nullptr; }
}

extern "C" int mpasb_get_nccl_unique_id(void* out128) {
    NcclApi* a = nccl_api(); if (!a) return 1;
    ncclUniqueId id; if (a->GetUniqueId(&id) != ncclSuccess) return 2;
    memcpy(out128, &id, sizeof(id)); return 0;
}
extern "C" int mpasb_comm_init(mpasb_handle h, int rank, int world_size, const void* id128) {
    cudaSetDevice(h->device);
    NcclApi* a = nccl_api(); if (!a) { h->err = "libnccl.so.2 not found"; return 1; }
    ncclUniqueId id; memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    ncclResult_t r = a->CommInitRank(&comm, world_size, id, rank);
    if (r != ncclSuccess) { h->err = std::string("ncclCommInitRank: ") + a->GetErrorString(r); return 2; }
    h->halo.comm = comm; h->halo.rank = rank; h->halo.world = world_size;
    return 0;
}

extern "C" int mpasb_set_halo_lists(mpasb_handle h, int kind, int n_neighbors, const int* neighbor_rank, int n_layers,
                                    const int* n_send, const int* send_src, const int* n_recv, const int* recv_dst) {
    cudaSetDevice(h->device);
    if (kind < 0 || kind > 2) { h->err = "bad halo kind"; return 1; }
    HaloKind& K = h->halo.kind[kind];
    halo_free_plans(h->halo);
    K.nbr.assign(neighbor_rank, neighbor_rank + n_neighbors);
    K.n_layers = n_layers;
    K.n_send.assign(n_send, n_send + n_neighbors * n_layers);
    K.n_recv.assign(n_recv, n_recv + n_neighbors * n_layers);
    K.send_off.assign(n_neighbors * n_layers + 1, 0); K.recv_off.assign(n_neighbors * n_layers + 1, 0);
    for (int i = 0; i < n_neighbors * n_layers; i++) { K.send_off[i + 1] = K.send_off[i] + K.n_send[i]; K.recv_off[i + 1] = K.recv_off[i] + K.n_recv[i]; }
    const int ts = K.send_off.back(), tr = K.recv_off.back();
    K.h_send.resize(ts); K.h_recv.resize(tr);
    for (int i = 0; i < ts; i++) K.h_send[i] = send_src[i] - 1;     // ABI is 1-based
    for (int i = 0; i < tr; i++) K.h_recv[i] = recv_dst[i] - 1;
    h->halo.active = true;
    h->ac_lists_ok = false;            // boundary / interior column lists of the cell solve follow the send lists
    h->dd_lists_ok = false;            // ... and so do the edges the damping kernel must finish before the ru_p exchange
    return 0;
}

// Build (once per group and time-level parity) the pack/unpack segment tables and buffers.
static int halo_build_plan(H* h, const GroupDef& gd, HaloGroupPlan& P) {
    HaloState& hs = h->halo;
    const int LDK = h->D.LDK, nl = h->D.nl;
    std::vector<int> peers;
    for (const GroupField& gf : gd.fields) for (int r : hs.kind[gf.kind].nbr) if (std::find(peers.begin(), peers.end(), r) == peers.end()) peers.push_back(r);
    std::sort(peers.begin(), peers.end());
    P.peers = peers;
    std::vector<HaloSeg> pack, unpack;
    std::vector<int> idx_s, idx_r;
    size_t soff = 0, roff = 0;
    for (size_t pi = 0; pi < peers.size(); pi++) {
        const int peer = peers[pi];
        P.send_off.push_back(soff); P.recv_off.push_back(roff);
        for (const GroupField& gf : gd.fields) {
            FieldRec* f = find_field(h, gf.name); if (!f) return 1;
            const HaloKind& K = hs.kind[gf.kind];
            int ni = -1;
            for (size_t n = 0; n < K.nbr.size(); n++) if (K.nbr[n] == peer) ni = (int)n;
            if (ni < 0) continue;
            int nplanes = 1, width = nl; size_t plane = 0; int stride = LDK;
            if (f->inner == IN_NL1) width = nl + 1;
            else if (f->inner == IN_S_NL) { nplanes = h->dims.num_scalars; plane = outer_of(h, f->loc) * LDK; }
            else if (f->inner == IN_NL_TWO) { nplanes = 2; plane = outer_of(h, f->loc) * LDK; }
            else if (f->inner != IN_NL) { h->err = std::string("halo: unsupported field shape ") + gf.name; return 1; }
            real* base = gf.lev >= 1 ? (real*)f->d[gf.lev - 1] : nullptr;
            if (gf.lev == 0) {                             // every per-scalar pair of scale_arr
                if (strcmp(gf.name, "scale_arr")) { h->err = "halo: level 0 is scale_arr only"; return 1; }
                base = h->D.mb_scale; nplanes = 2 * h->D.mb_planes;
            }
            for (int p = 0; p < nplanes; p++)
                for (int l = 0; l < K.n_layers; l++) {
                    if (!(gf.layers & (1 << l))) continue;
                    const int li = ni * K.n_layers + l;
                    if (K.n_send[li]) {
                        HaloSeg s; s.field = base + p * plane; s.idx_off = (int)idx_s.size(); s.count = K.n_send[li]; s.width = width; s.peer = (int)pi; s.buf_off = soff; s.stride = stride;
                        idx_s.insert(idx_s.end(), K.h_send.begin() + K.send_off[li], K.h_send.begin() + K.send_off[li + 1]);
                        pack.push_back(s); soff += (size_t)s.count * width;
                        P.max_seg = std::max(P.max_seg, (size_t)s.count * width);
                    }
                    if (K.n_recv[li]) {
                        HaloSeg s; s.field = base + p * plane; s.idx_off = (int)idx_r.size(); s.count = K.n_recv[li]; s.width = width; s.peer = (int)pi; s.buf_off = roff; s.stride = stride;
                        idx_r.insert(idx_r.end(), K.h_recv.begin() + K.recv_off[li], K.h_recv.begin() + K.recv_off[li + 1]);
                        unpack.push_back(s); roff += (size_t)s.count * width;
                        P.max_seg = std::max(P.max_seg, (size_t)s.count * width);
                    }
                }
        }
        P.send_cnt.push_back(soff - P.send_off.back()); P.recv_cnt.push_back(roff - P.recv_off.back());
    }
    if (!idx_s.empty()) { CUDA_OK(cudaMalloc(&P.d_idx_send, idx_s.size() * sizeof(int))); CUDA_OK(cudaMemcpy(P.d_idx_send, idx_s.data(), idx_s.size() * sizeof(int), cudaMemcpyHostToDevice)); }
    if (!idx_r.empty()) { CUDA_OK(cudaMalloc(&P.d_idx_recv, idx_r.size() * sizeof(int))); CUDA_OK(cudaMemcpy(P.d_idx_recv, idx_r.data(), idx_r.size() * sizeof(int), cudaMemcpyHostToDevice)); }
    P.n_pack = (int)pack.size(); P.n_unpack = (int)unpack.size();
    if (soff) CUDA_OK(cudaMalloc(&P.d_sendbuf, soff * sizeof(real)));
    if (roff) CUDA_OK(cudaMalloc(&P.d_recvbuf, roff * sizeof(real)));
    if (P.n_pack) { CUDA_OK(cudaMalloc(&P.d_pack, P.n_pack * sizeof(HaloSeg))); CUDA_OK(cudaMemcpy(P.d_pack, pack.data(), P.n_pack * sizeof(HaloSeg), cudaMemcpyHostToDevice)); }
    if (P.n_unpack) { CUDA_OK(cudaMalloc(&P.d_unpack, P.n_unpack * sizeof(HaloSeg))); CUDA_OK(cudaMemcpy(P.d_unpack, unpack.data(), P.n_unpack * sizeof(HaloSeg), cudaMemcpyHostToDevice)); }
    return 0;
}

static int halo_exchange(H* h, const char* group, cudaStream_t stream) {
    HaloState& hs = h->halo;
    const GroupDef* gd = nullptr;
    for (const GroupDef& g : group_table()) if (!strcmp(g.name, group)) gd = &g;
    if (!gd) { h->err = std::string("unknown halo group ") + group; return 1; }
    auto& plans = hs.plans[hs.parity];
    auto it = plans.find(group);
    if (it == plans.end()) {
        HaloGroupPlan P;
        if (halo_build_plan(h, *gd, P)) return 1;
        it = plans.emplace(group, P).first;
    }
    HaloGroupPlan& P = it->second;
    if (P.peers.empty()) return 0;
    const unsigned gx = (unsigned)std::max<size_t>(1, std::min<size_t>((P.max_seg + 255) / 256, 64));
    if (hs.p2p) {
        // one put kernel and one get kernel; no library call, no intermediate buffers
        if ((int)P.peers.size() > P2P_MAXP) { h->err = "p2p exchange: more than 8 neighbours"; return 1; }
        P2PPeers pp; memset(&pp, 0, sizeof(pp));
        pp.n = (int)P.peers.size();
        for (int p = 0; p < pp.n; p++) {
            const int q = P.peers[p];
            if (P.send_cnt[p] > hs.slot_elems || P.recv_cnt[p] > hs.slot_elems) { h->err = "p2p exchange: message larger than the mailbox slot"; return 1; }
            pp.seq_send[p] = P.send_cnt[p] ? ++hs.seq_send[q] : 0;
            pp.seq_recv[p] = P.recv_cnt[p] ? ++hs.seq_recv[q] : 0;
            pp.remote[p] = hs.peer_mbox[q] + ((size_t)hs.rank * 2 + (pp.seq_send[p] & 1)) * hs.slot_elems;
            pp.local[p] = hs.mbox + ((size_t)q * 2 + (pp.seq_recv[p] & 1)) * hs.slot_elems;
            pp.remote_arrived[p] = hs.peer_flags[q] + hs.rank;
            pp.remote_consumed[p] = hs.peer_flags[q] + hs.world + hs.rank;
            pp.local_arrived[p] = hs.flags + q;
            pp.local_consumed[p] = hs.flags + hs.world + q;
            pp.send_off[p] = P.send_off[p]; pp.recv_off[p] = P.recv_off[p];
        }
        if (P.n_pack) { k_halo_put<<<dim3(gx, std::min(P.n_pack, 256)), 256, 0, stream>>>(P.d_pack, P.d_idx_send, P.n_pack, pp, hs.done); h->launches++; }
        if (P.n_unpack) { k_halo_get<<<dim3(gx, std::min(P.n_unpack, 256)), 256, 0, stream>>>(P.d_unpack, P.d_idx_recv, P.n_unpack, pp, hs.done + 1); h->launches++; }
        return 0;
    }
    if (!hs.comm) { h->err = "halo lists are set but mpasb_comm_init was not called"; return 1; }
    NcclApi* a = nccl_api();
    if (P.n_pack) { k_halo_pack<<<dim3(gx, std::min(P.n_pack, 256)), 256, 0, stream>>>(P.d_pack, P.d_idx_send, P.d_sendbuf, P.n_pack); h->launches++; }
    a->GroupStart();
    for (size_t p = 0; p < P.peers.size(); p++) {
        if (P.recv_cnt[p]) a->Recv(P.d_recvbuf + P.recv_off[p], P.recv_cnt[p], (sizeof(real) == 8 ? ncclFloat64 : ncclFloat32), P.peers[p], (ncclComm_t)hs.comm, stream);
        if (P.send_cnt[p]) a->Send(P.d_sendbuf + P.send_off[p], P.send_cnt[p], (sizeof(real) == 8 ? ncclFloat64 : ncclFloat32), P.peers[p], (ncclComm_t)hs.comm, stream);
    }
    ncclResult_t r = a->GroupEnd();
    if (r != ncclSuccess) { h->err = std::string("nccl: ") + a->GetErrorString(r); return 1; }
    if (P.n_unpack) { k_halo_unpack<<<dim3(gx, std::min(P.n_unpack, 256)), 256, 0, stream>>>(P.d_unpack, P.d_idx_recv, P.d_recvbuf, P.n_unpack); h->launches++; }
    return 0;
}

// ---- CUDA-IPC peer-to-peer set-up: (1) every rank reports its largest message, (2) allocates mailbox + flags for the
// largest one over all ranks and exports two IPC handles, (3) maps the handles of all ranks.
extern "C" long mpasb_p2p_max_message(mpasb_handle h) {
    cudaSetDevice(h->device);
    HaloState& hs = h->halo;
    size_t mx = 0;
    for (const GroupDef& gd : group_table()) {
        auto& plans = hs.plans[hs.parity];
        auto it = plans.find(gd.name);
        if (it == plans.end()) {
            HaloGroupPlan P;
            if (halo_build_plan(h, gd, P)) return -1;
            it = plans.emplace(gd.name, P).first;
        }
        for (size_t p = 0; p < it->second.peers.size(); p++) mx = std::max(mx, std::max(it->second.send_cnt[p], it->second.recv_cnt[p]));
    }
    return (long)mx;
}
extern "C" int mpasb_p2p_prepare(mpasb_handle h, long slot_elems, void* out_handles128) {
    cudaSetDevice(h->device);
    HaloState& hs = h->halo;
    if (hs.world < 2 || hs.world > P2P_MAXP + 1 || slot_elems <= 0) { h->err = "mpasb_p2p_prepare: needs 2..9 ranks (mpasb_comm_init first) and a positive slot"; return 1; }
    hs.slot_elems = ((size_t)slot_elems + 1) / 2 * 2;                  // keeps every slot 16-byte aligned
    const size_t mbytes = (size_t)hs.world * 2 * hs.slot_elems * sizeof(real), fbytes = (size_t)2 * hs.world * sizeof(unsigned long long);
    CUDA_OK(cudaMalloc(&hs.mbox, mbytes)); CUDA_OK(cudaMalloc(&hs.flags, fbytes)); CUDA_OK(cudaMalloc(&hs.done, 2 * sizeof(unsigned)));
    CUDA_OK(cudaMemset(hs.mbox, 0, mbytes)); CUDA_OK(cudaMemset(hs.flags, 0, fbytes)); CUDA_OK(cudaMemset(hs.done, 0, 2 * sizeof(unsigned)));
    CUDA_OK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t hm, hf;
    CUDA_OK(cudaIpcGetMemHandle(&hm, hs.mbox)); CUDA_OK(cudaIpcGetMemHandle(&hf, hs.flags));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "two handles fill 128 bytes");
    memcpy(out_handles128, &hm, 64); memcpy((char*)out_handles128 + 64, &hf, 64);
    return 0;
}
extern "C" int mpasb_p2p_open(mpasb_handle h, const void* all_handles /* world x 128 bytes, rank order */) {
    cudaSetDevice(h->device);
    HaloState& hs = h->halo;
    if (!hs.mbox) { h->err = "mpasb_p2p_open before mpasb_p2p_prepare"; return 1; }
    hs.peer_mbox.assign(hs.world, nullptr); hs.peer_flags.assign(hs.world, nullptr);
    hs.seq_send.assign(hs.world, 0); hs.seq_recv.assign(hs.world, 0);
    for (int r = 0; r < hs.world; r++) {
        if (r == hs.rank) { hs.peer_mbox[r] = hs.mbox; hs.peer_flags[r] = hs.flags; continue; }
        cudaIpcMemHandle_t hm, hf;
        memcpy(&hm, (const char*)all_handles + (size_t)r * 128, 64); memcpy(&hf, (const char*)all_handles + (size_t)r * 128 + 64, 64);
        void* pm = nullptr; void* pf = nullptr;
        CUDA_OK(cudaIpcOpenMemHandle(&pm, hm, cudaIpcMemLazyEnablePeerAccess));
        CUDA_OK(cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess));
        hs.peer_mbox[r] = (real*)pm; hs.peer_flags[r] = (unsigned long long*)pf;
    }
    return 0;
}
// collective decision of the host: switch the peer-to-peer path on only when every rank mapped every peer
extern "C" int mpasb_p2p_enable(mpasb_handle h, int on) {
    HaloState& hs = h->halo;
    if (on && (hs.peer_mbox.empty() || !hs.mbox)) { h->err = "mpasb_p2p_enable before mpasb_p2p_open"; return 1; }
    hs.p2p = on != 0;
    return 0;
}

extern "C" int mpasb_exchange_halo_group_async(mpasb_handle h, const char* group_name) {
    cudaSetDevice(h->device);
    if (!h->halo.active) return 0;
    return exchange(h, group_name) ? 1 : 0;
}
extern "C" int mpasb_exchange_halo_group(mpasb_handle h, const char* group_name) {
    cudaSetDevice(h->device);
    if (!h->halo.active) return 0;
    if (exchange(h, group_name)) return 1;
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}
