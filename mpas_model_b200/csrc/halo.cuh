// halo.cuh -- device-resident halo exchange state.
// Restates the data the reference keeps per field and per group
// (src/framework/mpas_halo_types.inc:12-86): per (neighbour, halo layer) send and
// receive index lists, aggregated per exchange group into one message per neighbour.
#pragma once
#include <map>
#include <string>
#include <vector>
#include "kernels_diag.cuh"

struct HaloKind {                       // one of cell / edge / vertex
    std::vector<int> nbr;               // neighbour ranks
    int n_layers = 0;
    std::vector<int> n_send, n_recv;    // [nbr][layer]
    std::vector<int> send_off, recv_off;   // offsets into d_send / d_recv (elements)
    std::vector<int> h_send, h_recv;    // 0-based local indices: owned elements to send / halo elements to fill
};

struct HaloGroupPlan {
    std::vector<int> peers;             // union of neighbour ranks
    std::vector<size_t> send_off, send_cnt, recv_off, recv_cnt;   // per peer, in reals
    real* d_sendbuf = nullptr; real* d_recvbuf = nullptr;
    HaloSeg* d_pack = nullptr; HaloSeg* d_unpack = nullptr;
    int* d_idx_send = nullptr; int* d_idx_recv = nullptr;   // index lists of this group, all kinds concatenated
    int n_pack = 0, n_unpack = 0;
    size_t max_seg = 0;
};

struct HaloState {
    bool active = false;
    int rank = 0, world = 1;
    HaloKind kind[3];
    void* comm = nullptr;               // ncclComm_t
    void* nccl_lib = nullptr;
    int parity = 0;                     // flips with mpas_pool_shift_time_levels
    // CUDA-IPC peer-to-peer exchange (k_halo_put / k_halo_get): own mailbox and flags, the peers' mappings, message counters
    bool p2p = false;
    size_t slot_elems = 0;
    real* mbox = nullptr; unsigned long long* flags = nullptr; unsigned* done = nullptr;     // done[0]: put, done[1]: get
    std::vector<real*> peer_mbox; std::vector<unsigned long long*> peer_flags;
    std::vector<unsigned long long> seq_send, seq_recv;     // per peer rank
    std::map<std::string, HaloGroupPlan> plans[2];
};

static void halo_destroy(HaloState& hs);
