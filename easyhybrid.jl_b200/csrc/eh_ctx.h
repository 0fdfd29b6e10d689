// eh_ctx.h -- the library context (opaque `eh_ctx` of include/easyhybrid_cuda.h) and what the translation units of the
// host runtime share: eh_lib.cu (launch sequences, host-batch ring, data-parallel exchange, C ABI) and eh_plan.cu (model
// descriptor -> kernel variant + layout tables).  Internal; nothing here is part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/easyhybrid_cuda.h"
#include "eh_variants.h"
#include "eh_jit.h"
#include "eh_epoch_kernel.cuh"
#include "eh_eval_kernel.cuh"
#include "eh_wide.h"

using namespace eh;

namespace eh {
namespace rt {


struct Split {
    float* rec = nullptr;
    int64_t N = 0;
    float shift_y[MAXT] = {0, 0, 0, 0};
    float shift_x[MAXP] = {0};
    bool has_nan = false;
};

struct HostStage {  // device staging for eh_step_host*: raw arrays + packed records
    float* d_X = nullptr;
    float* d_planes = nullptr;
    float* d_rec = nullptr;
    int* d_cnt = nullptr;  // [MAXT] valid-target counts written by the packer
    float* d_bscal = nullptr;
    float* d_loss = nullptr;
    cudaEvent_t ready = nullptr;  // the batch's H2D copies have landed (recorded on the copy stream)
    cudaEvent_t freed = nullptr;  // the step that consumed this slot has retired (recorded on the compute stream)
    bool used = false;
    int64_t cap = 0;
};
constexpr int EH_HOST_SLOTS = 4;  // batches in flight between the copy engine and the step kernels
#ifndef EH_NPACK_STREAMS
#define EH_NPACK_STREAMS 4
#endif
constexpr int EH_NPACK = EH_NPACK_STREAMS;  // host-batch packers / copies in flight (one stream each)

// eh_step_host_async, grouped form: page-locked batches are packed (zero copy) into a ring of staging slots; every
// EH_RING_GROUP batches ONE persistent launch runs that many optimiser steps over the group's slots, while the packers
// of the next group keep the PCIe link busy.  Three groups: one training, one being packed, one draining.
constexpr int EH_RING_GROUP = 16, EH_RING_NGRP = 3;
constexpr int64_t EH_RING_MAX_BATCH = 1 << 18;   // batches beyond this take the one-launch-pair-per-batch form
// consumer mode of eh_step_host_async: ONE persistent launch per burst trains on the ring slots as the packers publish them
struct HostStream {
    bool active = false;       // a consumer kernel is running
    bool off = false;          // EH_HOST_NO_STREAM=1: grouped launches instead
    int64_t B = 0;             // batch size of the running burst
    unsigned global = 0;       // batches ever handed to this mode (slot = global % slots); tags and `done` count in it
    int count = 0;             // steps of the running burst so far
    int* h_total = nullptr;    // page-locked: number of steps of the burst, written when the burst is closed
    unsigned* d_ready = nullptr;   // [slots]
    unsigned* d_done = nullptr;    // [1]
};
constexpr int EH_STREAM_MAX_STEPS = 4096;   // steps per consumer launch (statistics buffer); longer bursts are cut there

struct HostRing {
    float* d_rec = nullptr;    // [NGRP][GROUP * cap][R4]; slot k of a group starts at record k * B (B = the group's batch size)
    float* d_bscal = nullptr;  // [NGRP * GROUP][BS_STRIDE]
    int* d_cnt = nullptr;      // [NGRP * GROUP][MAXT + 1]
    cudaEvent_t packed[EH_NPACK] = {};  // last packer of the open group on each pack stream
    cudaEvent_t freed[EH_RING_NGRP] = {nullptr, nullptr, nullptr};
    bool used[EH_RING_NGRP] = {false, false, false};
    int64_t cap = 0;           // samples per slot
    int g = 0, k = 0;          // open group, batches packed into it so far
    unsigned rr = 0;           // round robin over the pack streams
    int launches = 0;          // groups launched in the running burst
    int limit = 1;             // size at which the open group is launched: 1, 2, 4, 8, 16, 16, ... within a burst (the first
                               // steps start while later batches are still crossing PCIe); back to 1 at eh_sync
    int64_t B = 0;             // batch size of the open group
    float* loss0 = nullptr;    // page-locked loss cell of the group's first step (the others follow contiguously)
    bool off = false;          // EH_HOST_NO_GROUPS=1
};

}  // namespace rt
}  // namespace eh
using namespace eh::rt;

struct eh_ctx {
    std::string err;
    int device = 0;
    int nsm = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaStream_t pack_stream[EH_NPACK] = {};   // packers of consecutive host batches rotate over these ([0] == copy_stream)
    const Variant* var = nullptr;   // engine chosen at eh_create (FFMA2 one sample per lane, or tensor pipe)
    const Variant* var2 = nullptr;  // FFMA2 two samples per lane: same layouts, used for large batches
    const Variant* var_tc = nullptr;  // tensor engine (tcgen05 tiles of 128 samples): same layouts, persistent kernel, large batches
    // traced process model compiled at run time (eh_jit.cu): `var` then points at `jit_var`, a copy of the generic
    // variant of the shape (same layouts) without launchers -- the kernels are the handles in `jit`
    bool jit_on = false;
    Variant jit_var{};
    eh::JitKernels jit;
    std::string jit_cubin, jit_names[3];
    // wide-hidden-layer path (bf16 tcgen05 GEMMs, eh_wide.cu): `var` then points at `wide_var`, a descriptor
    // without kernels that only carries the record / slot geometry the shared host code reads
    eh::wide::WideNet* wide = nullptr;
    Variant wide_var{};
    eh::wide::WideModel wide_model{};
    std::vector<int> h_wmap;        // [nflat][4] {kind, layer, image row, image column} of the embedded chain
    // model
    int n_pred_raw = 0, n_forc_raw = 0, n_targ = 0;
    int nflat = 0, ntheta = 0, nglob = 0;
    int real_in = 0;
    int n_chains = 1, chain_in0[4] = {0, 0, 0, 0}, chain_nin[4] = {0, 0, 0, 0};   // chain k owns inputs [in0, in0 + nin) of the embedded chain
    std::vector<int> h_wsrc, h_pmap;
    std::vector<float> h_pspan;
    PSlot slots[MAXPS];
    float pmc[4] = {0, 0, 0, 0};
    int loss_kind[MAXT] = {0, 0, 0, 0};   // as the kernels see it (LOSS_AFFINE for the prediction-statistics losses)
    int loss_kind_abi[MAXT] = {0, 0, 0, 0};
    bool l2_on = false;       // native weight_l2 extra loss: one launch pair per step (k_update adds the term)
    float l2_aggw = 1.f, l2_loss_coef = 0.f;
    std::vector<float> h_l2coef;
    float* d_l2coef = nullptr;
    bool stat_loss = false;   // some target's seeds need statistics of the predictions: forward pre-pass per step, no persistent kernel
    double* d_statpart = nullptr;
    int statpart_cap = 0;
    int agg_mean = 0;
    int opt_kind = 0, adamw_coupled = 1;
    float eta = 0.01f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f, lambda = 0.f;
    int use_bn = 0;
    unsigned flags = 0;
    int src_kind[24], src_idx[24], ncols = 0;
    int nparam_desc = 0;
    std::vector<int> slot_of_param;  // desc parameter index -> canonical slot or -1
    // device state
    int *d_wsrc = nullptr, *d_pmap = nullptr;
    float* d_pspan = nullptr;
    float *d_theta = nullptr, *d_m = nullptr, *d_v = nullptr, *d_grad = nullptr;
    OptState* d_ost = nullptr;
    // persistent epoch kernel
    std::vector<int> h_cells, h_slot_of_flat;
    int *d_cells = nullptr, *d_slot_of_flat = nullptr, *d_losskind = nullptr;
    float *d_pbuf = nullptr, *d_stats = nullptr;
    int epoch_tiles = 0, epoch_grid = 0, epoch_warps = 0;  // last persistent launch geometry
    const Variant* geo_var = nullptr;                      // cached launch geometry of the persistent kernel
    int64_t geo_B = 0;
    int geo_mode = -1, geo_G = 0, geo_w = 0, geo_tile = 0, geo_pg = 0;
    // the persistent launch as a three-node CUDA graph (event record, kernel, event record): the whole launch reaches the
    // GPU at once, so the timed interval holds no host submission latency (a cooperative launch costs ~30 us of host time),
    // and a graph launch is cheaper on the host than cudaLaunchCooperativeKernel.  Rebuilt when the geometry changes.
    cudaGraph_t pg_graph = nullptr;
    cudaGraphExec_t pg_exec = nullptr;
    cudaGraphNode_t pg_knode = nullptr;
    const void* pg_func = nullptr;
    int pg_G = 0, pg_threads = 0;
    size_t pg_smem = 0;
    bool pg_off = false;
    size_t stats_cap = 0;
    bool persist_ok = false;
    int pm_id = 0;
    float *d_partial = nullptr, *d_gvec = nullptr;
    float* d_bscal = nullptr;
    size_t bscal_cap = 0;
    float* d_bn_batch = nullptr;
    size_t bn_batch_cap = 0;
    int* d_idx = nullptr;
    long long* d_idx64 = nullptr;
    size_t idx_cap = 0;
    int* d_err = nullptr;
    float* d_loss = nullptr;
    float* h_loss = nullptr;  // pinned
    size_t loss_cap = 0;
    double* d_evalpart = nullptr;
    float* d_bn_test = nullptr;  // BS_STRIDE row with running stats for test mode
    Split split[2];
    int64_t perm_n = 0;
    // epoch staging: the train records in the order of the resident index stream (d_idx), for the persistent kernel
    float* d_stage = nullptr;
    size_t stage_cap = 0;            // records
    unsigned idx_gen = 1, stage_gen = 0;   // d_stage mirrors d_idx when the generations agree
    bool stage_on = true;            // EH_NO_STAGE=1: the persistent kernel gathers through the index stream instead
    int64_t perm_B = 0;  // batch size the bscal rows were prepared for (0 = none)
    std::vector<float> bn_mean, bn_var;
    // host-step pipeline
    HostStage hs[EH_HOST_SLOTS];
    int hs_next = 0;
    HostRing ring;
    bool small_prog = false;     // register-tile path with an interpreted process model (PmProgram variants)
    int scale_rt = 0;            // scale_nn_outputs as the generic variants take it
    unsigned pass_mask[3] = {0, 0, 0};   // generic variants: pass-through units per hidden layer (chains of unequal depth)
    PmProgData h_prog;           // the program, host copy
    PmProgData* d_prog = nullptr;
    bool host_zero_copy = true;  // EH_HOST_NO_ZEROCOPY=1: always stage host batches through the copy engine
    HostStream hstream;
    std::vector<std::pair<float*, float*>> pending_loss;  // (pinned src, user dst)
    struct PendingBn { const float* loss; const float* mom; int64_t B; };
    std::vector<PendingBn> pending_bn;                      // host batches whose BatchNorm batch moments still have to be folded in
    float* h_async_bn = nullptr;                            // pinned ring [async_cap][2 * MAXP]: (mean, biased var) per input
    float* h_bn0 = nullptr;                                 // pinned [2 * MAXP] + loss cell for the synchronous eh_step_host
    float* h_async_loss = nullptr;                          // pinned ring
    size_t async_cap = 0, async_used = 0;
    // timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    std::vector<cudaEvent_t> seg_ev;  // index-segment arrival events of the pipelined epoch
    float* d_snap = nullptr;          // trainable-state snapshot of the pipelined epoch
    float last_ms = 0.f, last_step_ms = 0.f;
    int64_t last_launches = 0;
    int profiling = 0;
    std::vector<cudaEvent_t> prof_ev;
    // epoch graph
    cudaGraphExec_t gexec = nullptr;
    int64_t g_n = 0, g_B = 0;
    int g_pdl = 0;
    bool g_has_pdl = false;
    const int* g_idx = nullptr;
    const float* g_bscal = nullptr;
    const float* g_loss = nullptr;
    // data parallel: inbox block = [2 parities][8 ranks][npartp] {value, tag} slots, IPC-shared
    int rank = 0, world = 1;
    void* dp_block = nullptr;
    void* dp_peer[EH_MAX_WORLD] = {nullptr};
    unsigned dp_steps = 0;  // steps exchanged so far (absolute flag tags)
    unsigned epoch_tag = 0; // steps run by the persistent kernel so far (tags of the in-GPU exchange; never reset)
    unsigned* d_dperr = nullptr;
};

namespace eh {
namespace rt {
// error text into the ctx (or, without one, into the per-thread slot eh_last_error(NULL) reads); returns `s`
eh_status fail(eh_ctx* c, eh_status s, const char* fmt, ...);
void set_create_error(const std::string& s);
// descriptor -> plan (eh_plan.cu): kernel variant, flat layout, gather / scatter tables, slots, loss and optimiser settings
eh_status build_plan(eh_ctx* c, const eh_model_desc* d);
}  // namespace rt
}  // namespace eh

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail(c, EH_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
