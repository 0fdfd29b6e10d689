// eh_layout.h -- compile-time shape math shared by the fused step kernel (device)
// and the host planner (eh_plan.cu).  No includes, NVRTC-clean.
//
// A "shape" is one Dense chain  P -> H -> ... -> H -> NOUT  with NH hidden
// layers, every hidden width padded up to H (multiple of 4) with zero weights.
// Reference: prepare_hidden_chain, src/models/NNModels.jl:220-231.
#pragma once

#ifndef EH_HD
#ifdef __CUDACC__
#define EH_HD __host__ __device__
#else
#define EH_HD
#endif
#endif

namespace eh {

enum : int { ACT_IDENTITY = 0, ACT_TANH = 1, ACT_SIGMOID = 2, ACT_RELU = 3, ACT_SWISH = 4 };
enum : int { ROLE_NEURAL = 0, ROLE_GLOBAL = 1, ROLE_FIXED = 2 };
enum : int { LOSS_MSE = 0, LOSS_RMSE = 1, LOSS_MAE = 2, LOSS_NSELOSS = 3,
              // ABI kinds whose seeds depend on statistics of the batch's predictions (a forward pre-pass computes them)
              LOSS_PEARSONLOSS = 4, LOSS_KGELOSS = 5, LOSS_PBKGELOSS = 6,
              // what the step kernels see for those targets (and for rmse over several targets): seeds that are affine in
              // (1, yhat, y) with three per-batch coefficients, dL/dyhat_i = sa + sb yhat_i + sc y_i
              LOSS_AFFINE = 7 };
enum : int { PM_RBQ10 = 0, PM_EXPO = 1, PM_LINEAR = 2, PM_LINEAR2 = 3, PM_EXPO2 = 4, PM_PROGRAM = 100 };

constexpr int MAXPS = 8;   // process-parameter slots
constexpr int MAXT = 4;    // targets
constexpr int MAXF = 8;    // forcings
constexpr int CHUNK = 32;  // samples per warp pass (one per lane)
constexpr int ROWSTRIDE = CHUNK + 4;  // floats per staging row (bank skew)

EH_HD constexpr int rup4(int x) { return (x + 3) & ~3; }

// runtime/compile-time description of one chain shape.  LR = 1: the weight gradient of the
// (narrow) linear output layer is accumulated in registers by each lane instead of through
// staging tiles (saves the a_NH / delta_out staging rows and their tiles).
struct ShapeDims {
    int P, NH, H, NOUT, LR;
    EH_HD constexpr int nlayers() const { return NH + 1; }
    EH_HD constexpr int nlt() const { return LR ? NH : NH + 1; }  // layers whose dW goes through tiles
    // fan-in / padded fan-out of dense layer l (1-based)
    EH_HD constexpr int din(int l) const { return l == 1 ? P : H; }
    EH_HD constexpr int dout4(int l) const { return l == NH + 1 ? rup4(NOUT) : H; }
    // rows of the augmented input [a; 1; 0..] of layer l, padded to 4
    EH_HD constexpr int ka(int l) const { return rup4(din(l) + 1); }
    EH_HD constexpr int nj(int l) const { return dout4(l) / 4; }
    EH_HD constexpr int nk(int l) const { return ka(l) / 4; }
    // staging groups (4 rows each; group g starts at row 5g): for l = 1..nlt: A_{l-1} then D_l
    EH_HD constexpr int gA(int l) const
    {
        int g = 0;
        for (int i = 1; i < l; i++) g += nk(i) + nj(i);
        return g;
    }
    EH_HD constexpr int gD(int l) const { return gA(l) + nk(l); }
    EH_HD constexpr int ngroups() const { return gA(nlt() + 1); }
    EH_HD constexpr int nrows() const { return 5 * ngroups(); }
    // dW blocks (4x4): layer-major, then j-block, then k-block
    EH_HD constexpr int blk0(int l) const
    {
        int b = 0;
        for (int i = 1; i < l; i++) b += nj(i) * nk(i);
        return b;
    }
    EH_HD constexpr int nblocks() const { return blk0(nlt() + 1); }
    // shared-memory weight image (floats)
    EH_HD constexpr int off_w1f() const { return 0; }                     // [P][H]
    EH_HD constexpr int off_b1() const { return P * H; }                  // [H]
    EH_HD constexpr int off_wf(int l) const { return P * H + H + (l - 2) * (2 * H * H + H); }  // l = 2..NH, [H][H] k-major
    EH_HD constexpr int off_b(int l) const { return off_wf(l) + H * H; }  // [H]
    EH_HD constexpr int off_wb(int l) const { return off_b(l) + H; }      // [H][H] j-major
    EH_HD constexpr int off_wo() const { return P * H + H + (NH - 1) * (2 * H * H + H); }  // [NOUT][H]
    EH_HD constexpr int off_bo() const { return off_wo() + NOUT * H; }    // [4]
    EH_HD constexpr int nweights() const { return off_bo() + 4; }
    // per-CTA partial vector: nblocks*16 dW entries, statistics, then (LR) the output layer [NOUT][H+1]
    EH_HD constexpr int npart_dw() const { return nblocks() * 16; }
    EH_HD constexpr int nlast() const { return LR ? rup4(NOUT * (H + 1)) : 0; }
    EH_HD constexpr int off_last() const { return npart_dw() + MAXT + MAXPS; }
    EH_HD constexpr int npart() const { return off_last() + nlast(); }
};

// parameter block in device memory: [nflat theta/phi | MAXPS uniform slot values | MAXPS*4 derived
// process-model scalars]; the tail is refreshed by whoever updates phi
constexpr int PMS_PER_SLOT_L = 4;
constexpr int PARAM_TAIL = MAXPS + MAXPS * PMS_PER_SLOT_L;

// statistics appended after the dW blocks in a partial vector:
//   [T] sum of r^2 (or |r| for mae targets), [MAXPS] sum of g*dy/dslot for GLOBAL slots
constexpr int NSTAT = MAXT + MAXPS;
// longest partial vector k_update reduces in its shared memory (checked per variant at compile time and by the planner)
constexpr int UPD_MAX_NPART = 4096;

}  // namespace eh
