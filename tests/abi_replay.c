/* abi_replay.c -- the drop-in boundary exercised FROM C: replays, call by call, the sequence the reference-side binding
 * (julia/EasyHybridCUDA.jl) issues for one training run, against include/easyhybrid_cuda.h / libeasyhybrid_cuda.so.
 *
 *   reference call                                                     -> C ABI call
 *   constructHybridModel + TrainConfig (src/config/TrainingConfig.jl)  -> eh_create(eh_model_desc)
 *   prepare_data / split_data, once   (src/data/prepare_data.jl:3-10)   -> eh_upload(TRAIN), eh_upload(VAL)
 *   LuxCore.setup -> ComponentArray(ps)                                 -> eh_set_params
 *   compute_loss + Zygote.gradient    (parity hook)                     -> eh_loss_grad
 *   run_epoch!(loader, ...)           (src/training/epoch.jl:13-33)     -> eh_epoch(perm of the DataLoader, batchsize)
 *   evaluate_epoch                    (src/training/epoch.jl:52-66)     -> eh_eval(TRAIN), eh_eval(VAL)
 *   train_state.parameters / optimizer_state at the end                 -> eh_get_params, eh_get_opt_state
 *   Lux.Training.single_train_step!(::FusedCUDA, ...) per host batch    -> eh_step_host (second ctx, same batches)
 *
 * Inputs are synthetic and written to <out>.bin so that the test (tests/test_abi_replay.py) runs the CPU checker on
 * exactly the same bytes; results go to <out>.txt as "key v0 v1 ..." lines.
 * Build: gcc -O1 -I include tests/abi_replay.c -o abi_replay -L easyhybrid.jl_b200 -l:libeasyhybrid_cuda.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "easyhybrid_cuda.h"

static uint32_t lcg_state = 12345u;
static float urand(void)
{
    lcg_state = lcg_state * 1664525u + 1013904223u;
    return (float)(lcg_state >> 8) * (1.0f / 16777216.0f);
}

#define CHECK(call)                                                                                     \
    do {                                                                                                \
        eh_status st__ = (call);                                                                        \
        if (st__ != EH_OK) {                                                                            \
            fprintf(stderr, "%s -> status %d: %s\n", #call, (int)st__, eh_last_error(ctx));             \
            return 2;                                                                                   \
        }                                                                                               \
    } while (0)

static eh_status create_rbq10(eh_ctx** out)
{
    /* RbQ10 hybrid of the README: NN(sw_pot, dsw_pot) -> rb in [0, 13]; global Q10 in [1, 4]; forcing ta; target reco */
    static const int32_t in_cols[2] = {0, 1}, hidden[2] = {16, 16};
    static eh_chain_desc chain;
    static const int32_t role[2] = {EH_ROLE_NEURAL, EH_ROLE_GLOBAL}, role_index[2] = {0, 0};
    static const float deflt[2] = {3.0f, 2.0f}, lower[2] = {0.0f, 1.0f}, upper[2] = {13.0f, 4.0f};
    static const eh_pm_arg pm_args[3] = {{0, 0}, {0, 1}, {1, 0}}; /* rb, Q10, ta */
    static const int32_t loss[1] = {EH_LOSS_MSE};
    eh_model_desc d;
    memset(&d, 0, sizeof d);
    chain.n_in = 2; chain.in_cols = in_cols; chain.n_hidden = 2; chain.hidden = hidden; chain.n_out = 1;
    chain.activation = EH_ACT_TANH; chain.input_batchnorm = 0;
    d.abi_version = EH_ABI_VERSION;
    d.n_pred = 2; d.n_forc = 1; d.n_targ = 1;
    d.n_chains = 1; d.chains = &chain;
    d.n_params = 2; d.role = role; d.role_index = role_index; d.deflt = deflt; d.lower = lower; d.upper = upper;
    d.scale_nn_outputs = 1;
    d.process_model = EH_PM_RBQ10; d.n_pm_args = 3; d.pm_args = pm_args; d.pm_consts[0] = 15.0f; /* tref */
    d.loss_per_target = loss; d.agg = EH_AGG_SUM;
    d.opt_kind = EH_OPT_ADAM; d.eta = 0.01f; d.beta1 = 0.9f; d.beta2 = 0.999f; d.eps = 1e-8f;
    d.adamw_decay_coupled_eta = 1;
    d.device = 0; d.flags = 0;
    return eh_create(out, &d);
}

int main(int argc, char** argv)
{
    const char* stem = argc > 1 ? argv[1] : "abi_replay_out";
    const int64_t n = 6000, nval = 1500, B = 512, nepochs = 2;
    char path[1024];
    eh_ctx* ctx = NULL;

    /* ---- synthetic data (only + and *: nothing here depends on the host libm) ---- */
    float* X = malloc(sizeof(float) * 2 * (n + nval));
    float* ta = malloc(sizeof(float) * (n + nval));
    float* reco = malloc(sizeof(float) * (n + nval));
    for (int64_t i = 0; i < n + nval; i++) {
        const float sw = 20.0f + 60.0f * urand(), dsw = 4.0f * (urand() - 0.5f), t = -5.0f + 30.0f * urand();
        const float d = t - 15.0f, q = 1.0f + 0.07f * d + 0.0025f * d * d;
        X[2 * i] = sw; X[2 * i + 1] = dsw; ta[i] = t;
        reco[i] = (3.0f + 0.02f * (sw - 50.0f)) * q + 0.1f * (urand() - 0.5f);
        if (i % 97 == 5) reco[i] = NAN; /* missing target: valid_mask, src/training/train.jl:221-232 */
    }

    if (create_rbq10(&ctx) != EH_OK) {
        fprintf(stderr, "eh_create failed: %s\n", eh_last_error(NULL));
        return 3; /* no sm_100 device: the test skips */
    }
    const int64_t nflat = eh_num_params(ctx);
    float* flat0 = malloc(sizeof(float) * nflat);
    for (int64_t i = 0; i < nflat; i++) flat0[i] = 0.6f * (urand() - 0.5f);
    int64_t* perm = malloc(sizeof(int64_t) * n * nepochs);
    for (int64_t e = 0; e < nepochs; e++) { /* DataLoader(shuffle = true): a fresh permutation per epoch, 1-based */
        int64_t* p = perm + e * n;
        for (int64_t i = 0; i < n; i++) p[i] = i + 1;
        for (int64_t i = n - 1; i > 0; i--) {
            lcg_state = lcg_state * 1664525u + 1013904223u;
            const int64_t j = (int64_t)((lcg_state >> 8) % (uint32_t)(i + 1));
            const int64_t t = p[i]; p[i] = p[j]; p[j] = t;
        }
    }
    snprintf(path, sizeof path, "%s.bin", stem);
    FILE* fb = fopen(path, "wb");
    if (!fb) return 4;
    const int64_t hdr[5] = {n, nval, B, nepochs, nflat};
    fwrite(hdr, sizeof hdr, 1, fb);
    fwrite(X, sizeof(float), 2 * (n + nval), fb);
    fwrite(ta, sizeof(float), n + nval, fb);
    fwrite(reco, sizeof(float), n + nval, fb);
    fwrite(flat0, sizeof(float), nflat, fb);
    fwrite(perm, sizeof(int64_t), n * nepochs, fb);
    fclose(fb);
    snprintf(path, sizeof path, "%s.txt", stem);
    FILE* fo = fopen(path, "w");
    if (!fo) return 4;

    /* ---- the resident path: what run_epoch! becomes ---- */
    const float* forc_tr[1] = {ta};
    const float* targ_tr[1] = {reco};
    const float* forc_va[1] = {ta + n};
    const float* targ_va[1] = {reco + n};
    CHECK(eh_upload(ctx, EH_SPLIT_TRAIN, n, X, forc_tr, targ_tr));
    CHECK(eh_upload(ctx, EH_SPLIT_VAL, nval, X + 2 * n, forc_va, targ_va));
    CHECK(eh_set_params(ctx, flat0, nflat));
    fprintf(fo, "variant %s\n", eh_kernel_variant(ctx));
    fprintf(fo, "nflat %lld\n", (long long)nflat);

    float loss0 = 0.f;
    float* grad = malloc(sizeof(float) * nflat);
    CHECK(eh_loss_grad(ctx, perm, B, &loss0, grad));
    fprintf(fo, "loss0 %.9g\ngrad0", loss0);
    for (int64_t i = 0; i < nflat; i++) fprintf(fo, " %.9g", grad[i]);
    fprintf(fo, "\n");

    const int64_t nb = (n + B - 1) / B;
    float* losses = malloc(sizeof(float) * nb * nepochs);
    for (int64_t e = 0; e < nepochs; e++) CHECK(eh_epoch(ctx, perm + e * n, n, B, losses + e * nb));
    fprintf(fo, "epoch_losses");
    for (int64_t i = 0; i < nb * nepochs; i++) fprintf(fo, " %.9g", losses[i]);
    fprintf(fo, "\n");

    float* flat1 = malloc(sizeof(float) * nflat);
    float* m = malloc(sizeof(float) * nflat);
    float* v = malloc(sizeof(float) * nflat);
    int64_t t_steps = 0;
    CHECK(eh_get_params(ctx, flat1, nflat));
    CHECK(eh_get_opt_state(ctx, m, v, nflat, &t_steps));
    fprintf(fo, "steps %lld\nparams", (long long)t_steps);
    for (int64_t i = 0; i < nflat; i++) fprintf(fo, " %.9g", flat1[i]);
    fprintf(fo, "\n");

    double stats[EH_EVAL_STATS];
    float* yhat = malloc(sizeof(float) * (n > nval ? n : nval));
    CHECK(eh_eval(ctx, EH_SPLIT_VAL, yhat, stats, NULL));
    fprintf(fo, "val_stats");
    for (int i = 0; i < EH_EVAL_STATS; i++) fprintf(fo, " %.17g", stats[i]);
    fprintf(fo, "\nval_yhat_head");
    for (int i = 0; i < 8; i++) fprintf(fo, " %.9g", yhat[i]);
    fprintf(fo, "\n");
    eh_destroy(ctx);

    /* ---- the per-step path: single_train_step! on host batches (collect_dim_data, epoch.jl:1-11) ---- */
    if (create_rbq10(&ctx) != EH_OK) return 3;
    CHECK(eh_set_params(ctx, flat0, nflat));
    float* bx = malloc(sizeof(float) * 2 * B);
    float* bt = malloc(sizeof(float) * B);
    float* by = malloc(sizeof(float) * B);
    fprintf(fo, "host_losses");
    for (int64_t e = 0; e < nepochs; e++)
        for (int64_t k = 0; k < nb; k++) {
            const int64_t b0 = k * B, bn = (b0 + B <= n ? B : n - b0);
            for (int64_t i = 0; i < bn; i++) {
                const int64_t s = perm[e * n + b0 + i] - 1;
                bx[2 * i] = X[2 * s]; bx[2 * i + 1] = X[2 * s + 1]; bt[i] = ta[s]; by[i] = reco[s];
            }
            const float* bf[1] = {bt};
            const float* bg[1] = {by};
            float l = 0.f;
            CHECK(eh_step_host(ctx, bn, bx, bf, bg, &l));
            fprintf(fo, " %.9g", l);
        }
    fprintf(fo, "\n");
    CHECK(eh_get_params(ctx, flat1, nflat));
    fprintf(fo, "host_params");
    for (int64_t i = 0; i < nflat; i++) fprintf(fo, " %.9g", flat1[i]);
    fprintf(fo, "\n");
    eh_destroy(ctx);
    fclose(fo);
    printf("abi_replay ok: %s.txt\n", stem);
    return 0;
}
