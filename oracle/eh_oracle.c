/*
 * eh_oracle.c -- CPU ORACLE for the EasyHybrid hybrid training step.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke()
 * check in __graft_entry__.py and the cpu_baseline / --impl reference legs of
 * bench.py may load it.  The product (libeasyhybrid_cuda.so) never links,
 * loads or calls anything in this directory.
 *
 * PARITY STATUS: PARTLY PINNED.
 *   pinned   : loss_fn closed forms incl. mask (test/test_loss_fn.jl:6-8,15-74,
 *              90-145), scale_single_param / minmax / hard_sigmoid known answers
 *              (test/test_generic_hybrid_model.jl:24-35,109-126), the
 *              _compute_loss sum-over-targets identities
 *              (test/test_compute_loss.jl:49-79) -- see tests/golden/.
 *   UNPINNED : Dense/BatchNorm numerics, Zygote gradients, Optimisers updates,
 *              DataLoader batch order, splitobs indices.  The reference is
 *              Julia; neither julia nor its un-vendored dependencies (Lux
 *              1.21, Zygote 0.7, Optimisers via OptimizationOptimisers 0.3.7,
 *              MLUtils 0.4.8; Project.toml:46-78, no Manifest) exist in this
 *              environment, and the reference's own tests assert no numeric
 *              value for them.  For those parts this file
 *              restates the published algorithms and is
 *              cross-witnessed by an independent torch-CPU float64 autograd
 *              implementation in tests/test_oracle_witness.py.
 *              julia/parity_dump.jl produces the real Lux/Zygote trace for
 *              anyone with Julia.
 *
 * What it restates (reference file:line):
 *   forward              src/models/GenericHybridModel.jl:370-431, 458-530
 *   Dense chain          src/models/NNModels.jl:220-231
 *   parameter squashing  src/models/GenericHybridModel.jl:348-365
 *   initial phi          src/models/GenericHybridModel.jl:236-256
 *   losses               src/losses/loss_fn.jl:58-179
 *   loss assembly / agg  src/losses/compute_loss.jl:20-66, 115-145
 *   NaN mask             src/training/train.jl:221-232
 *   step loop            src/training/epoch.jl:13-37
 *   optimiser defaults   src/config/TrainingConfig.jl:43
 *
 * Two precisions are compiled from eh_oracle_core.inc: float (the reference's
 * Float32 path) and double ("truth").  OpenMP parallelises over sample blocks
 * for the timed CPU baseline; the reduction order is fixed (per-block partials,
 * pairwise over blocks) so results do not depend on the thread count.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/easyhybrid_cuda.h" /* descriptor types only */

#define EHO_BLK 64
#define EHO_MAXT 16
#define EHO_MAXL 16

typedef struct {
    int n_in, n_layers, activation, input_batchnorm;
    int in_cols[64];
    int width[EHO_MAXL + 1];
    int64_t w_off[EHO_MAXL], b_off[EHO_MAXL];
    int act_off[EHO_MAXL + 1]; /* row offset of each layer's activations in ws.act */
} eho_chain;

typedef struct {
    int n_pred, n_forc, n_targ, n_chains, n_params;
    eho_chain* chains;
    int *role, *nn_chain, *nn_row, *glob_pos;
    float *deflt, *lower, *upper;
    int scale_nn_outputs;
    eh_pm_instr* prog;
    int pm_len;
    int pm_outputs[EHO_MAXT];
    int loss[EHO_MAXT];
    int agg;
    int opt_kind, adamw_coupled;
    float eta, beta1, beta2, eps, lambda;
    /* native extra loss lambda * weight_l2(ps.<chains>; normalize) (extract_weights.jl:55-91, compute_loss.jl:31-34) */
    float l2_lambda; int l2_normalize; unsigned l2_mask;
    int64_t n_flat, phi_off;
    int n_glob;
    int tot_act, maxw, tot_in, any_bn;
    /* BN running state, [tot_in] each */
    float *bn_rmean, *bn_rvar;
} eho_plan;

typedef struct {
    int64_t N;
    const float* X;            /* P x N column-major */
    const float* const* forc;  /* F x [N] */
    const float* const* targ;  /* T x [N] */
} eho_data;

/* ---- built-in process models as canned programs ------------- */
static int emit(eh_pm_instr* p, int* n, int op, int a, int b, float imm)
{
    p[*n].op = op; p[*n].a = a; p[*n].b = b; p[*n].imm = imm;
    return (*n)++;
}
static int emit_arg(eh_pm_instr* p, int* n, const eh_pm_arg* a)
{
    return emit(p, n, a->kind == 0 ? EH_OP_PARAM : EH_OP_FORCING, a->index, 0, 0.f);
}

static int build_program(eho_plan* pl, const eh_model_desc* d)
{
    if (d->process_model == EH_PM_PROGRAM) {
        pl->pm_len = d->pm_len;
        pl->prog = (eh_pm_instr*)malloc(sizeof(eh_pm_instr) * (size_t)(d->pm_len + 1));
        memcpy(pl->prog, d->pm_prog, sizeof(eh_pm_instr) * (size_t)d->pm_len);
        for (int t = 0; t < d->n_targ; t++) pl->pm_outputs[t] = d->pm_outputs[t];
        return 0;
    }
    eh_pm_instr* p = (eh_pm_instr*)calloc(32, sizeof(eh_pm_instr));
    int n = 0;
    const eh_pm_arg* a = d->pm_args;
    switch (d->process_model) {
    case EH_PM_RBQ10: {
        /* reco = rb .* Q10 .^ (0.1f0 .* (ta .- tref))  (README.md:148-151) */
        if (d->n_pm_args != 3 || d->n_targ != 1) return -1;
        int rb = emit_arg(p, &n, &a[0]), q10 = emit_arg(p, &n, &a[1]), ta = emit_arg(p, &n, &a[2]);
        int tref = emit(p, &n, EH_OP_CONST, 0, 0, d->pm_consts[0]);
        int c01 = emit(p, &n, EH_OP_CONST, 0, 0, 0.1f);
        int dt = emit(p, &n, EH_OP_SUB, ta, tref, 0);
        int e = emit(p, &n, EH_OP_MUL, c01, dt, 0);
        int pw = emit(p, &n, EH_OP_POW, q10, e, 0);
        pl->pm_outputs[0] = emit(p, &n, EH_OP_MUL, rb, pw, 0);
    } break;
    case EH_PM_EXPO: {
        /* Resp_obs = Resp0 .* exp.(k .* T)  (projects/ExpoHybrid/ExpoHybridEstim.jl:69-85) */
        if (d->n_pm_args != 3 || d->n_targ != 1) return -1;
        int r0 = emit_arg(p, &n, &a[0]), k = emit_arg(p, &n, &a[1]), T = emit_arg(p, &n, &a[2]);
        int kt = emit(p, &n, EH_OP_MUL, k, T, 0);
        int ex = emit(p, &n, EH_OP_EXP, kt, 0, 0);
        pl->pm_outputs[0] = emit(p, &n, EH_OP_MUL, r0, ex, 0);
    } break;
    case EH_PM_EXPO2: {
        /* (var1 = Resp0 .* exp.(k .* T), var2 = 2 .* var1): the two-target form of the Expo model used by the
         * wide-MLP configuration (same construction as test/test_compute_loss.jl:209-211 for the linear model) */
        if (d->n_pm_args != 3 || d->n_targ != 2) return -1;
        int r0 = emit_arg(p, &n, &a[0]), k = emit_arg(p, &n, &a[1]), T = emit_arg(p, &n, &a[2]);
        int kt = emit(p, &n, EH_OP_MUL, k, T, 0);
        int ex = emit(p, &n, EH_OP_EXP, kt, 0, 0);
        int two = emit(p, &n, EH_OP_CONST, 0, 0, 2.0f);
        pl->pm_outputs[0] = emit(p, &n, EH_OP_MUL, r0, ex, 0);
        pl->pm_outputs[1] = emit(p, &n, EH_OP_MUL, two, pl->pm_outputs[0], 0);
    } break;
    case EH_PM_LINEAR:
    case EH_PM_LINEAR2: {
        /* a .* x1 .+ b (test/test_generic_hybrid_model.jl:10-12);
         * (var1 = a.*x1.+b, var2 = 2a.*x1.+b) (test/test_compute_loss.jl:209-211) */
        int nt = d->process_model == EH_PM_LINEAR ? 1 : 2;
        if (d->n_pm_args != 3 || d->n_targ != nt) return -1;
        int av = emit_arg(p, &n, &a[0]), bv = emit_arg(p, &n, &a[1]), x = emit_arg(p, &n, &a[2]);
        int ax = emit(p, &n, EH_OP_MUL, av, x, 0);
        pl->pm_outputs[0] = emit(p, &n, EH_OP_ADD, ax, bv, 0);
        if (nt == 2) {
            int two = emit(p, &n, EH_OP_CONST, 0, 0, 2.0f);
            int a2 = emit(p, &n, EH_OP_MUL, two, av, 0);
            int a2x = emit(p, &n, EH_OP_MUL, a2, x, 0);
            pl->pm_outputs[1] = emit(p, &n, EH_OP_ADD, a2x, bv, 0);
        }
    } break;
    default: free(p); return -1;
    }
    pl->prog = p;
    pl->pm_len = n;
    return 0;
}

void eho_plan_free(eho_plan* pl)
{
    if (!pl) return;
    free(pl->chains); free(pl->role); free(pl->nn_chain); free(pl->nn_row); free(pl->glob_pos);
    free(pl->deflt); free(pl->lower); free(pl->upper); free(pl->prog);
    free(pl->bn_rmean); free(pl->bn_rvar);
    free(pl);
}

eho_plan* eho_plan_new(const eh_model_desc* d)
{
    if (!d || d->n_targ > EHO_MAXT || d->n_targ < 1) return NULL;
    eho_plan* pl = (eho_plan*)calloc(1, sizeof(*pl));
    pl->n_pred = d->n_pred; pl->n_forc = d->n_forc; pl->n_targ = d->n_targ;
    pl->n_chains = d->n_chains; pl->n_params = d->n_params;
    pl->chains = (eho_chain*)calloc((size_t)(d->n_chains > 0 ? d->n_chains : 1), sizeof(eho_chain));
    int64_t off = 0;
    int act = 0, maxw = 1, tot_in = 0;
    for (int c = 0; c < d->n_chains; c++) {
        const eh_chain_desc* cd = &d->chains[c];
        eho_chain* ch = &pl->chains[c];
        if (cd->n_hidden + 1 > EHO_MAXL || cd->n_in > 64) { eho_plan_free(pl); return NULL; }
        ch->n_in = cd->n_in; ch->n_layers = cd->n_hidden + 1;
        ch->activation = cd->activation; ch->input_batchnorm = cd->input_batchnorm;
        if (cd->input_batchnorm) pl->any_bn = 1;
        for (int k = 0; k < cd->n_in; k++) ch->in_cols[k] = cd->in_cols[k];
        ch->width[0] = cd->n_in;
        for (int l = 0; l < cd->n_hidden; l++) ch->width[l + 1] = cd->hidden[l];
        ch->width[ch->n_layers] = cd->n_out;
        for (int l = 0; l <= ch->n_layers; l++) {
            ch->act_off[l] = act; act += ch->width[l];
            if (ch->width[l] > maxw) maxw = ch->width[l];
        }
        for (int l = 0; l < ch->n_layers; l++) {
            ch->w_off[l] = off; off += (int64_t)ch->width[l] * ch->width[l + 1];
            ch->b_off[l] = off; off += ch->width[l + 1];
        }
        tot_in += cd->n_in;
    }
    pl->tot_act = act; pl->maxw = maxw; pl->tot_in = tot_in;
    pl->phi_off = off;
    pl->role = (int*)calloc((size_t)d->n_params, sizeof(int));
    pl->nn_chain = (int*)calloc((size_t)d->n_params, sizeof(int));
    pl->nn_row = (int*)calloc((size_t)d->n_params, sizeof(int));
    pl->glob_pos = (int*)calloc((size_t)d->n_params, sizeof(int));
    pl->deflt = (float*)calloc((size_t)d->n_params, sizeof(float));
    pl->lower = (float*)calloc((size_t)d->n_params, sizeof(float));
    pl->upper = (float*)calloc((size_t)d->n_params, sizeof(float));
    int ng = 0;
    for (int p = 0; p < d->n_params; p++) {
        pl->role[p] = d->role[p];
        pl->deflt[p] = d->deflt[p]; pl->lower[p] = d->lower[p]; pl->upper[p] = d->upper[p];
        if (d->role[p] == EH_ROLE_NEURAL) { pl->nn_chain[p] = d->role_index[p] >> 16; pl->nn_row[p] = d->role_index[p] & 0xffff; }
        if (d->role[p] == EH_ROLE_GLOBAL) { pl->glob_pos[p] = d->role_index[p]; if (d->role_index[p] + 1 > ng) ng = d->role_index[p] + 1; }
    }
    pl->n_glob = ng;
    pl->n_flat = off + ng;
    pl->scale_nn_outputs = d->scale_nn_outputs;
    if (build_program(pl, d) != 0) { eho_plan_free(pl); return NULL; }
    for (int t = 0; t < d->n_targ; t++) pl->loss[t] = d->loss_per_target[t];
    pl->agg = d->agg;
    pl->opt_kind = d->opt_kind; pl->adamw_coupled = d->adamw_decay_coupled_eta;
    pl->eta = d->eta; pl->beta1 = d->beta1; pl->beta2 = d->beta2; pl->eps = d->eps; pl->lambda = d->lambda;
    if (d->abi_version >= 2) { pl->l2_lambda = d->l2_lambda; pl->l2_normalize = d->l2_normalize; pl->l2_mask = d->l2_chain_mask; }
    pl->bn_rmean = (float*)calloc((size_t)(tot_in > 0 ? tot_in : 1), sizeof(float));
    pl->bn_rvar = (float*)calloc((size_t)(tot_in > 0 ? tot_in : 1), sizeof(float));
    for (int i = 0; i < tot_in; i++) pl->bn_rvar[i] = 1.0f; /* Lux BatchNorm initial running_var = 1 */
    return pl;
}

int64_t eho_num_params(const eho_plan* pl) { return pl->n_flat; }

/* ---- instantiate the core twice ------------------------------------------ */
#define REAL float
#define SUF f32
#define RTANH tanhf
#define REXP expf
#define RLOG logf
#define RSQRT sqrtf
#define RSIN sinf
#define RCOS cosf
#include "eh_oracle_core.inc"
#undef REAL
#undef SUF
#undef RTANH
#undef REXP
#undef RLOG
#undef RSQRT
#undef RSIN
#undef RCOS

#define REAL double
#define SUF f64
#define RTANH tanh
#define REXP exp
#define RLOG log
#define RSQRT sqrt
#define RSIN sin
#define RCOS cos
#include "eh_oracle_core.inc"
#undef REAL
#undef SUF

static int nthreads_or_default(int n)
{
#ifdef _OPENMP
    return n > 0 ? n : omp_get_max_threads();
#else
    (void)n; return 1;
#endif
}
int eho_max_threads(void) { return nthreads_or_default(0); }

/* ---- public: loss + gradient on a batch ----------------------------------- */
/* idx0: 0-based indices (NULL = first B samples).  precision: 32 or 64.
 * flat is float in both cases (the stored parameters are Float32).          */
double eho_loss_grad(const eho_plan* pl, const float* flat, int64_t N, const float* X, const float* const* forc,
                     const float* const* targ, const int64_t* idx0, int64_t B, double* grad_out, int precision,
                     int nthreads)
{
    eho_data d = {N, X, forc, targ};
    int64_t nv = 0;
    nthreads = nthreads_or_default(nthreads);
    double L;
    if (precision == 64) {
        double* f = (double*)malloc(sizeof(double) * (size_t)pl->n_flat);
        double* g = (double*)malloc(sizeof(double) * (size_t)pl->n_flat);
        for (int64_t i = 0; i < pl->n_flat; i++) f[i] = flat[i];
        L = loss_grad_f64(pl, f, &d, idx0, B, g, &nv, NULL, nthreads);
        if (grad_out) for (int64_t i = 0; i < pl->n_flat; i++) grad_out[i] = g[i];
        free(f); free(g);
    } else {
        float* g = (float*)malloc(sizeof(float) * (size_t)pl->n_flat);
        L = loss_grad_f32(pl, flat, &d, idx0, B, g, &nv, NULL, nthreads);
        if (grad_out) for (int64_t i = 0; i < pl->n_flat; i++) grad_out[i] = g[i];
        free(g);
    }
    return L;
}

/* ---- public: run_epoch!-equivalent (Float32) ------------------------------ */
/* perm0: 0-based permutation of length n; batches of B, last partial; all-masked
 * batches skipped (epoch.jl:17-19).  Updates flat/m/v/t in place.  losses may be NULL. */
int eho_train_steps(eho_plan* pl, float* flat, float* m, float* v, int64_t* t, int64_t N, const float* X,
                    const float* const* forc, const float* const* targ, const int64_t* perm0, int64_t n, int64_t B,
                    float* losses, int nthreads)
{
    eho_data d = {N, X, forc, targ};
    if (!perm0) return -1;
    nthreads = nthreads_or_default(nthreads);
    float* g = (float*)malloc(sizeof(float) * (size_t)pl->n_flat);
    float* bnb = pl->any_bn ? (float*)calloc((size_t)2 * pl->tot_in, sizeof(float)) : NULL;
    int64_t nsteps = (n + B - 1) / B;
    for (int64_t k = 0; k < nsteps; k++) {
        int64_t b = (k + 1) * B <= n ? B : n - k * B;
        int64_t nv = 0;
        double L = loss_grad_f32(pl, flat, &d, perm0 + k * B, b, g, &nv, bnb, nthreads);
        if (nv == 0) { if (losses) losses[k] = NAN; continue; }
        if (losses) losses[k] = (float)L;
        opt_step_f32(pl, flat, m, v, t, g);
        if (bnb) { /* Lux BatchNorm running stats: momentum 0.1, unbiased variance */
            for (int i = 0; i < pl->tot_in; i++) {
                float mu = bnb[2 * i], var = bnb[2 * i + 1];
                float unb = b > 1 ? var * (float)b / (float)(b - 1) : var;
                pl->bn_rmean[i] = 0.9f * pl->bn_rmean[i] + 0.1f * mu;
                pl->bn_rvar[i] = 0.9f * pl->bn_rvar[i] + 0.1f * unb;
            }
        }
    }
    free(g); free(bnb);
    return 0;
}

/* optimiser step alone (for unit tests) */
void eho_opt_step(const eho_plan* pl, float* flat, float* m, float* v, int64_t* t, const float* g)
{
    opt_step_f32(pl, flat, m, v, t, g);
}

void eho_get_bn_state(const eho_plan* pl, float* mean, float* var)
{
    memcpy(mean, pl->bn_rmean, sizeof(float) * (size_t)pl->tot_in);
    memcpy(var, pl->bn_rvar, sizeof(float) * (size_t)pl->tot_in);
}
void eho_set_bn_state(eho_plan* pl, const float* mean, const float* var)
{
    memcpy(pl->bn_rmean, mean, sizeof(float) * (size_t)pl->tot_in);
    memcpy(pl->bn_rvar, var, sizeof(float) * (size_t)pl->tot_in);
}

/* ---- public: test-mode forward on a whole split (evaluate_acc) ------------ */
/* yhat: [T][N]; par_out (nullable): [n_params][N] scaled process parameters */
int eho_forward(const eho_plan* pl, const float* flat, int64_t N, const float* X, const float* const* forc,
                float* yhat, float* par_out, int precision, int nthreads)
{
    eho_data d = {N, X, forc, NULL};
    nthreads = nthreads_or_default(nthreads);
    const int B = EHO_BLK;
    int64_t nblk = (N + B - 1) / B;
    int tot_in = pl->tot_in;
    if (precision == 64) {
        double* f = (double*)malloc(sizeof(double) * (size_t)pl->n_flat);
        for (int64_t i = 0; i < pl->n_flat; i++) f[i] = flat[i];
        double* mu = (double*)calloc((size_t)(tot_in + 1), sizeof(double));
        double* rs = (double*)calloc((size_t)(tot_in + 1), sizeof(double));
        for (int i = 0; i < tot_in; i++) { mu[i] = pl->bn_rmean[i]; rs[i] = 1.0 / sqrt((double)pl->bn_rvar[i] + 1e-5); }
#pragma omp parallel num_threads(nthreads)
        {
            ws_t_f64* w = ws_new_f64(pl);
            double* rec = (double*)calloc((size_t)(pl->n_pred + pl->n_forc) * B, sizeof(double));
#pragma omp for schedule(static)
            for (int64_t b = 0; b < nblk; b++) {
                int64_t base = b * B;
                int nb = (int)((N - base) < B ? (N - base) : B);
                gather_f64(pl, &d, NULL, base, nb, rec, NULL);
                block_forward_f64(pl, f, w, rec, pl->any_bn ? mu : NULL, rs, nb);
                for (int t = 0; t < pl->n_targ; t++)
                    for (int s = 0; s < nb; s++) yhat[(size_t)t * N + base + s] = (float)w->yhat[(size_t)t * B + s];
                if (par_out)
                    for (int p = 0; p < pl->n_params; p++)
                        for (int s = 0; s < nb; s++) par_out[(size_t)p * N + base + s] = (float)w->par[(size_t)p * B + s];
            }
            free(rec); ws_free_f64(w);
        }
        free(f); free(mu); free(rs);
    } else {
        float* mu = (float*)calloc((size_t)(tot_in + 1), sizeof(float));
        float* rs = (float*)calloc((size_t)(tot_in + 1), sizeof(float));
        for (int i = 0; i < tot_in; i++) { mu[i] = pl->bn_rmean[i]; rs[i] = 1.0f / sqrtf(pl->bn_rvar[i] + 1e-5f); }
#pragma omp parallel num_threads(nthreads)
        {
            ws_t_f32* w = ws_new_f32(pl);
            float* rec = (float*)calloc((size_t)(pl->n_pred + pl->n_forc) * B, sizeof(float));
#pragma omp for schedule(static)
            for (int64_t b = 0; b < nblk; b++) {
                int64_t base = b * B;
                int nb = (int)((N - base) < B ? (N - base) : B);
                gather_f32(pl, &d, NULL, base, nb, rec, NULL);
                block_forward_f32(pl, flat, w, rec, pl->any_bn ? mu : NULL, rs, nb);
                for (int t = 0; t < pl->n_targ; t++)
                    for (int s = 0; s < nb; s++) yhat[(size_t)t * N + base + s] = w->yhat[(size_t)t * B + s];
                if (par_out)
                    for (int p = 0; p < pl->n_params; p++)
                        for (int s = 0; s < nb; s++) par_out[(size_t)p * N + base + s] = w->par[(size_t)p * B + s];
            }
            free(rec); ws_free_f32(w);
        }
        free(mu); free(rs);
    }
    return 0;
}

/* ---- public: the reference's loss_fn table in Float64 --------------------- */
/* src/losses/loss_fn.jl:58-179.  kind: 0 mse 1 rmse 2 mae 3 nseLoss 4 nse 5 r2
 * 6 pearson 7 pearsonLoss 8 kgeLoss 9 kge 10 pbkgeLoss 11 pbkge 12 alpha 13 beta */
double eho_loss_fn(int kind, const double* yhat, const double* y, const uint8_t* mask, int64_t n)
{
    double nv = 0, sy = 0, sh = 0;
    for (int64_t i = 0; i < n; i++) if (mask[i]) { nv += 1; sy += y[i]; sh += yhat[i]; }
    double my = sy / nv, mh = sh / nv;
    double sse = 0, sae = 0, syy = 0, shh = 0, syh = 0;
    for (int64_t i = 0; i < n; i++) if (mask[i]) {
        double r = yhat[i] - y[i];
        sse += r * r; sae += fabs(r);
        syy += (y[i] - my) * (y[i] - my); shh += (yhat[i] - mh) * (yhat[i] - mh);
        syh += (y[i] - my) * (yhat[i] - mh);
    }
    double cor = syh / sqrt(syy * shh);
    double sd_o = sqrt(syy / (nv - 1)), sd_s = sqrt(shh / (nv - 1)); /* Statistics.std: corrected */
    double alpha = sd_s / sd_o, beta = mh / my;
    switch (kind) {
    case 0: return sse / nv;
    case 1: return sqrt(sse / nv);
    case 2: return sae / nv;
    case 3: return sse / syy;
    case 4: return 1.0 - sse / syy;
    case 5: return 1.0 - sse / syy;
    case 6: return cor;
    case 7: return 1.0 - cor;
    case 8: return sqrt((cor - 1) * (cor - 1) + (alpha - 1) * (alpha - 1) + (beta - 1) * (beta - 1));
    case 9: return 1.0 - sqrt((cor - 1) * (cor - 1) + (alpha - 1) * (alpha - 1) + (beta - 1) * (beta - 1));
    case 10: return sqrt((cor - 1) * (cor - 1) + (beta - 1) * (beta - 1));
    case 11: return 1.0 - sqrt((cor - 1) * (cor - 1) + (beta - 1) * (beta - 1));
    case 12: return alpha;
    case 13: return beta;
    default: return NAN;
    }
}

/* ---- public: parameter squashing helpers (GenericHybridModel.jl:9-18,348-365) */
float eho_scale_single_param(float raw, float lower, float upper)
{
    return lower + (upper - lower) * (1.0f / (1.0f + expf(-raw)));
}
float eho_inv_sigmoid(float y) { return logf(y / (1.0f - y)); }
float eho_scale_single_param_minmax(float deflt, float lower, float upper)
{
    return eho_inv_sigmoid((deflt - lower) / (upper - lower));
}
double eho_hard_sigmoid(double x)
{
    double v = 0.2 * x + 0.5;
    return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
}
