// eh_update_kernel.cuh -- K2: fixed-order second-pass reduction + loss scalars +
// optimiser update (one CTA), K0: per-batch data statistics, and the record packer.
//
// K2 replaces `Optimisers.update!(opt_state, ps, grads)` inside
// Lux.Training.single_train_step! (call site src/training/epoch.jl:20-26) and the
// scalar part of the loss (src/losses/loss_fn.jl:58-81, compute_loss.jl:50-53).
// Optimiser rules follow Optimisers.jl (unpinned).
#pragma once
#include "eh_step_kernel.cuh"

namespace eh {

struct UpdateArgs {
    const float* partial;  // [G][npart] from K1
    int G, npart, npart_dw;
    float* gvec;           // [npart] reduced vector (DP exchange buffer / UPD_FROM_VECTOR input)
    int mode;              // UPD_*
    int apply;             // 0: loss + gradient only (eh_loss_grad)
    int nflat, ntheta;     // flat entries; the first ntheta are chain weights, the rest phi
    const int* pmap;       // [nflat] index into the partial vector
    const float* pspan;    // [nflat] phi entries: (upper - lower); 0 for theta
    float* theta;
    float* m;
    float* v;
    OptState* ost;
    const float* bscal;    // this batch's scalar row
    float* loss_out;       // device slot for this step's loss
    float* grad_out;       // nullable: flat gradient of this step
    int T, agg_mean;
    int loss_kind[MAXT];
    int opt_kind, adamw_coupled;
    float eta, beta1, beta2, eps, lambda;
    // native extra loss lambda * weight_l2 (extract_weights.jl:55-91): loss = agg([L, E]), E = l2_loss_coef * sum over the
    // flagged entries of theta^2; gradient: aggw * (g + l2coef[p] * theta[p]) with l2coef = 2 * l2_loss_coef on the flagged entries
    const float* l2coef;   // [nflat] or NULL
    float l2_aggw, l2_loss_coef;
    // parameter-block tail (uniform slot values + derived process-model scalars)
    const int* slot_of_flat;  // [nflat] phi entries: canonical slot, -1 otherwise
    PSlot slot[MAXPS];
    int pm_id;
};

// value + derived scalars of uniform slot sl -> tail of the parameter block
__device__ inline void write_slot_tail(float* pblock, int nflat, int pm_id, int sl, float val)
{
    float o4[4];
    pm_prep_slot(pm_id, sl, val, o4);
    pblock[nflat + sl] = val;
    for (int i = 0; i < 4; i++) pblock[nflat + MAXPS + sl * PMS_PER_SLOT + i] = o4[i];
}

// refresh the whole tail from the current phi (after eh_set_params)
struct TailArgs {
    float* pblock;
    int nflat, ntheta, pm_id;
    const int* slot_of_flat;
    PSlot slot[MAXPS];
};
__global__ void k_param_tail(const TailArgs a)
{
    int t = threadIdx.x;
    if (t < MAXPS) {
        const PSlot sl = a.slot[t];
        if (sl.role == ROLE_FIXED) write_slot_tail(a.pblock, a.nflat, a.pm_id, t, sl.fixedv);
        else if (sl.role == ROLE_NEURAL) write_slot_tail(a.pblock, a.nflat, -1, t, 0.f);
    }
    for (int p = a.ntheta + t; p < a.nflat; p += blockDim.x) {
        int s = a.slot_of_flat[p];
        if (s >= 0) {
            const PSlot sl = a.slot[s];
            write_slot_tail(a.pblock, a.nflat, a.pm_id, s, sl.lo + sl.span * (1.f / (1.f + expf(-a.pblock[p]))));
        }
    }
}

__global__ void __launch_bounds__(512, 1) k_update(const UpdateArgs a)
{
    __shared__ float red[UPD_MAX_NPART];
    __shared__ float s_loss, s_post;
    __shared__ int s_skip;
    __shared__ float s_l2w[16];
    pdl_wait();  // K1 of this step must be complete and flushed
    pdl_launch_dependents();  // let the next K1 start its data prefetch
    if (a.l2coef) {
        // sum of squares of the flagged weights (before this step's update), fixed order: thread stride, warp tree, 16 warps
        float s = 0.f;
        for (int p = threadIdx.x; p < a.nflat; p += blockDim.x)
            if (a.l2coef[p] != 0.f) s = fmaf(a.theta[p], a.theta[p], s);
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) s_l2w[threadIdx.x >> 5] = s;
    }

    if (a.mode != UPD_FROM_VECTOR) {
        for (int p = threadIdx.x; p < a.npart; p += blockDim.x) {
            // fixed order over CTAs -> bitwise reproducible
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int g = 0;
            for (; g + 4 <= a.G; g += 4) {
                s0 += a.partial[(size_t)(g + 0) * a.npart + p];
                s1 += a.partial[(size_t)(g + 1) * a.npart + p];
                s2 += a.partial[(size_t)(g + 2) * a.npart + p];
                s3 += a.partial[(size_t)(g + 3) * a.npart + p];
            }
            for (; g < a.G; g++) s0 += a.partial[(size_t)g * a.npart + p];
            float s = (s0 + s1) + (s2 + s3);
            red[p] = s;
            if (a.mode == UPD_REDUCE_ONLY) a.gvec[p] = s;
        }
        if (a.mode == UPD_REDUCE_ONLY) return;
    } else {
        for (int p = threadIdx.x; p < a.npart; p += blockDim.x) red[p] = a.gvec[p];
    }
    __syncthreads();

    if (threadIdx.x == 0) {
        // loss_fn (loss_fn.jl:58-81) from the reduced sums, agg over targets (compute_loss.jl:50-53)
        float L = 0.f, ntot = 0.f, post = 1.f;
        for (int t = 0; t < a.T; t++) {
            float n = a.bscal[BS_N + t], ss = a.bscal[BS_SS + t], acc = red[a.npart_dw + t];
            ntot += n;
            float lt;
            switch (a.loss_kind[t]) {
            case LOSS_MSE: lt = acc / n; break;
            case LOSS_RMSE:
                lt = sqrtf(acc / n);
                post = 1.f / (2.f * lt);  // d sqrt(mse) = d mse / (2 rmse); single-target only (host checks)
                break;
            case LOSS_MAE: lt = acc / n; break;
            case LOSS_AFFINE: lt = a.bscal[BS_AFF + 3 * MAXT + t]; break;  // value left by the pre-pass (k_stat_seeds)
            default: lt = acc / ss; break;  // nseLoss = SSE / SS_tot
            }
            L += lt;
        }
        if (a.agg_mean) L /= (float)a.T;
        if (a.l2coef) {
            float s = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += s_l2w[w];
            L = a.l2_aggw * (L + a.l2_loss_coef * s);
        }
        // all-masked batch: skipped by run_epoch! (epoch.jl:17-19, 35-37)
        int skip = (ntot == 0.f);
        s_skip = skip;
        s_post = post;
        s_loss = skip ? __int_as_float(0x7fc00000) : L;
        if (a.loss_out) *a.loss_out = s_loss;
    }
    __syncthreads();
    const int skip = s_skip;
    const float post = s_post;
    const float b1t = a.ost->b1t, b2t = a.ost->b2t;
    __syncthreads();

    for (int p = threadIdx.x; p < a.nflat; p += blockDim.x) {
        float g = red[a.pmap[p]] * post;
        float th = a.theta[p];
        if (p >= a.ntheta) {
            // chain rule through scale_single_param: (u - l) sigma(raw) (1 - sigma(raw))
            float sg = 1.f / (1.f + expf(-th));
            g *= a.pspan[p] * sg * (1.f - sg);
        }
        if (a.l2coef) g = a.l2_aggw * fmaf(a.l2coef[p], th, g);
        if (a.grad_out) a.grad_out[p] = g;
        if (!a.apply || skip) continue;
        float dx;
        if (a.opt_kind == OPT_ADAM || a.opt_kind == OPT_ADAMW) {
            float mt = a.beta1 * a.m[p] + (1.f - a.beta1) * g;
            float vt = a.beta2 * a.v[p] + (1.f - a.beta2) * g * g;
            a.m[p] = mt;
            a.v[p] = vt;
            dx = mt / (1.f - b1t) / (sqrtf(vt / (1.f - b2t)) + a.eps) * a.eta;
            if (a.opt_kind == OPT_ADAMW) dx += (a.adamw_coupled ? a.eta * a.lambda : a.lambda) * th;
        } else if (a.opt_kind == OPT_RMSPROP) {
            float q = a.beta2 * a.v[p] + (1.f - a.beta2) * g * g;
            a.v[p] = q;
            dx = g * a.eta / (sqrtf(q) + a.eps);
        } else {
            dx = a.eta * g;
        }
        th -= dx;
        a.theta[p] = th;
        if (p >= a.ntheta) {
            const int s = a.slot_of_flat[p];
            if (s >= 0) write_slot_tail(a.theta, a.nflat, a.pm_id, s, a.slot[s].lo + a.slot[s].span * (1.f / (1.f + expf(-th))));
        }
    }
    if (threadIdx.x == 0 && a.apply) {
        if (skip) {
            a.ost->skipped += 1;
        } else {
            a.ost->b1t = b1t * a.beta1;
            a.ost->b2t = b2t * a.beta2;
            a.ost->t += 1;
        }
    }
}

// ---- K0: per-batch statistics that depend on the data only ------------------------
// One CTA per batch.  Per target: n_valid, SS_tot = sum (y - mean_valid(y))^2 (nseLoss,
// loss_fn.jl:79-81) and the seed scale c_t; per chain input the BatchNorm
// batch mean / rstd (Lux BatchNorm training mode, biased variance, eps = 1e-5).
struct StatArgs {
    const float* rec;   // packed records as floats
    int R4;             // floats per record
    const int* idx;     // sample ids of all batches, batch b at idx + b*Bfull (NULL: rec_base + ...)
    long long rec_base;
    long long n;        // total samples covered
    int Bfull;          // nominal batch size (last batch may be shorter)
    int P, F, T;
    float shift_y[MAXT];  // numerically convenient shift (split mean of each target)
    float shift_x[MAXP];
    int loss_kind[MAXT];
    int agg_mean;
    int use_bn;
    float* bscal;       // [nbatches][BS_STRIDE]
    float* bn_batch;    // nullable [nbatches][2*P]: raw (mean, biased var) for the running-stat update
    double* moments;    // nullable [nbatches][DP_MOMENTS]: write the raw sums instead of finalising (data-parallel
                        // mode: the caller adds them over the ranks, k_bscal_from_moments finalises)
};
// raw per-batch sums: [cnt, S1, S2] per target, [S1, S2] per chain input, number of rows
constexpr int DP_MOMENTS = 3 * MAXT + 2 * MAXP + 1;

// finalise one per-batch scalar row from (possibly rank-summed) raw sums taken with the shifts in `a`
__device__ inline void bscal_from_sums(const StatArgs& a, const double* m, int b)
{
    float* out = a.bscal + (size_t)b * BS_STRIDE;
    const double rows = m[3 * MAXT + 2 * MAXP];
    for (int t = 0; t < MAXT; t++) {
        const double c = m[3 * t], u = m[3 * t + 1], w = m[3 * t + 2];
        double sstot = c > 0 ? w - u * u / c : 0.0;
        double aggw = a.agg_mean ? 1.0 / a.T : 1.0;
        double ct = 0.0;
        if (t < a.T) ct = (a.loss_kind[t] == LOSS_NSELOSS) ? aggw / sstot : aggw / c;
        out[BS_C + t] = (float)ct;
        out[BS_N + t] = (float)c;
        out[BS_SS + t] = (float)sstot;
    }
    for (int k = 0; k < MAXP; k++) {
        const double u = m[3 * MAXT + 2 * k], w = m[3 * MAXT + 2 * k + 1];
        double mu = 0.0, var = 1.0 - 1e-5;
        if (a.use_bn && k < a.P && rows > 0) {
            mu = u / rows;
            var = w / rows - mu * mu;
            mu += (double)a.shift_x[k];
            if (var < 0) var = 0;
        }
        out[BS_BN + 2 * k] = (float)mu;
        out[BS_BN + 2 * k + 1] = (float)(1.0 / sqrt((double)(float)var + 1e-5));
        if (a.bn_batch && k < a.P) {
            a.bn_batch[(size_t)b * 2 * a.P + 2 * k] = (float)mu;
            a.bn_batch[(size_t)b * 2 * a.P + 2 * k + 1] = (float)var;
        }
    }
}

__global__ void k_bscal_from_moments(const StatArgs a, const double* moments, int nbatches)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nbatches) bscal_from_sums(a, moments + (size_t)b * DP_MOMENTS, b);
}

__global__ void __launch_bounds__(256) k_batch_stats(const StatArgs a)
{
    __shared__ double sh[8][3 * MAXT + 2 * MAXP];
    const int b = blockIdx.x;
    const long long base = (long long)b * a.Bfull;
    const int nb = (int)((a.n - base) < a.Bfull ? (a.n - base) : a.Bfull);
    double cnt[MAXT], s1[MAXT], s2[MAXT], x1[MAXP], x2[MAXP];
#pragma unroll
    for (int t = 0; t < MAXT; t++) cnt[t] = s1[t] = s2[t] = 0.0;
#pragma unroll
    for (int k = 0; k < MAXP; k++) x1[k] = x2[k] = 0.0;
    for (int s = threadIdx.x; s < nb; s += blockDim.x) {
        long long i = a.idx ? (long long)a.idx[base + s] : a.rec_base + base + s;
        const float* r = a.rec + i * a.R4;
#pragma unroll
        for (int t = 0; t < MAXT; t++)
            if (t < a.T) {
                float y = r[a.P + a.F + t];
                if (y == y) {
                    double d = (double)y - (double)a.shift_y[t];
                    cnt[t] += 1.0; s1[t] += d; s2[t] += d * d;
                }
            }
        if (a.use_bn) {
#pragma unroll
            for (int k = 0; k < MAXP; k++)
                if (k < a.P) {
                    double d = (double)r[k] - (double)a.shift_x[k];
                    x1[k] += d; x2[k] += d * d;
                }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < MAXT; t++) {
        double c = warp_sum_d(cnt[t]), u = warp_sum_d(s1[t]), w = warp_sum_d(s2[t]);
        if (lane == 0) { sh[warp][3 * t] = c; sh[warp][3 * t + 1] = u; sh[warp][3 * t + 2] = w; }
    }
#pragma unroll
    for (int k = 0; k < MAXP; k++) {
        double u = warp_sum_d(x1[k]), w = warp_sum_d(x2[k]);
        if (lane == 0) { sh[warp][3 * MAXT + 2 * k] = u; sh[warp][3 * MAXT + 2 * k + 1] = w; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        double m[DP_MOMENTS];
        for (int i = 0; i < 3 * MAXT + 2 * MAXP; i++) {
            double v = 0;
            for (int q = 0; q < nw; q++) v += sh[q][i];
            m[i] = v;
        }
        m[3 * MAXT + 2 * MAXP] = (double)nb;
        if (a.moments) {
            for (int i = 0; i < DP_MOMENTS; i++) a.moments[(size_t)b * DP_MOMENTS + i] = m[i];
        } else {
            bscal_from_sums(a, m, b);
        }
    }
}

// ---- record packer: prepare_data's ((X, forcings), targets) -> canonical AoS records ----
// canonical column c of a record comes from source plane src_plane[c] (0..P_raw-1: row of X,
// P_raw.. : forcing / target vectors stored after X) ; X is P_raw x N column-major.
struct PackArgs {
    const float* X;       // [N][P_raw]
    const float* planes;  // [(F_raw + T)][N]
    long long N;
    int P_raw;
    int ncols;            // used columns of a record
    int R4;
    int src_kind[24];     // 0: X row, 1: plane, 2: zeros, 3: NaNs
    int src_idx[24];
    float* rec;           // [N][R4]
    long long rec_base;   // first record to write
};

// column c of sample i: kind 0 = row of X, 1 = forcing / target plane, 2 = zero column, 3 = NaN column (the padding
// inputs / forcings and the always-masked padding targets of the generic variants)
__device__ __forceinline__ float pack_load(const PackArgs& a, int c, long long i)
{
    const int k = a.src_kind[c];
    if (k >= 2) return k == 2 ? 0.f : __int_as_float(0x7fc00000);
    return k == 0 ? a.X[i * a.P_raw + a.src_idx[c]] : a.planes[(long long)a.src_idx[c] * a.N + i];
}

__global__ void __launch_bounds__(256) k_pack(const PackArgs a)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    float* r = a.rec + (a.rec_base + i) * a.R4;
    for (int c = 0; c < a.R4; c++) {
        float v = 0.f;
        if (c < a.ncols) v = pack_load(a, c, i);
        r[c] = v;
    }
}

// int64 1-based -> int32 0-based index conversion (Julia permutations)
__global__ void __launch_bounds__(256) k_idx_convert(const long long* in, int* out, long long n, long long nmax, int* err)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long v = in[i] - 1;
    if (v < 0 || v >= nmax) { *err = 1; v = 0; }
    out[i] = (int)v;
}

// Epoch staging (collect_dim_data, src/training/epoch.jl:1-11, once per permutation instead of once per batch):
// out[i] = rec[idx[i]], i.e. the records in batch order, so that the persistent kernel streams every batch as one
// contiguous range (one bulk copy per CTA and step) instead of gathering 16-byte records at random.
__global__ void __launch_bounds__(256) k_stage_records(const float4* rec, const int* idx, float4* out, long long n, int r44)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long src = idx[i];
    for (int q = 0; q < r44; q++) out[i * r44 + q] = __ldg(rec + src * r44 + q);
}

// per-step loss values from the reduced sums of the persistent kernel (loss_fn.jl:58-81)
__global__ void k_losses_from_stats(const float* stats, const float* bscal, long long first_step, int nb, int nsteps,
                                    int T, int agg_mean, const int* loss_kind_dev, float* loss_out)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsteps) return;
    const float* bs = bscal + (size_t)((first_step + s) % nb) * BS_STRIDE;
    float L = 0.f, ntot = 0.f;
    for (int t = 0; t < T; t++) {
        float n = bs[BS_N + t], ss = bs[BS_SS + t], acc = stats[(size_t)s * MAXT + t];
        ntot += n;
        int lk = loss_kind_dev[t];
        L += (lk == LOSS_NSELOSS) ? acc / ss : (lk == LOSS_RMSE ? sqrtf(acc / n) : acc / n);
    }
    if (agg_mean) L /= (float)T;
    loss_out[s] = ntot == 0.f ? __int_as_float(0x7fc00000) : L;
}

}  // namespace eh
