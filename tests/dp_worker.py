"""Worker for the data-parallel tests: `torchrun --nproc-per-node W tests/dp_worker.py [gloo|nccl]`.

gloo (CPU): every rank evaluates the ORACLE on its shard of each global batch, the per-rank (n, loss,
grad) are all-reduced and combined (easyhybrid_b200.dp.combine_mse) and must equal the oracle on the
union batch -- the contract the device-side exchange implements.
nccl (GPU): every rank trains its shard through the CUDA library in data-parallel mode; rank 0 checks the
per-step losses and the trained parameters against the oracle trained on the union batches, and that
all ranks hold bit-identical parameters."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import make_synth, rbq10_model  # noqa: E402


def main():
    backend = sys.argv[1] if len(sys.argv) > 1 else "gloo"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    if os.environ.get("EH_DP_DEBUG"):
        os.environ["EH_EPOCH_DEBUG"] = f"gpurun_out/dbg_dp_{rank}.bin"
    import easyhybrid_b200 as eh
    from easyhybrid_b200.dp import combine_mse, global_batch_indices
    from oracle import oracle as orc

    model = rbq10_model(eh)
    n_local, B, steps = (1 << 20, 65536, 24) if os.environ.get("EH_DP_DEBUG") else (6000, 1000, 11)
    shards = [eh.prepare_data(model, make_synth(n_local, seed=100 + r)) for r in range(world)]
    perms = [np.random.default_rng(200 + r).permutation(n_local) for r in range(world)]
    flat0 = model.initialparameters(np.random.default_rng(5))
    # the union dataset, for the single-process oracle
    Xall = np.concatenate([s[0][0] for s in shards])
    fall = {"ta": np.concatenate([s[0][1]["ta"] for s in shards])}
    yall = {"reco": np.concatenate([s[1]["reco"] for s in shards])}
    o = orc.Oracle(model, opt=eh.Adam(0.01))

    if backend == "gloo":
        xf, y = shards[rank]
        for k in range(3):
            idx = perms[rank][k * B:(k + 1) * B][: B - 100 * rank]      # ragged shards
            L, g = o.loss_grad(flat0, xf, y, idx, precision=64)
            t = torch.tensor(np.concatenate([[len(idx) * L, len(idx)], len(idx) * g]))
            dist.all_reduce(t)                                          # sum of n_r L_r, n_r, n_r g_r
            Lg, gg = t[0].item() / t[1].item(), t[2:].numpy() / t[1].item()
            gi = np.concatenate([r * n_local + perms[r][k * B:(k + 1) * B][: B - 100 * r] for r in range(world)])
            Lw, gw = o.loss_grad(flat0, (Xall, fall), yall, gi, precision=64)
            assert abs(Lg - Lw) <= 1e-12 * abs(Lw), (Lg, Lw)
            assert np.abs(gg - gw).max() <= 1e-11 * np.abs(gw).max()
            Lc, gc = combine_mse([len(idx)], [L], [g])
            assert Lc == L
        if rank == 0:
            print("DP_GLOO_OK")
    else:
        xf, y = shards[rank]
        sess = eh.FusedSession(model, opt=eh.Adam(0.01), device=local)
        sess.upload(0, xf, y)
        sess.set_params(flat0)
        sess.comm_init(rank, world, dist)
        sess.set_perm(perms[rank])
        dist.barrier()
        losses = sess.run_steps(B, 0, steps)
        ps = sess.get_params()
        allps = [None] * world
        dist.all_gather_object(allps, ps.tobytes())
        if rank == 0 and os.environ.get("EH_DP_DEBUG"):
            print("DP_NCCL_OK (debug run, no oracle check)", losses[-3:])
        elif rank == 0:
            assert all(b == allps[0] for b in allps), "replicas diverged"
            ref = flat0.copy()
            want = []
            nb = n_local // B
            for s in range(steps):
                gi = global_batch_indices(perms, [n_local] * world, B, s % nb)
                want.append(o.train_steps(ref, (Xall, fall), yall, gi, world * B)[0])
            np.testing.assert_allclose(losses, np.array(want), rtol=2e-4)
            assert abs(float(ps[-1]) - float(ref[-1])) <= 1e-4, (ps[-1], ref[-1])
            print("DP_NCCL_OK", losses[:3], want[:3])
        if not os.environ.get("EH_DP_DEBUG"):
            # the same steps streamed as page-locked HOST batches (collect_dim_data |> gdev per step): packed in place
            # over PCIe, trained in grouped persistent launches, exchange fused as above
            sess.set_params(flat0)
            sess.set_opt_state(None, None, 0)
            dist.barrier()
            nbl = n_local // B
            hl = sess.pinned(np.zeros(steps, dtype=np.float32))
            keep = []
            for s in range(steps):
                idx = perms[rank][(s % nbl) * B:(s % nbl + 1) * B]
                hb = sess.host_batch(sess.pinned(xf[0][idx]), [sess.pinned(xf[1]["ta"][idx])], [sess.pinned(y["reco"][idx])])
                keep.append(hb)
                sess.step_host_async(hb, hl, s)
            sess.sync()
            np.testing.assert_allclose(np.asarray(hl), losses, rtol=2e-6)
            ps_h = sess.get_params()
            np.testing.assert_allclose(ps_h, ps, rtol=0, atol=2e-6)
            allh = [None] * world
            dist.all_gather_object(allh, ps_h.tobytes())
            if rank == 0:
                assert all(b == allh[0] for b in allh), "replicas diverged (host batches)"
                print("DP_NCCL_HOST_OK")
        sess.close()
        if not os.environ.get("EH_DP_DEBUG"):
            # ---- the same model written as a traced callable and compiled at run time (NVRTC): the fused exchange lives in the
            # run-time compiled persistent kernel as well; same trajectory as the built-in form ----
            def traced_rbq10(*, ta, rb, Q10, tref=15.0):
                return {"reco": rb * Q10 ** (0.1 * (ta - tref)) + 0.0 * ta}
            modelj = eh.constructHybridModel(["sw_pot", "dsw_pot"], ["ta"], ["reco"], traced_rbq10, dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0)),
                                             ["rb"], ["Q10"], hidden_layers=[16, 16], activation="tanh", scale_nn_outputs=True)
            xf, y = shards[rank]
            sj = eh.FusedSession(modelj, opt=eh.Adam(0.01), device=local, jit=True)
            assert sj.kernel_variant().startswith("nvrtc/"), sj.kernel_variant()
            sj.upload(0, xf, y)
            sj.set_params(flat0)
            sj.comm_init(rank, world, dist)
            sj.set_perm(perms[rank])
            dist.barrier()
            lj = sj.run_steps(B, 0, steps)
            pj = sj.get_params()
            allj = [None] * world
            dist.all_gather_object(allj, pj.tobytes())
            np.testing.assert_allclose(lj, losses, rtol=2e-5)
            if rank == 0:
                assert all(b == allj[0] for b in allj), "replicas diverged (run-time compiled kernels)"
                print("DP_NCCL_JIT_OK", lj[:3])
            sj.close()
        # ---- second scenario: NaN targets + input BatchNorm + nseLoss: per-batch statistics of the GLOBAL batch ----
        if not os.environ.get("EH_DP_DEBUG"):
            model2 = rbq10_model(eh, bn=True)
            shards2 = [eh.prepare_data(model2, make_synth(n_local, seed=300 + r, nan_frac=0.04), drop_missing_rows=False) for r in range(world)]
            X2 = np.concatenate([s[0][0] for s in shards2])
            f2 = {"ta": np.concatenate([s[0][1]["ta"] for s in shards2])}
            y2 = {"reco": np.concatenate([s[1]["reco"] for s in shards2])}
            o2 = orc.Oracle(model2, training_loss="nseLoss", opt=eh.Adam(0.01))
            xf, y = shards2[rank]
            sess = eh.FusedSession(model2, training_loss="nseLoss", opt=eh.Adam(0.01), device=local)
            sess.upload(0, xf, y)
            sess.set_params(flat0)
            sess.comm_init(rank, world, dist)
            sess.set_perm(perms[rank])
            sess.dp_exchange_batch_stats(B, dist)
            dist.barrier()
            losses = sess.run_steps(B, 0, 6)
            ps = sess.get_params()
            allps = [None] * world
            dist.all_gather_object(allps, ps.tobytes())
            if rank == 0:
                assert all(b == allps[0] for b in allps), "replicas diverged (scenario 2)"
                ref = flat0.copy()
                want = []
                for s in range(6):
                    gi = global_batch_indices(perms, [n_local] * world, B, s % (n_local // B))
                    want.append(o2.train_steps(ref, (X2, f2), y2, gi, world * B)[0])
                np.testing.assert_allclose(losses, np.array(want), rtol=3e-4)
                assert abs(float(ps[-1]) - float(ref[-1])) <= 2e-4, (ps[-1], ref[-1])
                print("DP_NCCL_STATS_OK", losses[:3], want[:3])
            sess.close()
            # ---- third scenario: wide chain (bf16 tcgen05 path), two targets, PerTarget(nseLoss, mse): the 0.2 M-entry
            # gradient is all-reduced over NVLink peer memory, the batch statistics are global ----
            def expo2(n, seed):
                rg = np.random.default_rng(seed)
                T = rg.random(n) * 40 - 10
                SM = rg.random(n) * 0.8 + 0.1
                resp = 1.1 * np.exp(-8.0 * (SM - 0.6) ** 2) * np.exp(0.07 * T)
                return {k: v.astype(np.float32) for k, v in dict(
                    T=T, SM=SM, Resp_obs=resp + rg.standard_normal(n) * 0.05 * resp.mean(),
                    Resp_obs2=2.0 * resp + rg.standard_normal(n) * 0.05 * resp.mean()).items()}
            model3 = eh.constructHybridModel({"Resp0": ["SM"]}, ["T"], ["Resp_obs", "Resp_obs2"], eh.Expo_resp_model2,
                                             dict(k=(0.01, 0.0, 0.2), Resp0=(2.0, 0.0, 8.0)), ["k"],
                                             hidden_layers=[256, 256], activation="tanh", scale_nn_outputs=False)
            nl3, B3, st3 = 4096, 1024, 8
            shards3 = [eh.prepare_data(model3, expo2(nl3, 400 + r)) for r in range(world)]
            perms3 = [np.random.default_rng(500 + r).permutation(nl3) for r in range(world)]
            X3 = np.concatenate([s[0][0] for s in shards3])
            f3 = {"T": np.concatenate([s[0][1]["T"] for s in shards3])}
            y3 = {t: np.concatenate([s[1][t] for s in shards3]) for t in model3.targets}
            loss3 = eh.PerTarget("nseLoss", "mse")
            o3 = orc.Oracle(model3, training_loss=loss3, opt=eh.Adam(0.001))
            flat3 = model3.initialparameters(np.random.default_rng(6))
            xf, y = shards3[rank]
            sess = eh.FusedSession(model3, training_loss=loss3, opt=eh.Adam(0.001), device=local)
            sess.upload(0, xf, y)
            sess.set_params(flat3)
            sess.comm_init(rank, world, dist)
            sess.set_perm(perms3[rank])
            sess.dp_exchange_batch_stats(B3, dist)
            dist.barrier()
            losses = sess.run_steps(B3, 0, st3)
            ps = sess.get_params()
            allps = [None] * world
            dist.all_gather_object(allps, ps.tobytes())
            if rank == 0:
                assert all(b == allps[0] for b in allps), "replicas diverged (wide path)"
                ref = flat3.copy()
                want = []
                for s in range(st3):
                    gi = global_batch_indices(perms3, [nl3] * world, B3, s % (nl3 // B3))
                    want.append(o3.train_steps(ref, (X3, f3), y3, gi, world * B3)[0])
                np.testing.assert_allclose(losses, np.array(want), rtol=5e-2)   # bf16 tolerance (tests/test_gpu_wide.py)
                assert abs(float(ps[-1]) - float(ref[-1])) <= 2e-2, (ps[-1], ref[-1])
                print("DP_NCCL_WIDE_OK", losses[:3], want[:3])
            sess.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
