"""Secondary bench workloads (not the driver's headline): fusion data-parallel training (BASELINE configs[3]),
AstroMiNN training (configs[2]) and the preprocessing sweep (configs[4]).  Same timing rules as bench.py."""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch

FLOPS_FUSION_TRAIN = 27.6e9   # 3 x forward
FLOPS_ASTROMINN_TRAIN = 1.31e9


def _dist():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist, world, rank, local


def run_extra(args, peaks, ClockSampler):
    import applecider_b200 as ab
    from applecider_b200 import _lib, fn, ops, synth
    from applecider_b200.ddp import FlatGradSync, ddp_train_step

    dist, world, rank, local = _dist()
    W, K = max(args.warmup, 3), args.steps

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "preprocess":
        run_preprocess(args, peaks, rank)
        return

    B = args.batch if args.batch != 4096 else (512 if args.workload == "train" else 1024)
    if args.workload == "train":
        model = ab.AppleCider(ab.default_config(), hidden_dim=64, fusion="avg", compute_dtype=args.dtype)
        opt_fn = lambda ps: torch.optim.Adam(ps, lr=1e-3, weight_decay=0.01)  # brew_cider.py:1211
        flops = FLOPS_FUSION_TRAIN
    else:
        cfg = ab.default_config()
        cfg["model"]["AstroMiNN"]["compute_dtype"] = args.dtype
        model = ab.AstroMiNN(cfg)
        opt_fn = None
        flops = FLOPS_ASTROMINN_TRAIN
    model.load_state_dict(synth.det_state_dict(model, 0), strict=True)
    model = model.cuda().train()
    topt = opt_fn([p for p in model.parameters() if p.requires_grad]) if opt_fn else model.this_optimizer
    if getattr(args, "torch_optim", False):
        sync = FlatGradSync(model)
        optimizer = topt
    else:  # fused step: two kernels over flat buffers, bf16 weight shadow refreshed in the same pass
        from applecider_b200.optim import fused_from_torch

        optimizer = fused_from_torch(topt, bf16_shadow=(args.dtype == "bf16"))
        sync = optimizer.grads
    if world > 1 and not getattr(args, "no_overlap", False):
        sync.enable_overlap()  # bucketed all-reduce launched from the autograd hooks, overlapped with the rest of the backward

    x, pad, lens = synth.photometry_batch(B, seed=1337 + rank)
    host = {"x": x, "pad": pad, "meta": synth.metadata(B, seed=1337 + rank), "img": synth.cutouts(B, seed=1337 + rank),
            "spec": synth.spectra(B, seed=1337 + rank, L=4096),
            "tgt": torch.nn.functional.one_hot(synth.labels(B, seed=1337 + rank), 5).float()}
    if args.workload == "cnn_train":
        host = {k: host[k] for k in ("meta", "img", "tgt")}
    pinned = {k: v.pin_memory() for k, v in host.items()}
    dev = {k: v.cuda() for k, v in pinned.items()}
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())

    def fwd_loss(d):
        if args.workload == "train":
            out = model(d["x"], d["pad"], d["meta"], d["img"], d["spec"])
        else:
            out = model((d["meta"], d["img"], d["tgt"]))
        return fn.soft_cross_entropy(out, d["tgt"])

    def step(d):
        return ddp_train_step(sync, lambda: fwd_loss(d), optimizer)

    for _ in range(W):
        step(dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss = step(dev)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        d = {k: v.cuda(non_blocking=True) for k, v in pinned.items()}
        lv = step(d).item()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()
    if rank == 0:
        value = world * B * K / (ms / 1e3)
        out = {
            "metric": "fusion_training_samples_per_sec" if args.workload == "train" else "astrominn_training_samples_per_sec",
            "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"{args.workload}_b{B}_per_gpu_{args.dtype}", "global_batch": B * world,
                       "step": "forward + backward (C-ABI kernels) + NCCL all-reduce(avg) of the flat gradient + "
                               + ("torch optimizer step" if getattr(args, "torch_optim", False) else "fused Adam kernel (acb_adam_step)") + "; dropout on",
                       "optimizer": "Adam(lr 1e-3, wd 0.01)" if args.workload == "train" else "AdamW 11 groups (astrominn.py:151-218)",
                       "grad_elements": sync.numel, "parallelism": f"dp{world}",
                       "fraction_of_tensor_roofline": value / world * flops / (peaks["tf_sustained"] * 1e12)},
            "e2e": {"value": world * B * K / (e2e_ms / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clocks, "final_loss": float(loss),
        }
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def run_preprocess(args, peaks, rank):
    """P1-P5 over synthetic alerts in HBM-resident chunks; alerts/s and achieved GB/s per kernel."""
    from applecider_b200 import preprocess as pp, synth

    if rank != 0:
        return
    n = 100_000
    reps = max(1, args.steps)
    res = {}

    def timeit(f):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    raws = synth.raw_light_curves(20_000, seed=1)
    raws = raws * (n // len(raws))
    raw, off = pp.ragged(raws)
    mean, std = torch.tensor([2.9, 0.9, 1.5, 0.08]).cuda(), torch.tensor([1.1, 0.8, 0.5, 0.04]).cuda()
    ms = timeit(lambda: pp.prep_lightcurves(raw, off, 100.0, mean, std))
    byts = raw.numel() * 4 + n * 257 * 29
    res["P1_lightcurve"] = {"alerts_per_s": n / ms * 1e3, "GBps": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / peaks["hbm_gbs"]}

    specs = synth.raw_spectra(2000, seed=2)
    ns = 20_000
    specs = specs * (ns // len(specs))
    wl, offs = pp.ragged([s[:, 0] for s in specs])
    fx, _ = pp.ragged([s[:, 1] for s in specs])
    grid = pp.wave_grid()
    mx = int((offs[1:] - offs[:-1]).max())
    ms = timeit(lambda: pp.resample_spectra(wl, fx, offs, grid, mx))
    byts = wl.numel() * 16 + ns * 3481 * 4
    res["P3_spectrum_resample"] = {"alerts_per_s": ns / ms * 1e3, "GBps": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / peaks["hbm_gbs"]}

    img = synth.cutouts(4096, seed=3, normalise=False).cuda().repeat(8, 1, 1, 1)
    ms = timeit(lambda: pp.normalize_cutouts(img, "median"))
    byts = img.numel() * 8
    res["P4_cutout_median_norm"] = {"alerts_per_s": img.shape[0] / ms * 1e3, "GBps": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / peaks["hbm_gbs"]}

    ev = torch.randn(20_000_000, 14, device="cuda")
    ms = timeit(lambda: pp.feature_stats(ev))
    res["P5_feature_stats"] = {"rows_per_s": ev.shape[0] / ms * 1e3, "GBps": ev.numel() * 4 / ms / 1e6, "frac_hbm": ev.numel() * 4 / ms / 1e6 / peaks["hbm_gbs"]}

    rng = np.random.default_rng(5)
    nobj = 50_000
    lens = rng.integers(5, 120, size=nobj)
    tot = int(lens.sum())
    offd = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).cuda()
    mjd = torch.from_numpy(np.concatenate([np.sort(rng.uniform(0, 90, size=k)) for k in lens])).cuda()
    mag = torch.from_numpy(rng.normal(19, 0.8, size=tot)).cuda()
    magerr = torch.from_numpy(np.abs(rng.normal(0.08, 0.04, size=tot)) + 0.005).cuda()
    fid = torch.from_numpy(rng.choice([1, 2, 3], size=tot, p=[0.45, 0.45, 0.1]).astype(np.int32)).cuda()
    ms = timeit(lambda: pp.prep_events(mjd, mag, magerr, fid, offd))
    byts = tot * (28 + 17)
    res["P2_event_merge"] = {"alerts_per_s": nobj / ms * 1e3, "GBps": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / peaks["hbm_gbs"]}
    print(json.dumps({"metric": "preprocessing_alerts_per_sec", "unit": "alerts/s", "n_gpus": 1, "data": "synthetic", "kernels": res,
                      "config": {"workload": "preprocess_sweep", "note": "inputs resident in HBM; CUDA-event timing, 3 warm-ups"}}))
