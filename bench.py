#!/usr/bin/env python
"""bench.py — headline benchmark of the AppleCiDEr hot path on B200, plus the other BASELINE configs in the same record.

Headline (BASELINE.json configs[1], `value` / `e2e` / `roofline`): full multimodal fusion INFERENCE (photometry transformer +
metadata towers/MoE + 63x63x3 cutout ConvNeXt-T + SpectraNet + late-fusion head), batch 4096 synthetic ZTF-shaped alerts per
GPU, bf16 tcgen05 path.  One "step" = one forward over one batch.  N > 1 (torchrun): alerts are independent, every rank runs
its own batch (weak scaling, no collective on the data path); value = all alerts of all ranks / max-over-ranks device time.

Blocks measured in the SAME run and printed in the same JSON line (bench_blocks.py):
  "train"       configs[3]: fusion training, data parallel, B = 512 per GPU: forward + backward + NCCL all-reduce of the flat
                gradient (bucketed, overlapped with the backward) + fused Adam -- at every N, with the exposed all-reduce time
  "cnn_train"   configs[2]: AstroMiNN (ConvNeXt-T cutout CNN + towers + MoE) training, B = 1024            (N = 1 only)
  "preprocess"  configs[4]: P1-P5 over 1 M synthetic alerts, streamed in HBM-resident chunks               (N = 1 only)
  "eager_b200"  the same-box library bar: the reference's modules (oracle port) as plain PyTorch on this GPU (N = 1 only)

  python bench.py --gpus 1 --steps 10 --warmup 3
  python bench.py --impl reference        # CPU oracle port of the reference path, bounded sample
  python bench.py --blocks none           # headline only
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# ---- algorithmic forward FLOPs (2 x MAC), SURVEY.md §8d ---------------------------------------------------------------
FLOPS_SPECTRA_STAGE1_CONV = 2.0 * 1024 * 64 * (3 + 31 + 251) * 128  # 4.782 GFLOP per spectrum
FLOPS_SPECTRANET = 8.222e9          # L = 4096, src config
FLOPS_CONVNEXT = 0.4350e9 + 1.28e6  # ConvNeXt-T @ 63x63 + split head
FLOPS_TOWERS_MOE = 46.8e3 + 84.1e3 + 158.3e3
FLOPS_FUSION_HEAD = 18.8e3


def photo_flops(lens):
    """Transformer forward FLOPs of a batch from its ACTUAL event counts (varlen packing: no padded work)."""
    l0 = np.asarray(lens, dtype=np.float64)
    l1 = l0 + 1.0  # + CLS
    per = 4.0 * (2 * l1 * 128 * 384 + 4 * l1 * l1 * 128 + 2 * l1 * 128 * 128 + 4 * l1 * 128 * 512) + 2 * l0 * 7 * 128
    return float(per.sum())


def fusion_fwd_flops(lens):
    n = len(lens)
    return photo_flops(lens) + n * (FLOPS_SPECTRANET + FLOPS_CONVNEXT + FLOPS_TOWERS_MOE + FLOPS_FUSION_HEAD)


def astrominn_fwd_flops(n):
    return n * (FLOPS_CONVNEXT + FLOPS_TOWERS_MOE)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


def load_ncu_traffic(batch):
    """dram bytes per launch of the dominant kernel from the committed ncu page (profiles/ncu_traffic.json, written by
    tools/ncu_traffic.py); None when the page was captured at another batch size or is absent."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p)).get("dominant")
    if not d or d.get("dram_read_bytes") is None:
        return None, None
    grid0 = int(d["grid"].strip("()").split(",")[0])
    if grid0 not in (batch * 1024 // 128, batch * 1024 // 256):  # stage-1 conv: B*1024 rows in 128- or 256-row CTAs
        return None, d["file"]
    return d["dram_read_bytes"] + d["dram_write_bytes"], d["file"]


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": (float(np.median(sm)) if sm else None), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(B, seed):
    from applecider_b200 import synth

    x, pad, lens = synth.photometry_batch(B, seed=seed)
    return {
        "x": x, "pad": pad, "lens": lens.numpy(), "tokens": int(lens.sum()) + B,
        "meta": synth.metadata(B, seed=seed), "img": synth.cutouts(B, seed=seed), "spec": synth.spectra(B, seed=seed, L=4096),
    }


def run_reference(args):
    """CPU arm: the oracle port of the reference modules (oracle/models.py, pinned against the real reference by
    tests/golden) on all host cores, fp32, eval/no-grad.  NOTE: every step is a BOUNDED SAMPLE (`--cpu-sample` alerts, default
    32) of the 4096-alert workload -- the reference needs ~35 s per full batch on 16 threads -- normalised to alerts/s."""
    from applecider_b200 import synth
    from oracle import models as om

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = args.cpu_sample
    model = om.AppleCider(om.default_config(), hidden_dim=64, fusion="avg").eval()
    model.load_state_dict(synth.det_state_dict(model, 0))
    inp = make_inputs(sample, 1337)
    call = lambda: model(inp["x"], inp["pad"], inp["meta"], inp["img"], inp["spec"])  # noqa: E731
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 1))):
            call()
        steps = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(steps):
            call()
        dt = time.perf_counter() - t0
    val = sample * steps / dt
    out = {
        "impl": "reference", "metric": "fusion_inference_alerts_per_sec", "value": val, "unit": "alerts/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"fusion_inference_b4096_bf16 (CPU arm: same model/inputs, BOUNDED SAMPLE of {sample} alerts per step)",
                   "batch_per_step": sample},
        "cpu_baseline": {"value": val, "unit": "alerts/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} alerts/step x {steps} steps, torch CPU fp32, {cores} threads (oracle port of the reference modules)"},
        "e2e": {"value": val, "unit": "alerts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample", type=int, default=32, dest="cpu_sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--blocks", default="auto",
                    help="comma list of train,cnn_train,preprocess,eager | all | none | auto (N=1: all, N>1: train)")
    ap.add_argument("--train-batch", type=int, default=512, dest="train_batch")
    ap.add_argument("--cnn-batch", type=int, default=1024, dest="cnn_batch")
    ap.add_argument("--prep-alerts", type=int, default=1_000_000, dest="prep_alerts")
    ap.add_argument("--no-overlap", action="store_true", dest="no_overlap", help="train block, N > 1: one all-reduce after the backward instead of bucketed overlap")
    ap.add_argument("--torch-optim", action="store_true", dest="torch_optim", help="train blocks: torch.optim step instead of the fused kernel")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"], help="train blocks: replay the whole step as a CUDA graph")
    ap.add_argument("--workload", default="infer", choices=["infer", "train", "cnn_train", "preprocess", "eager"],
                    help="infer = headline + blocks; any other value runs that block alone and prints its own line")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import bench_blocks as bb

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = bb.Ctx(args=args, peaks=load_peaks(), dist=dist, world=world, rank=rank, local_rank=local_rank, ClockSampler=ClockSampler)

    if args.workload != "infer":
        out = {"train": bb.train_block, "cnn_train": bb.cnn_train_block, "preprocess": bb.preprocess_block, "eager": bb.eager_block}[args.workload](ctx)
        if rank == 0:
            print(json.dumps(out))
        if dist is not None:
            dist.destroy_process_group()
        return

    import applecider_b200 as ab
    from applecider_b200 import _lib, ops, synth

    W = max(args.warmup, 3)
    K = args.steps
    B = args.batch
    peaks = ctx.peaks

    model = ab.AppleCider(ab.default_config(), hidden_dim=64, fusion="avg", compute_dtype=args.dtype)
    model.load_state_dict(synth.det_state_dict(model, 0), strict=True)  # random-init-like deterministic weights
    model = model.cuda().eval()

    host = make_inputs(B, 1337 + rank)
    pinned = {k: v.pin_memory() for k, v in host.items() if torch.is_tensor(v)}
    dev = {k: v.cuda() for k, v in pinned.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned.values())
    ntok = host["tokens"]  # the collate's packed token count (sum of lengths + B): exact token matrix, nothing read back

    def step_dev():
        return model(dev["x"], dev["pad"], dev["meta"], dev["img"], dev["spec"], total_tokens=ntok)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(W):
            step_dev()
        # ---- device-resident throughput ("value") ----
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        _lib.reset_launch_count()
        ops.profile_start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            step_dev()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count()
        regions = ops.profile_stop()
        clocks = sampler.stop() if rank == 0 else None
        # ---- end to end through the public API with host buffers ("e2e") ----
        # model.predict_batches: the public streaming call -- per step one H2D copy of the pinned host batch (on a side
        # stream, overlapping the previous batch's compute) and one D2H read of that step's logits, all inside the timed region
        host_batch = (pinned["x"], pinned["pad"], pinned["meta"], pinned["img"], pinned["spec"])
        for _ in model.predict_batches([host_batch] * 2):
            pass
        barrier()
        t0 = time.perf_counter()
        n_out = 0
        for lg in model.predict_batches(host_batch for _ in range(K)):
            n_out += lg.shape[0]
        barrier()
        e2e_s = time.perf_counter() - t0
        assert n_out == K * B

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()

    out = None
    if rank == 0:
        value = world * B * K / (ms / 1e3)
        e2e = world * B * K / (e2e_ms / 1e3)
        k_ms = float(np.mean(regions.get("spectra.conv.cin64", [float("nan")])))
        achieved = FLOPS_SPECTRA_STAGE1_CONV * B / (k_ms / 1e3) / 1e12
        flops_batch = fusion_fwd_flops(host["lens"])
        traffic, traffic_src = load_ncu_traffic(B)
        out = {
            "metric": "fusion_inference_alerts_per_sec", "value": value, "unit": "alerts/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {
                "workload": f"fusion_inference_b{B}_{args.dtype}", "batch_per_gpu": B, "spectrum_len": 4096, "cutout": "3x63x63",
                "photometry": "Hyrax (B,257,7) layout, lognormal lengths, varlen-packed on device (token count from the collate, no host sync)",
                "tokens_per_batch": ntok,
                "l2": "inputs (293 MB/batch) and activations (>6 GB/step) exceed the 126 MB L2; no explicit flush",
                "parallelism": f"independent alert shards x{world}, no collective",
                "algorithmic_gflop_per_alert": flops_batch / B / 1e9,
                "fraction_of_fusion_tensor_roofline": (value / world) * (flops_batch / B) / (peaks["tf_sustained"] * 1e12),
            },
            "e2e": {"value": e2e, "unit": "alerts/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": B * 5 * 4},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {
                "kernel": "gemm_tc_kernel<128,2,2> — SpectraNet stage-1 multi-kernel Conv1d(64->3x128, k=3/31/251) implicit GEMM (tcgen05, two 128-row sub-tiles per CTA)",
                "bound": "tensor", "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sustained"],
                "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                "kernel_ms": k_ms, "algorithmic_flops_per_launch": FLOPS_SPECTRA_STAGE1_CONV * B,
                "traffic": traffic, "traffic_unit": "bytes/launch (dram__bytes_read.sum + dram__bytes_write.sum)", "traffic_source": traffic_src,
                "region_ms": {k: float(np.mean(v)) for k, v in regions.items()},
            },
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import models as om  # CPU baseline leg only

            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            cm = om.AppleCider(om.default_config(), hidden_dim=64, fusion="avg").eval()
            cm.load_state_dict(synth.det_state_dict(cm, 0))
            n = args.cpu_sample
            with torch.no_grad():
                a = [host["x"][:n], host["pad"][:n], host["meta"][:n], host["img"][:n], host["spec"][:n]]
                cm(*a)
                t0 = time.perf_counter()
                reps = 2
                for _ in range(reps):
                    cm(*a)
                dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": n * reps / dt, "unit": "alerts/s", "cores": cores, "kind": "port",
                                   "sample": f"BOUNDED SAMPLE: first {n} alerts of the same batch x {reps} reps, torch CPU fp32, {cores} threads (oracle port)"}
            del cm
    # ---- the other BASELINE configs, same run, same record ----
    del model, dev
    torch.cuda.empty_cache()
    wanted = bb.resolve_blocks(args.blocks, world)
    blocks = {}
    for name in wanted:
        fn = {"train": bb.train_block, "cnn_train": bb.cnn_train_block, "preprocess": bb.preprocess_block, "eager": bb.eager_block}[name]
        try:
            res = fn(ctx)
        except Exception as e:  # a failing block must not lose the headline (the error is part of the record)
            if name == "train" and world > 1:
                raise  # collective: every rank must fail together
            res = {"error": f"{type(e).__name__}: {e}"[:400]}
        blocks["eager_b200" if name == "eager" else name] = res
        torch.cuda.empty_cache()
    if rank == 0:
        out.update(blocks)
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
