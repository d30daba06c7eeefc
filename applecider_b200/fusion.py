"""Late-fusion AppleCider classifier (reference: _archive/notebooks/brew_cider.py:807-862 head over the
three src/ encoders — DECISION-1 in SURVEY.md §8b) and the fusion collate (models/Time2Vec.py:18-45)."""
from __future__ import annotations

import copy

import os

import torch
import torch.nn as nn

from . import ops
from .astrominn import AstroMiNN
from .photo import HyraxBaselineCLS
from .spectra import SpectraNet


class AppleCider(nn.Module):
    """forward(photometry, photometry_mask, metadata, images, spectra) -> (B, num_classes) logits."""

    def __init__(self, config, hidden_dim=64, fusion="avg", num_classes=5, compute_dtype=None, spectra_variant="src"):
        super().__init__()
        if fusion not in ("avg", "concat"):
            raise NotImplementedError(fusion)
        cfg = copy.deepcopy(config)
        cfg["model"]["HyraxBaselineCLS"]["mode"] = "all"  # encoder returns norm(z[:,0]) (HyraxBaselineCLS.py:37,81)
        cfg["model"]["HyraxBaselineCLS"]["use_probabilities"] = False
        cfg["model"]["HyraxBaselineCLS"]["pretrained_weights_path_"] = False
        cfg["model"]["AstroMiNN"]["use_probabilities"] = False
        if compute_dtype is not None:
            for k in ("HyraxBaselineCLS", "AstroMiNN", "SpectraNet"):
                cfg["model"][k]["compute_dtype"] = compute_dtype
        self.config = cfg
        self.fusion, self.hidden_dim, self.num_classes = fusion, hidden_dim, num_classes
        self.classification = True
        self.photometry_encoder = HyraxBaselineCLS(cfg)
        self.spectra_variant = spectra_variant
        if spectra_variant == "B":  # legacy checkpoints: variant-B encoder -> 256-d embedding (brew_cider.py:585-708,826)
            from .legacy import SpectraClassificationB

            self.spectra_encoder = SpectraClassificationB({"mode": "all", "classes": list(range(num_classes))}, compute_dtype=compute_dtype)
        elif spectra_variant == "src":
            self.spectra_encoder = SpectraNet(cfg)
        else:
            raise ValueError(f"unknown spectra_variant {spectra_variant!r}")
        self.img_metadata_encoder = AstroMiNN(cfg)
        sc = cfg["model"]["SpectraNet"]
        spec_out = 256 if spectra_variant == "B" else (1 if sc["redshift"] else sc["class_order"])
        self.photometry_proj = nn.Linear(cfg["model"]["HyraxBaselineCLS"]["d_model"], hidden_dim)
        self.spectra_proj = nn.Linear(spec_out, hidden_dim)
        self.img_metadata_proj = nn.Linear(5, hidden_dim)
        self.fc = nn.Linear(hidden_dim * 3 if fusion == "concat" else hidden_dim, num_classes)

    # inference: the three encoders are independent and can run on three streams (measured +1.8 % throughput at B=4096, but the
    # co-running small kernels slow the tensor-bound spectra convs by ~5 %, which muddies per-kernel timing) -> opt-in
    CONCURRENT_ENCODERS = int(os.environ.get("ACB_CONCURRENT_ENCODERS", "0"))  # 1 = side streams, 2 = high-priority side streams

    def _encode(self, photometry, photometry_mask, metadata, images, spectra, total_tokens=None):
        if self.CONCURRENT_ENCODERS and not torch.is_grad_enabled() and spectra.is_cuda and self.spectra_variant == "src":
            return self._encode_concurrent(photometry, photometry_mask, metadata, images, spectra, total_tokens)
        # no step of the forward reads anything back from the device (the token packing works on a row capacity, see
        # photo.pack), so the launch order is free; the spectra encoder goes first only because its ~20 long kernels keep the
        # GPU busy while the ~100 short photometry / ConvNeXt / tower launches are queued behind them
        s = self.spectra_encoder(spectra) if self.spectra_variant == "B" else self.spectra_encoder((spectra, None, None))
        if s.dim() == 1:
            s = s[:, None].contiguous()
        p = self.photometry_encoder((photometry, photometry_mask, None), total_tokens=total_tokens)
        im = self.img_metadata_encoder((metadata, images, None))
        return p, im, s

    def _encode_concurrent(self, photometry, photometry_mask, metadata, images, spectra, total_tokens=None):
        """Spectra on the caller's stream (it is 70 % of the work), photometry and image+metadata on two side streams."""
        dev = spectra.device
        main = torch.cuda.current_stream(dev)
        side = getattr(self, "_side_streams", None)
        if side is None:
            pr = -1 if int(self.CONCURRENT_ENCODERS) >= 2 else 0
            side = self._side_streams = (torch.cuda.Stream(dev, priority=pr), torch.cuda.Stream(dev, priority=pr))
        start = torch.cuda.Event()
        start.record(main)
        outs = [None, None]
        done = []
        for k, st in enumerate(side):
            st.wait_event(start)
            with torch.cuda.stream(st):
                if k == 0:
                    outs[0] = self.photometry_encoder((photometry, photometry_mask, None), total_tokens=total_tokens)
                else:
                    outs[1] = self.img_metadata_encoder((metadata, images, None))
                ev = torch.cuda.Event()
                ev.record(st)
                done.append(ev)
        s = self.spectra_encoder((spectra, None, None))
        if s.dim() == 1:
            s = s[:, None].contiguous()
        for ev, o in zip(done, outs):
            main.wait_event(ev)
            o.record_stream(main)
        for t in (photometry, photometry_mask, metadata, images):  # inputs were produced on `main`, read on the side streams
            if t is not None:
                t.record_stream(side[0])
                t.record_stream(side[1])
        return outs[0], outs[1], s

    def _head(self, p, im, s, want_emb):
        B = p.shape[0]
        logits = torch.empty((B, self.num_classes), dtype=torch.float32, device=p.device)
        emb = torch.empty((3, B, self.hidden_dim), dtype=torch.float32, device=p.device) if want_emb else None
        ops.call(
            "acb_fusion_head", p, p.shape[1], im, im.shape[1], s, s.shape[1], self.photometry_proj.weight, self.photometry_proj.bias,
            self.img_metadata_proj.weight, self.img_metadata_proj.bias, self.spectra_proj.weight, self.spectra_proj.bias,
            self.fc.weight, self.fc.bias, self.hidden_dim, int(self.fusion == "concat"), self.num_classes, logits, emb, B,
        )
        return logits, emb

    def get_embeddings(self, photometry, photometry_mask, metadata, images, spectra, total_tokens=None):
        p, im, s = self._encode(photometry, photometry_mask, metadata, images, spectra, total_tokens)
        _, emb = self._head(p, im, s, True)
        return emb[0], emb[1], emb[2]

    def forward(self, photometry, photometry_mask, metadata, images, spectra, total_tokens=None):
        """total_tokens (optional, an extension of the reference signature): the packed token count of the batch from the
        collate -- number of unmasked events + B -- or any upper bound of it; sizes the token matrix exactly.  Without it the
        capacity is B*(L+1) and the kernels skip the unused rows by a device-side count; either way nothing syncs."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .train import fusion_forward_train

            return fusion_forward_train(self, photometry, photometry_mask, metadata, images, spectra, total_tokens)
        p, im, s = self._encode(photometry, photometry_mask, metadata, images, spectra, total_tokens)
        return self._head(p, im, s, False)[0]


    @torch.no_grad()
    def predict_batches(self, host_batches):
        """Streaming inference over HOST batches: yields one (B, num_classes) fp32 CPU logits tensor per batch (a fresh
        tensor each time: the two pinned staging buffers behind it are reused, the yielded copies are the caller's).

        host_batches: iterable of (photometry, photometry_mask, metadata, images, spectra) CPU tensors (pinned memory
        makes the copies asynchronous).  The host->device copy of batch i+1 runs on a side stream while batch i is being
        computed (two device slots); the logits come back with an asynchronous device->host copy and are yielded one
        batch late, after their own event -- nothing blocks the device queue.
        """
        dev = self.fc.weight.device
        if dev.type != "cuda":
            raise RuntimeError("applecider_b200: the model must live on a CUDA device (no CPU fallback)")
        compute = torch.cuda.current_stream(dev)
        copy_stream = getattr(self, "_copy_stream", None)
        if copy_stream is None:
            copy_stream = self._copy_stream = torch.cuda.Stream(dev)
        slots = [None, None]            # device input buffers, reused every other batch
        h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
        slot_free = [None, None]        # compute finished reading the slot
        outs = [None, None]
        out_done = [None, None]

        ntok = [None, None]

        def stage(i, batch):
            k = i & 1
            # packed token count from the HOST mask (the collate's by-product: ~1 MB of bools): sizes the token matrix exactly
            ntok[k] = int(batch[1].numel() - int(batch[1].count_nonzero())) + batch[1].shape[0]
            with torch.cuda.stream(copy_stream):
                if slot_free[k] is not None:
                    copy_stream.wait_event(slot_free[k])
                if slots[k] is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(slots[k], batch)):
                    slots[k] = [torch.empty(h.shape, dtype=h.dtype, device=dev) for h in batch]
                for d, h in zip(slots[k], batch):
                    d.copy_(h, non_blocking=True)
                h2d_done[k].record(copy_stream)

        it = iter(host_batches)
        nxt = next(it, None)
        if nxt is None:
            return
        stage(0, nxt)
        i = 0
        pending = None
        while nxt is not None:
            k = i & 1
            nxt = next(it, None)
            if nxt is not None:
                stage(i + 1, nxt)  # overlaps with the compute of batch i
            compute.wait_event(h2d_done[k])
            d = slots[k]
            logits = self.forward(d[0], d[1], d[2], d[3], d[4], total_tokens=ntok[k])
            slot_free[k] = torch.cuda.Event()
            slot_free[k].record(compute)
            if outs[k] is None or outs[k].shape != logits.shape:
                outs[k] = torch.empty(logits.shape, dtype=torch.float32).pin_memory()
            outs[k].copy_(logits, non_blocking=True)
            out_done[k] = torch.cuda.Event()
            out_done[k].record(compute)
            if pending is not None:
                out_done[pending].synchronize()
                yield outs[pending].clone()
            pending = k
            i += 1
        out_done[pending].synchronize()
        yield outs[pending].clone()


def fusion_collate(batch, mean, std, device="cuda"):
    """Fusion collate contract of models/Time2Vec.py:18-45 with the hard-coded stats path turned into
    arguments: batch of (photo[L,7], metadata, image, spectra[1,Ls], label) ->
    (photo[B,Lmax,7] normalised, mask[B,Lmax] bool True=pad, metadata, images, spectra, labels)."""
    photo, meta, img, spec, labels = zip(*batch)
    lens = [int(s.shape[0]) for s in photo]
    Lmax = max(lens)
    B = len(photo)
    x = torch.zeros((B, Lmax, 7), dtype=torch.float32)
    mask = torch.ones((B, Lmax), dtype=torch.bool)
    for i, s in enumerate(photo):
        x[i, : lens[i]] = torch.as_tensor(s, dtype=torch.float32)
        mask[i, : lens[i]] = False
    mean = torch.as_tensor(mean, dtype=torch.float32)
    std = torch.as_tensor(std, dtype=torch.float32)
    x[..., :4] = (x[..., :4] - mean) / (std + 1e-8)
    to = lambda t: t.to(device, non_blocking=True)  # noqa: E731
    return (to(x), to(mask), to(torch.stack([torch.as_tensor(m, dtype=torch.float32) for m in meta])),
            to(torch.stack([torch.as_tensor(i, dtype=torch.float32) for i in img])),
            to(torch.stack([torch.as_tensor(s, dtype=torch.float32) for s in spec])), to(torch.as_tensor(labels)))


EVENT_COLUMNS = ["dt", "dt_prev", "band_id", "logflux", "logflux_err", "band_ztfg", "band_ztfr", "band_ztfi",
                 "g_r", "g_r_err", "r_i", "r_i_err", "has_g_r", "has_r_i"]  # build_event_features minus obj_id/jd/fid (:680)
MODEL_EVENT_COLUMNS = ["dt", "dt_prev", "logflux", "logflux_err", "band_ztfg", "band_ztfr", "band_ztfi"]


def from_pad_collate(batch, mean, std, event_columns=None, log1p_dt=False, n_meta=24, device="cuda"):
    """MultiModalDataset.pad_collate dict (events[B,T,14], events_mask True=valid, image, metadata[B,46], label) ->
    (photometry[B,T,7] normalised, pad_mask[B,T] True=pad, metadata[B,n_meta], image, label) on the device.

    The column gather, the optional log1p of the time columns, the normalisation and the mask polarity flip run in one
    kernel (acb_collate_events); AstroMiNN's 24 metadata columns are the first 24 of ALERT_META_KEEP
    (preprocess_multimodal.py:615-640, consistent with the tower comments astrominn.py:249-254)."""
    names = list(event_columns) if event_columns is not None else EVENT_COLUMNS
    cols = torch.tensor([names.index(c) for c in MODEL_EVENT_COLUMNS], dtype=torch.int32, device=device)
    ev = batch["events"].to(device, non_blocking=True).float().contiguous()
    valid = batch["events_mask"].to(device, non_blocking=True).contiguous()
    B, T, Fe = ev.shape
    if Fe != len(names):
        raise ValueError(f"events have {Fe} columns, expected {len(names)}")
    x = torch.empty((B, T, 7), dtype=torch.float32, device=device)
    pad = torch.empty((B, T), dtype=torch.bool, device=device)
    m = torch.as_tensor(mean, dtype=torch.float32).to(device).contiguous()
    sd = torch.as_tensor(std, dtype=torch.float32).to(device).contiguous()
    ops.call("acb_collate_events", ev, valid.view(torch.uint8), B, T, Fe, cols, int(log1p_dt), m, sd, x, pad.view(torch.uint8))
    meta = batch["metadata"].to(device, non_blocking=True).float()[:, :n_meta].contiguous()
    return x, pad, meta, batch["image"].to(device, non_blocking=True).float().contiguous(), batch["label"].to(device, non_blocking=True)
