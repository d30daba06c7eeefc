"""GPU model parity (B200 box): drop-in modules vs the CPU oracle (live, same weights/inputs) and vs the
committed golden outputs of the REAL reference (tests/golden/*.npz).

Tolerances (SURVEY.md §7 hard part 6): fp32 path |err| <= 1e-4 * max(1, |ref|_inf); bf16 path
|err| <= 5e-3 * max(1, |ref|_inf) (measured 6e-4 .. 2e-3) with 100 % argmax agreement on rows whose top-2 margin exceeds
the same bound."""
import pytest
import torch

from util import assert_close, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"
FP32_TOL = 1e-4
# bf16 bounds, as a fraction of max(1, |ref|_inf), set from what round 2 measured on the B200 (round 1 used 3e-2 everywhere):
#   fusion logits (the benchmarked configuration; encoders feed L2-normalised embeddings): measured 6e-4  -> 5e-3
#   stand-alone encoders (raw logits / 768-d features after 4 transformer layers, 5 conv stages or 18 ConvNeXt blocks of
#   bf16-rounded activations): photometry 6.2e-3, SpectraNet 7.9e-3 (1.05e-2 through the opt-in fused kernels),
#   ConvNeXt features 9.3e-3                                                                       -> 1.5e-2
BF16_TOL_FUSION = 5e-3
BF16_TOL = 1.5e-2
BF16_TOL_PHOTO = BF16_TOL


def _pair(name, cfg_edit=None, dtype="fp32", **kw):
    """(product module on GPU, oracle module on CPU) with identical deterministic weights."""
    import applecider_b200 as ab
    from applecider_b200 import synth
    from oracle import models as om

    cfg = ab.default_config()
    if cfg_edit:
        cfg_edit(cfg)
    ocfg = om.default_config()
    if cfg_edit:
        cfg_edit(ocfg)
    for k in cfg["model"]:
        cfg["model"][k]["compute_dtype"] = dtype
    oracle = getattr(om, name)(ocfg, **kw).eval()
    sd = synth.det_state_dict(oracle, 0)
    oracle.load_state_dict(sd)
    prod = getattr(ab, name)(cfg, **kw)
    missing = prod.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return prod.to(DEV).eval(), oracle


def _argmax_agree(got, ref, tol):
    got, ref = got.float().cpu(), ref.float().cpu()
    top2 = ref.topk(2, -1).values
    margin = top2[:, 0] - top2[:, 1]
    keep = margin > 2 * tol * max(1.0, ref.abs().max().item())
    assert (got.argmax(-1)[keep] == ref.argmax(-1)[keep]).all(), "argmax disagreement on margin-filtered rows"


# ---- photometry ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL_PHOTO)])
def test_photo_matches_golden_and_oracle(golden_dir, dtype, tol):
    g = load_golden(golden_dir, "photo")
    prod, oracle = _pair("HyraxBaselineCLS", dtype=dtype)
    with torch.no_grad():
        got = prod((g["x"].to(DEV), g["pad"].to(DEV), None))
        ref_live = oracle((g["x"], g["pad"], None))
    assert_close(ref_live, g["logits"], 1e-5, "oracle vs golden(real reference)")
    assert_close(got, g["logits"], tol, f"photo {dtype} vs golden")
    assert_close(got, g["logits_slowpath"], tol, f"photo {dtype} vs golden slow path")
    _argmax_agree(got, g["logits"], tol)
    Ls = int((~g["pad"]).sum(1).max())
    with torch.no_grad():
        got_s = prod((g["x"][:, :Ls].contiguous().to(DEV), g["pad"][:, :Ls].contiguous().to(DEV), None))
    assert_close(got_s, g["logits_short"], tol, f"photo {dtype} short layout")


def test_photo_arbitrary_mask_and_embedding_mode():
    from applecider_b200 import synth

    def edit(c):
        c["model"]["HyraxBaselineCLS"]["mode"] = "all"

    prod, oracle = _pair("HyraxBaselineCLS", cfg_edit=edit)
    x, pad, _ = synth.photometry_batch(6, seed=21, L=64)
    gen = torch.Generator().manual_seed(1)
    pad = pad | (torch.rand(pad.shape, generator=gen) < 0.2)  # holes in the middle of sequences
    pad[2] = True  # fully padded: only the CLS token attends to itself
    with torch.no_grad():
        got = prod((x.to(DEV), pad.to(DEV), None))
        ref = oracle((x, pad, None))
    assert got.shape == (6, 128)
    assert_close(got, ref, FP32_TOL, "photo arbitrary mask (embedding mode)")


def test_photo_probabilities_and_legacy_signature():
    import applecider_b200 as ab
    from applecider_b200 import synth

    def edit(c):
        c["model"]["HyraxBaselineCLS"]["use_probabilities"] = True

    prod, oracle = _pair("HyraxBaselineCLS", cfg_edit=edit)
    x, pad, _ = synth.photometry_batch(5, seed=22, L=40)
    with torch.no_grad():
        got = prod((x.to(DEV), pad.to(DEV), None))
        ref = oracle((x, pad, None))
    assert_close(got, ref, FP32_TOL, "photo probabilities")
    assert_close(got.sum(1), torch.ones(5), 1e-5, "rows sum to 1")
    # legacy BaselineCLS(x, pad_mask) = head(norm(z[:,0]))
    legacy = ab.BaselineCLS(128, 8, 4, 5, 0.4)
    sd = {k: v for k, v in synth.det_state_dict(oracle, 0).items() if not k.startswith("fc.")}
    legacy.load_state_dict(sd, strict=True)
    legacy = legacy.to(DEV).eval()
    with torch.no_grad():
        got = legacy(x.to(DEV), pad.to(DEV))
        z = oracle.encode(x, pad)
        ref = torch.nn.functional.linear(z, sd["head.weight"], sd["head.bias"])
    assert_close(got, ref, FP32_TOL, "legacy BaselineCLS")


# ---- spectra ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_spectra_matches_golden(golden_dir, dtype, tol):
    g = load_golden(golden_dir, "spectra")
    prod, oracle = _pair("SpectraNet", dtype=dtype)
    with torch.no_grad():
        got4096 = prod((g["s4096"].to(DEV), None, None))
        got3481 = prod((g["s3481"].to(DEV), None, None))
    assert_close(got4096, g["logits4096"], tol, f"spectra {dtype} L=4096 vs golden")
    assert_close(got3481, g["logits3481"], tol, f"spectra {dtype} L=3481 vs golden")
    if dtype == "fp32":
        blk = prod.all_stages[0][0]
        y, Lo = blk.forward_cl(g["s4096"].to(DEV).view(2, 4096, 1), 2, 4096, torch.float32)
        assert Lo == 1024
        assert_close(y[0, :, 0], g["stage0_b0_c0"], tol, "stage-0 output b0 c0")
        assert_close(y[1, :, 63], g["stage0_b1_c63"], tol, "stage-0 output b1 c63")


def test_spectra_tiny_config_with_committed_weights(golden_dir):
    import applecider_b200 as ab

    g = load_golden(golden_dir, "spectra_tiny")
    cfg = ab.default_config()
    cfg["model"]["SpectraNet"].update(
        channels=[8, 16, 16, 32, 32], kernel_sizes_per_stage=[[3, 9, 33], [3, 7, 17], [3, 5, 9], [3, 5, 7], [3, 5, 7]], flat_dim=96, class_order=4
    )
    m = ab.SpectraNet(cfg)
    m.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith("w::")}, strict=True)
    m = m.to(DEV).eval()
    with torch.no_grad():
        got = m((g["x"].to(DEV), None, None))
    assert_close(got, g["logits"], FP32_TOL, "tiny SpectraNet (ragged L=777) vs real reference")


def test_spectra_redshift_head():
    from applecider_b200 import synth

    def edit(c):
        c["model"]["SpectraNet"]["redshift"] = True

    prod, oracle = _pair("SpectraNet", cfg_edit=edit)
    s = synth.spectra(2, seed=31, L=1024)
    with torch.no_grad():
        got = prod((s.to(DEV), None, None))
        ref = oracle((s, None, None))
    assert got.shape == (2,)
    assert_close(got, ref, FP32_TOL, "redshift regressor")


# ---- image + metadata ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_astrominn_matches_golden(golden_dir, dtype, tol):
    g = load_golden(golden_dir, "astrominn")
    prod, oracle = _pair("AstroMiNN", dtype=dtype)
    with torch.no_grad():
        feat = prod.image_tower.backbone.forward_features(g["image"].to(DEV), prod.compute_dtype)
        got = prod((g["metadata"].to(DEV), g["image"].to(DEV), None))
    assert_close(feat, g["backbone"], tol, f"ConvNeXt-T features {dtype} vs golden")
    assert_close(got, g["logits"], tol, f"AstroMiNN logits {dtype} vs golden")
    _argmax_agree(got, g["logits"], tol)


def test_astrominn_router_indices_exact():
    """top-2 expert selection is integer work: must match the oracle exactly (fp32 path)."""
    from applecider_b200 import ops, synth

    prod, oracle = _pair("AstroMiNN")
    meta, img = synth.metadata(32, seed=41, missing_frac=0.0), synth.cutouts(32, seed=41)
    with torch.no_grad():
        feats = prod.features(meta.to(DEV), img.to(DEV))
        ofeats = oracle.features(meta, img)
        assert_close(feats, ofeats, FP32_TOL, "concatenated tower features")
        r = prod.fusion_router
        gate = ops.gemm(ops.gemm(feats, r[0].weight, r[0].bias, act=ops.ACT_TANH), r[3].weight, r[3].bias, act=ops.ACT_SIGMOID)
        ogate = oracle.fusion_router(ofeats)
    sel = torch.topk(gate.cpu(), 2, -1).indices
    osel = torch.topk(ogate, 2, -1).indices
    top3 = ogate.topk(3, -1).values
    safe = (top3[:, 1] - top3[:, 2]) > 1e-5  # ignore numerically tied gates
    assert torch.equal(sel[safe], osel[safe])


# ---- fusion ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fusion", ["avg", "concat"])
@pytest.mark.parametrize("dtype,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL_FUSION)])
def test_fusion_matches_golden(golden_dir, fusion, dtype, tol):
    g = load_golden(golden_dir, f"fusion_{fusion}")
    import applecider_b200 as ab
    from applecider_b200 import synth
    from oracle import models as om

    oracle = om.AppleCider(om.default_config(), hidden_dim=64, fusion=fusion).eval()
    sd = synth.det_state_dict(oracle, 0)
    oracle.load_state_dict(sd)
    prod = ab.AppleCider(ab.default_config(), hidden_dim=64, fusion=fusion, compute_dtype=dtype)
    prod.load_state_dict(sd, strict=True)
    prod = prod.to(DEV).eval()
    args = [g[k].to(DEV) for k in ("x", "pad", "metadata", "image", "spectra")]
    with torch.no_grad():
        got = prod(*args)
        p, im, s = prod.get_embeddings(*args)
        ref_live = oracle(*[g[k] for k in ("x", "pad", "metadata", "image", "spectra")])
    assert_close(ref_live, g["logits"], 1e-5, "oracle fusion vs golden")
    assert_close(got, g["logits"], tol, f"fusion {fusion} {dtype} logits vs golden")
    if dtype == "fp32":
        assert_close(p, g["p_emb"], tol, "photometry embedding")
        assert_close(im, g["im_emb"], tol, "image+metadata embedding")
        assert_close(s, g["s_emb"], tol, "spectra embedding")


def test_spectra_fused_stage1_path(golden_dir):
    """The opt-in fused conv+LN+GELU kernel for the 64->3x128 stage gives the same logits."""
    from applecider_b200 import spectra as sp

    g = load_golden(golden_dir, "spectra")
    prod, _ = _pair("SpectraNet", dtype="bf16")
    old = sp.FUSE_STAGE1
    sp.FUSE_STAGE1 = True
    try:
        with torch.no_grad():
            got = prod((g["s4096"].to(DEV), None, None))
    finally:
        sp.FUSE_STAGE1 = old
    assert_close(got, g["logits4096"], BF16_TOL, "spectra bf16 with fused stage 1")


def test_predict_batches_streams_host_batches_and_matches_forward():
    """AppleCider.predict_batches (double-buffered H2D on a side stream) returns the logits of forward(), in order."""
    import applecider_b200 as ab
    from applecider_b200 import synth

    model = ab.AppleCider(ab.default_config(), hidden_dim=64, fusion="avg", compute_dtype="bf16")
    model.load_state_dict(synth.det_state_dict(model, 0), strict=True)
    model = model.cuda().eval()
    batches = []
    for i, B in enumerate([5, 5, 3, 5]):  # a shape change in the middle re-allocates the slot
        x, pad, _ = synth.photometry_batch(B, seed=50 + i)
        batches.append(tuple(t.pin_memory() for t in (x, pad, synth.metadata(B, seed=50 + i), synth.cutouts(B, seed=50 + i),
                                                     synth.spectra(B, seed=50 + i, L=4096))))
    got = [o.clone() for o in model.predict_batches(iter(batches))]
    assert len(got) == len(batches)
    with torch.no_grad():
        for b, g in zip(batches, got):
            ref = model(*[t.cuda() for t in b]).float().cpu()
            assert torch.equal(g, ref)
    assert list(model.predict_batches([])) == []


def test_spectra_stage0_persistent_kernel_matches_per_tile_kernel():
    """Stage-0 conv+LN: B=40 takes the persistent rotating-TMEM kernel (>= 148 signal windows), chunks of 8 take the
    one-CTA-per-tile kernel; both accumulate every K block in the same order, so the results are bit-identical."""
    import applecider_b200 as ab
    from applecider_b200 import synth

    cfg = ab.default_config()
    cfg["model"]["SpectraNet"]["compute_dtype"] = "bf16"
    m = ab.SpectraNet(cfg)
    m.load_state_dict(synth.det_state_dict(m, 0), strict=True)
    m = m.cuda().eval()
    x = synth.spectra(40, seed=77, L=4096).cuda()
    blk = m.all_stages[0][0]
    with torch.no_grad():
        big, L8 = blk._conv_ln_fused_bf16(None, 40, 4096, x.view(40, 4096))
        torch.cuda.synchronize()
        parts = [blk._conv_ln_fused_bf16(None, 8, 4096, x[i: i + 8].contiguous().view(8, 4096))[0] for i in range(0, 40, 8)]
    assert torch.isfinite(big.float()).all()
    assert torch.equal(big, torch.cat(parts, 0))
    # and the whole network still matches the fp32 oracle-grade path within the bf16 bound
    cfg32 = ab.default_config()
    cfg32["model"]["SpectraNet"]["compute_dtype"] = "fp32"
    m32 = ab.SpectraNet(cfg32)
    m32.load_state_dict(synth.det_state_dict(m32, 0), strict=True)
    m32 = m32.cuda().eval()
    with torch.no_grad():
        assert_close(m((x.view(40, 1, 4096), None, None)), m32((x.view(40, 1, 4096), None, None)), BF16_TOL, "persistent stage 0 through the network")


def test_spectra_stage0_fused_downsample_matches_unfused():
    """Stage 0 with the 1x1 downsample + MaxPool fused into the persistent kernel (no [B*L,192] activation in HBM) equals the
    conv+LN kernel followed by the pooled GEMM bit for bit: same bf16 operand rows, same K order, and max / +bias / rounding commute."""
    import applecider_b200 as ab
    from applecider_b200 import spectra as sp, synth

    cfg = ab.default_config()
    cfg["model"]["SpectraNet"]["compute_dtype"] = "bf16"
    m = ab.SpectraNet(cfg)
    m.load_state_dict(synth.det_state_dict(m, 0), strict=True)
    m = m.cuda().eval()
    x = synth.spectra(40, seed=78, L=4096).cuda()
    blk = m.all_stages[0][0]
    with torch.no_grad():
        old = sp.FUSE_STAGE0_DOWN
        try:
            sp.FUSE_STAGE0_DOWN = True
            zf, Lf = blk.forward_cl(None, 40, 4096, torch.bfloat16, raw_signal=x.view(40, 4096))
            sp.FUSE_STAGE0_DOWN = False
            zu, Lu = blk.forward_cl(None, 40, 4096, torch.bfloat16, raw_signal=x.view(40, 4096))
        finally:
            sp.FUSE_STAGE0_DOWN = old
    assert Lf == Lu == 1024 and zf.shape == zu.shape == (40, 1024, 64)
    assert torch.isfinite(zf.float()).all()
    assert torch.equal(zf, zu)


def test_spectra_stage1_persistent_fused_conv_ln_close_to_unfused():
    """Stage 1 (64 -> 3 x 128, k = 3/31/251): persistent fused conv + LayerNorm + GELU kernel (B = 40 -> 320 tiles) against the
    conv GEMM followed by the LayerNorm kernel.  Statistics are one-pass from fp32 accumulators in the fused kernel and
    two-pass from bf16-rounded conv outputs in the unfused path: agreement to bf16 rounding (2 ulp of the O(1) outputs)."""
    import applecider_b200 as ab
    from applecider_b200 import spectra as sp, synth

    cfg = ab.default_config()
    cfg["model"]["SpectraNet"]["compute_dtype"] = "bf16"
    m = ab.SpectraNet(cfg)
    m.load_state_dict(synth.det_state_dict(m, 0), strict=True)
    m = m.cuda().eval()
    x = (torch.randn(40, 1024, 64, device="cuda") * 0.7).to(torch.bfloat16)
    blk = m.all_stages[1][0]
    old = sp.FUSE_STAGE1
    try:
        with torch.no_grad():
            sp.FUSE_STAGE1 = True
            zf, _ = blk.forward_cl(x, 40, 1024, torch.bfloat16)
            sp.FUSE_STAGE1 = False
            zu, _ = blk.forward_cl(x, 40, 1024, torch.bfloat16)
    finally:
        sp.FUSE_STAGE1 = old
    assert zf.shape == zu.shape == (40, 256, 128)
    assert_close(zf, zu, 2e-2, "stage-1 fused vs unfused (after downsample + pool)")


def test_fusion_concurrent_encoder_streams_match_single_stream():
    """Opt-in three-stream inference (AppleCider.CONCURRENT_ENCODERS) returns the single-stream logits bit for bit."""
    import applecider_b200 as ab
    from applecider_b200 import synth

    model = ab.AppleCider(ab.default_config(), hidden_dim=64, fusion="avg", compute_dtype="bf16")
    model.load_state_dict(synth.det_state_dict(model, 0), strict=True)
    model = model.cuda().eval()
    B = 6
    x, pad, _ = synth.photometry_batch(B, seed=61)
    args = [t.cuda() for t in (x, pad, synth.metadata(B, seed=61), synth.cutouts(B, seed=61), synth.spectra(B, seed=61, L=4096))]
    with torch.no_grad():
        ref = model(*args).clone()
        try:
            model.CONCURRENT_ENCODERS = True
            for _ in range(3):
                got = model(*args)
            torch.cuda.synchronize()
        finally:
            model.CONCURRENT_ENCODERS = False
    assert torch.equal(got, ref)
