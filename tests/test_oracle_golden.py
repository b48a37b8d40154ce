"""CPU: the oracle restatement vs golden vectors produced by the unmodified reference
(oracle/make_golden.py).  This is the pin that lets the GPU parity tests trust the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from findtextcenternet_b200 import arch, synthetic
from oracle import detector_oracle as DO
from oracle import transformer_oracle as TO


def test_detector_state_keys_match_reference():
    with open(os.path.join(GOLDEN, "detector_state_keys.json")) as f:
        ref = [(k, tuple(s)) for k, s in json.load(f)]
    ours = [(s.key, tuple(s.shape)) for s in arch.text_detector_specs("xl")]
    assert len(ref) == 2444
    assert ours == ref


def test_transformer_state_keys_match_reference():
    from oracle.make_golden import TRANSFORMER_CFGS
    with open(os.path.join(GOLDEN, "transformer_state_keys.json")) as f:
        ref = json.load(f)
    for name, (dims, _) in TRANSFORMER_CFGS.items():
        ours = [(s.key, list(s.shape)) for s in arch.transformer_specs(**dims)]
        assert ours == [(k, s) for k, s in ref[name]], name


@pytest.mark.slow
@pytest.mark.parametrize("name", ["rand0", "test1"])
def test_detector_oracle_matches_reference(name, golden_detector, test1_tile, detector_sd):
    g = golden_detector
    if name == "rand0":
        x = synthetic.detector_input(1, 0, "rand")
    else:
        x = torch.from_numpy(test1_tile.astype(np.float32)[None] / 255.).permute(0, 3, 1, 2).float()
    h10, feat = DO.detector_forward(detector_sd, x)
    h10, feat = h10.numpy()[0], feat.numpy()[0]
    ref = g[name + "_heatmap10"]
    # peak channel: identical -inf pattern (exact index parity), values to fp32 round-off
    assert np.array_equal(np.isfinite(h10[1]), np.isfinite(ref[1]))
    fin = np.isfinite(ref[1])
    assert rel_l2(h10[1][fin], ref[1][fin]) < 1e-5
    other = [0] + list(range(2, 10))
    assert rel_l2(h10[other], ref[other]) < 1e-5
    assert rel_l2(feat[:, ::8, ::8], g[name + "_feat_s8"]) < 1e-5
    yx = g[name + "_feat_at_peaks_yx"]
    assert rel_l2(feat[:, yx[:, 0], yx[:, 1]].T, g[name + "_feat_at_peaks"]) < 1e-5
    # per-tile decode (pre-NMS) must contain every box the reference's run_detector kept
    loc, gf = DO.decode_tile(h10, feat)
    ref_loc = g[name + "_locations"]
    keys = {(int(l[1]), int(l[2])): i for i, l in enumerate(loc)}
    for r, rf in zip(ref_loc, g[name + "_glyphfeatures"]):
        i = keys[(int(r[1]), int(r[2]))]
        np.testing.assert_allclose(loc[i][:5], r[:5], rtol=1e-4)
        np.testing.assert_allclose(gf[i], rf, rtol=1e-4, atol=1e-5)


@pytest.mark.slow
def test_simple_decoder_and_fmask_match_reference(golden_detector, detector_sd):
    g = golden_detector
    gen = torch.Generator().manual_seed(7)
    label = torch.rand(1, 5, 192, 192, generator=gen)
    fmask = DO.get_fmask(label)
    assert np.array_equal(torch.nonzero(fmask)[:, 0].numpy(), g["rand0_fmask_idx"])
    with torch.no_grad():
        _, dec = DO.text_detector_forward(detector_sd, synthetic.detector_input(1, 0, "rand"), fmask)
    for i, d in enumerate(dec):
        assert rel_l2(d.numpy()[::16], g[f"rand0_decoder{i}_s16"]) < 1e-5


def test_crt_matches_reference(golden_transformer):
    g = golden_transformer
    b = g["crt_in"]
    assert np.array_equal(TO.calc_predid_np(b[0], b[1], b[2]), g["crt_out"])
    # exhaustive identity on a range of code points
    x = np.arange(0, 0x40000, 7, dtype=np.int64)
    assert np.array_equal(TO.calc_predid_np(x % 1091, x % 1093, x % 1097), x)


@pytest.mark.parametrize("name", ["tiny", "cfg4", "default"])
def test_transformer_oracle_matches_reference(name, golden_transformer):
    if name == "default":
        pytest.importorskip("torch")
    from oracle.make_golden import TRANSFORMER_CFGS
    g = golden_transformer
    dims, batch = TRANSFORMER_CFGS[name]
    full = dict(enc_input_dim=arch.ENCODER_DIM, embed_dim=768, head_num=12, enc_block_num=10, dec_block_num=10,
                max_enc_seq_len=400, max_dec_seq_len=400)
    full.update(dims)
    sd = synthetic.transformer_state_dict(0, **dims)
    enc, dec, _ = synthetic.transformer_inputs(batch, full["max_enc_seq_len"], full["max_dec_seq_len"], 0)
    logits = TO.transformer_forward(sd, full["head_num"], enc, dec)
    ld = full["max_dec_seq_len"]
    for i, lg in enumerate(logits):
        lg = lg.numpy()
        assert rel_l2(lg[:, ::max(1, ld // 8)], g[f"{name}_logits{i}_s"]) < 2e-5
        assert rel_l2(torch.logsumexp(torch.from_numpy(lg), -1).numpy(), g[f"{name}_lse{i}"]) < 2e-5
    if name != "default":
        ids = TO.predictor_forward(sd, full["head_num"], enc, max_decoderlen=ld)
        assert np.array_equal(ids.numpy(), g[f"{name}_pred_ids"])


@pytest.mark.parametrize("variant,boost", [("peaked", 20.0), ("medium", 10.5)])
def test_predictor_early_exit_variants(variant, boost, golden_transformer):
    from oracle.make_golden import TRANSFORMER_CFGS
    dims, batch = TRANSFORMER_CFGS["tiny"]
    sd = synthetic.transformer_state_dict(0, **dims)
    for i, m in enumerate(arch.MODULO_LIST):
        b = sd[f"decoder.out_layers.{i}.bias"].clone()
        b[0x3042 % m] += boost
        sd[f"decoder.out_layers.{i}.bias"] = b
    enc, _, _ = synthetic.transformer_inputs(batch, 24, 24, 0)
    trace = []
    ids = TO.predictor_forward(sd, dims["head_num"], enc, max_decoderlen=24, trace=trace)
    assert np.array_equal(ids.numpy(), golden_transformer[f"tiny_{variant}_pred_ids"])
    if variant == "peaked":
        assert len(trace) == 1   # "[0 early stop]"


def test_page_maps_oracle_matches_reference_run_detector():
    """oracle page_maps == lines_all / seps_all returned by the unmodified reference run_detector (stub backend, 4 tiles)."""
    import os
    from conftest import GOLDEN
    from findtextcenternet_b200 import synthetic
    from oracle import detector_oracle as DO
    gold = np.load(os.path.join(GOLDEN, "page_maps_seed0.npz"))
    offsets = [tuple(int(v) for v in o) for o in gold["offsets"]]
    ph, pw = (int(v) for v in gold["page_hw"])
    heat9 = synthetic.page_maps_inputs(int(gold["seed"]), len(offsets)).numpy()
    maps = DO.page_maps(heat9, offsets, pw, ph)
    assert maps.shape == (7, ph // 4, pw // 4)
    assert np.abs(maps[1][::3, ::3] - gold["lines_all_s3"]).max() < 1e-6
    assert np.abs(maps[2][::3, ::3] - gold["seps_all_s3"]).max() < 1e-6
    assert maps[1].max() > 0.99 and (maps[1] == 0).sum() == 0          # every page pixel is covered by some tile window


def _dense_page():
    """Inputs of the dense stub page of oracle/make_golden.py::golden_page, regenerated from its seeds."""
    from findtextcenternet_b200.process_ocr_b200 import page_tiles
    gold = np.load(os.path.join(GOLDEN, "page_dense_seed0.npz"))
    h, w = (int(v) for v in gold["image_hw"])
    page, offsets = page_tiles(synthetic.page_image(int(gold["seed"]), h, w))
    assert [tuple(int(v) for v in o) for o in gold["offsets"]] == offsets
    heat10 = synthetic.dense_page_heatmaps(int(gold["seed"]), len(offsets)).numpy()
    feats = torch.randn(len(offsets), 100, 192, 192, generator=torch.Generator().manual_seed(int(gold["feat_seed"]))).numpy()
    return gold, page, offsets, heat10, feats


def test_select_boxes_oracle_matches_reference_run_detector_real_page():
    """oracle.select_boxes (process_ocr_base.py:540-650) on the per-tile peaks of the 4-tile synthetic page == the final boxes of
    the unmodified reference run_detector (reference detector on CPU; golden page4_seed0.npz), bit for bit."""
    gold = np.load(os.path.join(GOLDEN, "page4_seed0.npz"))
    h, w = (int(v) for v in gold["image_hw"])
    from findtextcenternet_b200.process_ocr_b200 import page_tiles
    page, offsets = page_tiles(synthetic.page_image(int(gold["seed"]), h, w))
    assert tuple(page.shape[:2]) == tuple(int(v) for v in gold["page_hw"])
    maps7 = gold["maps7"]
    loc, gf = DO.select_boxes(gold["pre_locations"].astype(np.float64), gold["pre_glyphfeatures"], page.astype(np.float32),
                              maps7[2], maps7[3:7])
    assert loc.shape == gold["locations"].shape and len(loc) > 50 and len(loc) < len(gold["pre_locations"])
    assert np.array_equal(loc, gold["locations"]) and np.array_equal(gf, gold["glyphfeatures"])


def test_select_boxes_oracle_matches_reference_run_detector_dense_page():
    """Same on the dense stub page (1 150 candidates, 374 survivors): IoU, 75 %-intersection, fill-map, histogram, separator and
    code-maximum branches all fire; decode_tile + page_maps + select_boxes chained == reference run_detector output."""
    gold, page, offsets, heat10, feats = _dense_page()
    ph, pw = page.shape[:2]
    pre = [DO.decode_tile(heat10[i], feats[i], x, y, pw, ph) for i, (x, y) in enumerate(offsets)]
    pre_loc, pre_gf = np.concatenate([p[0] for p in pre]), np.concatenate([p[1] for p in pre])
    assert len(pre_loc) == int(gold["n_candidates"])
    heat9 = np.concatenate([heat10[:, :1], heat10[:, 2:]], 1)
    maps7 = DO.page_maps(heat9, offsets, pw, ph)
    assert np.array_equal(maps7[1][::3, ::3], gold["lines_all_s3"]) and np.array_equal(maps7[2][::3, ::3], gold["seps_all_s3"])
    loc, gf = DO.select_boxes(pre_loc, pre_gf, page.astype(np.float32), maps7[2], maps7[3:7])
    assert loc.shape == gold["locations"].shape
    assert np.array_equal(loc, gold["locations"])
    assert np.allclose(gf.astype(np.float64).sum(1), gold["glyphfeatures_sum"], rtol=0, atol=1e-9)
