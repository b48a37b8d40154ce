"""GPU: the reference's backend ABI (process_ocr_base.py:49-55, implemented by process_ocr_torch.py:43-54) on the B200 backend.

``OCR_b200_Processer.call_detector(np.float32[1,768,768,3] in 0..255) -> (heatmap[1,10,192,192], features[1,100,192,192])`` and
``call_transformer(np.float32[1,400,106]) -> np.int64[400]`` are what ``run_detector`` / ``call_OCR`` of the reference pipeline
call; they are checked against goldens produced by the UNMODIFIED reference modules (oracle/make_golden.py)."""
import io
import contextlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def proc32():
    from findtextcenternet_b200 import synthetic
    from findtextcenternet_b200.models.transformer import ModelDimensions
    from findtextcenternet_b200.process_ocr_b200 import OCR_b200_Processer
    dims = ModelDimensions().__dict__
    return OCR_b200_Processer(precision="fp32", detector_state_dict=synthetic.detector_state_dict(0),
                              transformer_state_dict=synthetic.transformer_state_dict(0, **dims), transformer_config=dims)


def test_call_detector_matches_reference_on_test1_tile(proc32, golden_detector, test1_tile):
    """img/test1.png tile (config #1 parity gate) through call_detector: numpy in, numpy out, reference shapes / dtypes; maps within
    1e-3 of the reference CPU run, the peak channel's finite set EXACTLY equal (the -inf pattern of models/detector.py:293-296)."""
    x = test1_tile.astype(np.float32)[None]
    heat, feat = proc32.call_detector(x)
    assert isinstance(heat, np.ndarray) and heat.dtype == np.float32 and heat.shape == (1, 10, 192, 192)
    assert isinstance(feat, np.ndarray) and feat.dtype == np.float32 and feat.shape == (1, 100, 192, 192)
    ref = golden_detector["test1_heatmap10"]
    assert np.array_equal(np.isfinite(heat[0, 1]), np.isfinite(ref[1]))
    fin = np.isfinite(ref[1])
    assert np.abs(heat[0, 1][fin] - ref[1][fin]).max() < 1e-3 * max(1.0, np.abs(ref[1][fin]).max())
    sel = [0] + list(range(2, 10))
    assert rel_l2(heat[0, sel], ref[sel]) < 1e-3
    assert rel_l2(feat[0, :, ::8, ::8], golden_detector["test1_feat_s8"]) < 1e-3
    yx = golden_detector["test1_feat_at_peaks_yx"]
    assert rel_l2(feat[0][:, yx[:, 0], yx[:, 1]].T, golden_detector["test1_feat_at_peaks"]) < 1e-3


def test_call_detector_feeds_reference_decode(proc32, golden_detector):
    """The maps returned by call_detector, pushed through the reference's per-tile decode (oracle.decode_tile, pinned to
    process_ocr_base.py:498-538), give the boxes the reference run_detector found on the rand0 tile before its greedy selection."""
    from findtextcenternet_b200 import synthetic
    from oracle import detector_oracle as DO
    x = (synthetic.detector_input(1, 0, "rand")[0].permute(1, 2, 0).numpy() * 255.).astype(np.float32)[None]
    heat, feat = proc32.call_detector(x)
    ref10 = golden_detector["rand0_heatmap10"]
    # input quantisation: the golden forward saw rand in [0,1); here (x*255)/255 differs in the last bit only
    loc, gf = DO.decode_tile(heat[0], feat[0])
    rloc, _ = DO.decode_tile(ref10, np.zeros((100, 192, 192), np.float32))
    key = lambda l: (int(l[1]), int(l[2]))
    sure = {key(l) for l in rloc if abs(l[0] - 0.4) > 1e-4}
    assert sure <= {key(l) for l in loc} and len(loc) <= len(rloc) + 3 and len(sure) > 50


def test_call_transformer_matches_reference_default_predictor(proc32, golden_transformer):
    """call_transformer on the code-default 768/12/10/400 model (models/transformer.py:255-264, const.py:9-10): int64[400] code
    points identical to the reference TransformerPredictor's (golden ``default_pred_ids``), printed stop line included."""
    from findtextcenternet_b200 import synthetic
    enc, _, _ = synthetic.transformer_inputs(1, 400, 400, seed=0)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ids = proc32.call_transformer(enc.numpy())
    assert isinstance(ids, np.ndarray) and ids.dtype == np.int64 and ids.shape == (400,)
    ref = golden_transformer["default_pred_ids"]
    assert ref.shape == (1, 400)
    assert np.array_equal(ids, ref[0])
    assert buf.getvalue().strip() == str(golden_transformer["default_pred_log"]).strip()


def test_call_transformer_batch_equals_single_calls(proc32):
    """call_transformer_batch (all chunks of a page in ONE predictor call, ftc_transformer_predict_each) == the reference's
    chunk-by-chunk call_transformer loop (process_ocr_base.py:235): every sequence stops by the rule the reference applies to its
    batch of one, so the code points are identical row by row."""
    from findtextcenternet_b200 import synthetic
    enc, _, _ = synthetic.transformer_inputs(3, 400, 400, seed=1)
    with contextlib.redirect_stdout(io.StringIO()):
        batch = proc32.call_transformer_batch(enc.numpy())
        single = np.stack([proc32.call_transformer(enc[i:i + 1].numpy()) for i in range(3)])
    assert batch.shape == (3, 400) and batch.dtype == np.int64
    assert np.array_equal(batch, single)


def test_predict_each_follows_per_sequence_stop_rules():
    """Sequences with DIFFERENT stop behaviour in one batch (tiny model; the output-head bias of U+3042 boosted so that the
    mask-predict loop exits early, as in the reference golden's 'peaked' / 'medium' variants): forward_each row i == forward on
    sequence i alone, including the number of passes and the stop reason."""
    import findtextcenternet_b200.models.transformer as T
    from findtextcenternet_b200 import arch, synthetic
    from oracle.make_golden import TRANSFORMER_CFGS
    dims, _ = TRANSFORMER_CFGS["tiny"]
    old = T.max_decoderlen
    T.max_decoderlen = 24
    try:
        for boost in (20.0, 10.5, 6.0, 0.0):
            sd = synthetic.transformer_state_dict(0, **dims)
            for i, m in enumerate(arch.MODULO_LIST):
                bias = sd[f"decoder.out_layers.{i}.bias"].clone()
                bias[0x3042 % m] += boost
                sd[f"decoder.out_layers.{i}.bias"] = bias
            model = T.Transformer(**T.ModelDimensions(**dims).__dict__)
            model.load_state_dict(sd)
            pred = T.TransformerPredictor(model.encoder, model.decoder).set_precision("fp32").cuda().eval()
            pred.verbose = False
            enc, _, _ = synthetic.transformer_inputs(5, 24, 24, seed=3)
            enc[1] *= 0.05            # a very different sequence: other confidence profile, other stop pass
            enc = enc.cuda()
            ids = pred.forward_each(enc).cpu().numpy()
            state = pred.last_state.numpy()
            for i in range(enc.shape[0]):
                one = pred(enc[i:i + 1]).cpu().numpy()[0]
                assert np.array_equal(ids[i], one), (boost, i)
                assert state[i, 1] == pred.last_passes and state[i, 2] == pred.last_stop_reason, (boost, i, state[i])
    finally:
        T.max_decoderlen = old
