"""GPU: run_detector of the reference pipeline on the device: OCR_b200_Processer.run_detector -- tiles cut on the
device, batched detector, ftc_peak_decode, ftc_page_maps, ftc_box_hists, ftc_select_boxes -- against goldens produced by the
UNMODIFIED reference run_detector (oracle/make_golden.py::golden_page)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def proc32():
    from findtextcenternet_b200 import synthetic
    from findtextcenternet_b200.process_ocr_b200 import OCR_b200_Processer
    return OCR_b200_Processer(precision="fp32", detector_state_dict=synthetic.detector_state_dict(0))


def _tiles(page, offsets):
    im = page.astype(np.float32)
    return im, [{"input": im[None, y:y + 768, x:x + 768, :], "offsetx": x, "offsety": y} for x, y in offsets]


def test_run_detector_dense_stub_page_equals_reference(proc32, monkeypatch):
    """Everything AFTER the detector, exactly: the stub backend's seeded heatmaps (1 150 candidates with heavy overlaps) are injected
    in place of the engine's forward; peak decode + page maps + histogram scores + greedy selection + separator veto + code maximum
    on the device must return the reference's 374 boxes in the reference's order (scores may differ in the last fp32 bit: tanhf on
    the device vs numpy)."""
    from test_oracle_golden import _dense_page
    gold, page, offsets, heat10, feats = _dense_page()
    heat9 = torch.from_numpy(np.concatenate([heat10[:, :1], heat10[:, 2:]], 1)).cuda()
    feat = torch.from_numpy(feats).cuda()
    state = {"i": 0}

    class StubEngine:
        def forward(self, tiles, want_heat10, nhwc255=False):
            b = tiles.shape[0]
            i = state["i"]
            state["i"] += b
            return heat9[i:i + b].contiguous(), feat[i:i + b].contiguous(), None

    monkeypatch.setattr(proc32.detector.detector, "engine", lambda dev: StubEngine())
    im, ds = _tiles(page, offsets)
    loc, gf, lines, seps = proc32.run_detector(ds, im, tile_batch=3)
    ref = gold["locations"]
    assert proc32.last_candidates == int(gold["n_candidates"])
    assert loc.shape == ref.shape and loc.dtype == np.float32 and gf.shape == (len(ref), 100)
    assert np.array_equal(loc[:, 1:3], ref[:, 1:3])                     # same boxes, same order
    np.testing.assert_allclose(loc, ref, rtol=2e-6, atol=1e-7)
    assert np.allclose(gf.astype(np.float64).sum(1), gold["glyphfeatures_sum"], rtol=0, atol=1e-9)
    assert np.abs(lines[::3, ::3] - gold["lines_all_s3"]).max() < 1e-6 and np.abs(seps[::3, ::3] - gold["seps_all_s3"]).max() < 1e-6


def test_box_hists_bit_identical_to_numpy(proc32):
    """ftc_box_hists (3 x 256-bin histograms + two-means gap in double) == the oracle's numpy imageHist for every candidate of the
    golden page, including the boxes whose loose crop wraps around the page border (Python slice semantics)."""
    from findtextcenternet_b200 import synthetic
    from findtextcenternet_b200.engine import box_hists
    from findtextcenternet_b200.process_ocr_b200 import page_tiles
    from oracle import detector_oracle as DO
    gold = np.load(os.path.join(GOLDEN, "page4_seed0.npz"))
    h, w = (int(v) for v in gold["image_hw"])
    page, _ = page_tiles(synthetic.page_image(int(gold["seed"]), h, w))
    loc = gold["pre_locations"].copy()
    loc[0, 1:5] = (3.0, 2.0, 11.0, 9.0)                   # loose crop starts at -1 / -2: wraps to an empty slice
    loc[1, 1:5] = (page.shape[1] - 2.0, page.shape[0] - 3.0, 9.0, 9.0)
    hists = box_hists(torch.from_numpy(page).cuda(), torch.from_numpy(loc).cuda()).cpu().numpy()
    loose, tight = DO.box_hists(loc.astype(np.float64), page.astype(np.float32))
    assert np.array_equal(hists[0], loose) and np.array_equal(hists[1], tight)


def test_run_detector_real_page_close_to_reference(proc32):
    """The whole device run_detector with the real fp32 detector on the 4-tile golden page against the reference's final boxes
    (reference detector on CPU).  A candidate whose score is within fp32 noise of the 0.4 cut-off, or two near-tied scores, can
    change which boxes survive the greedy pass, so the comparison is by box identity: >= 97 % of the reference's boxes are returned
    (same centre), values equal to 1e-3, and nothing else beyond 3 %."""
    from findtextcenternet_b200 import synthetic
    from findtextcenternet_b200.process_ocr_b200 import page_tiles
    gold = np.load(os.path.join(GOLDEN, "page4_seed0.npz"))
    h, w = (int(v) for v in gold["image_hw"])
    page, offsets = page_tiles(synthetic.page_image(int(gold["seed"]), h, w))
    im, ds = _tiles(page, offsets)
    loc, gf, lines, seps = proc32.run_detector(ds, im)
    ref, ref_gf = gold["locations"], gold["glyphfeatures"]
    key = lambda l: (int(l[1]), int(l[2]))
    got_by, ref_by = {key(l): i for i, l in enumerate(loc)}, {key(l): i for i, l in enumerate(ref)}
    common = [k for k in ref_by if k in got_by]
    print(f"reference {len(ref)} boxes, device {len(loc)}, common {len(common)}, candidates {proc32.last_candidates}")
    assert len(common) >= 0.97 * len(ref) and len(loc) <= 1.03 * len(ref)
    a = np.array([loc[got_by[k]] for k in common]); b = np.array([ref[ref_by[k]] for k in common])
    np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-4)
    assert rel_l2(np.array([gf[got_by[k]] for k in common]), np.array([ref_gf[ref_by[k]] for k in common])) < 1e-3
    assert np.abs(lines - gold["maps7"][1]).max() < 2e-5 and np.abs(seps - gold["maps7"][2]).max() < 2e-5
