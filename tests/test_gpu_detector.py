"""GPU parity of the detector engine (through the nn.Module mirror -> C-ABI) against the golden vectors produced by
the unmodified reference on CPU fp32 (tests/golden/detector_xl_seed0.npz, oracle/make_golden.py).

fp32 path (CUDA-core implicit GEMM): heatmap / features within 1e-3 relative (BASELINE.json north_star), peak index
set EXACTLY equal.  bf16 tcgen05 path: activations are STORED in bf16 between ~110 layers (8 mantissa bits, a random
walk of ~0.4 % roundings) -> measured rel-L2 0.5-1.7 % on the random image and 6 % on the nearly-constant white
test1.png tile; bound 1e-1 on the maps and peak-set Jaccard >= 0.9 (exact index parity is only claimed for fp32,
SURVEY.md section 7 "hard parts"; bf16 CUDA-core and bf16 tcgen05 agree with each other to 1e-2)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model(detector_sd):
    from findtextcenternet_b200.models.detector import TextDetectorModel, CenterNetDetector
    m = TextDetectorModel(pre_weights=False)
    m.load_state_dict(detector_sd, strict=True)
    m = m.cuda().eval()
    return m, CenterNetDetector(m.detector).eval()


def _input(name, test1_tile):
    from findtextcenternet_b200 import synthetic
    if name == "rand0":
        return synthetic.detector_input(1, 0, "rand")
    return torch.from_numpy(test1_tile.astype(np.float32)[None] / 255.).permute(0, 3, 1, 2).float()


@pytest.mark.parametrize("name", ["rand0", "test1"])
def test_detector_fp32_matches_reference(name, model, golden_detector, test1_tile):
    m, det = model
    m.detector.set_precision("fp32")
    g = golden_detector
    with torch.no_grad():
        h10, feat = det(_input(name, test1_tile).cuda())
    h10, feat = h10.cpu().numpy()[0], feat.cpu().numpy()[0]
    ref = g[name + "_heatmap10"]
    assert np.array_equal(np.isfinite(h10[1]), np.isfinite(ref[1])), "peak index set differs"
    fin = np.isfinite(ref[1])
    other = [0] + list(range(2, 10))
    assert rel_l2(h10[other], ref[other]) < 1e-3
    assert np.max(np.abs(h10[other] - ref[other])) < 1e-3 * max(1.0, np.abs(ref[other]).max())
    assert rel_l2(h10[1][fin], ref[1][fin]) < 1e-3
    assert rel_l2(feat[:, ::8, ::8], g[name + "_feat_s8"]) < 1e-3
    yx = g[name + "_feat_at_peaks_yx"]
    assert rel_l2(feat[:, yx[:, 0], yx[:, 1]].T, g[name + "_feat_at_peaks"]) < 1e-3
    s = g[name + "_feat_sum"]
    assert abs(feat.astype(np.float64).sum() - s[0]) < 1e-3 * s[1]


def test_detector_fp32_batch_invariance(model):
    """Batch of 2 (rand0 + text image) equals the two singles: no cross-image leakage (SE sums, tiles)."""
    from findtextcenternet_b200 import synthetic
    m, det = model
    m.detector.set_precision("fp32")
    x = torch.cat([synthetic.detector_input(1, 0, "rand"), synthetic.detector_input(1, 1, "text")]).cuda()
    with torch.no_grad():
        hb, fb = m.detector(x)
        h0, f0 = m.detector(x[0:1])
        h1, f1 = m.detector(x[1:2])
    assert rel_l2(hb[0].cpu().numpy(), h0[0].cpu().numpy()) < 1e-5
    assert rel_l2(hb[1].cpu().numpy(), h1[0].cpu().numpy()) < 1e-5
    assert rel_l2(fb[1].cpu().numpy(), f1[0].cpu().numpy()) < 1e-5


@pytest.mark.parametrize("precision", ["bf16_simt", "bf16"])
@pytest.mark.parametrize("name", ["rand0", "test1"])
def test_detector_bf16_close_to_reference(name, precision, model, golden_detector, test1_tile):
    """The benchmarked precision (bf16 storage, fp32 accumulation, tcgen05) against the reference golden.  Bounds calibrated on
    B200 in round 2 (tools/measure_bf16_gate.py, profiles/r02a_bf16_gate.jsonl): rand0 heat 0.61 %, features 0.73 % -> asserted
    at 2e-2; peak sets (3x3 local maxima, +-1 map pixel) Jaccard 0.988 at the 0.4 cut-off (161 peaks) and 0.982 over the 1 970
    peaks above sigmoid 0.27 -> asserted at 0.95.  The near-constant white test1 tile lies outside the noise domain the synthetic
    weights were BN-calibrated on (activations cancel; 3 peaks above the cut-off): 9.4 % measured -> 0.15, no peak statistics."""
    m, det = model
    m.detector.set_precision(precision)
    g = golden_detector
    with torch.no_grad():
        h10, feat = det(_input(name, test1_tile).cuda())
    h10, feat = h10.cpu().numpy()[0], feat.cpu().numpy()[0]
    ref = g[name + "_heatmap10"]
    other = [0] + list(range(2, 10))
    bound = 2e-2 if name == "rand0" else 0.15
    assert rel_l2(h10[other], ref[other]) < bound
    assert rel_l2(feat[:, ::8, ::8], g[name + "_feat_s8"]) < bound
    if name != "rand0":
        return

    def dilate(m):
        p = np.pad(m, 1)
        return np.max([p[dy:dy + m.shape[0], dx:dx + m.shape[1]] for dy in range(3) for dx in range(3)], axis=0)

    for thr, min_peaks in ((-0.405, 100), (-1.0, 200)):       # the cut-off (sigmoid 0.4) and a lower level with ~2 000 peaks
        a, b = np.isfinite(h10[1]) & (h10[1] > thr), np.isfinite(ref[1]) & (ref[1] > thr)
        matched = ((a & dilate(b)).sum() + (b & dilate(a)).sum()) / 2.0
        jaccard = matched / max(a.sum() + b.sum() - matched, 1)
        assert b.sum() >= min_peaks and jaccard >= 0.95, (thr, jaccard, int(a.sum()), int(b.sum()))


def test_tc_and_simt_bf16_agree(model):
    """Same bf16 weights/activations through CUDA cores and through tcgen05: only accumulation order differs."""
    from findtextcenternet_b200 import synthetic
    m, det = model
    x = synthetic.detector_input(2, 3, "rand").cuda()
    outs = {}
    for prec in ("bf16_simt", "bf16"):
        m.detector.set_precision(prec)
        with torch.no_grad():
            h, f = m.detector(x)
        outs[prec] = (h.cpu().numpy(), f.cpu().numpy())
    assert rel_l2(outs["bf16"][0], outs["bf16_simt"][0]) < 1e-2
    assert rel_l2(outs["bf16"][1], outs["bf16_simt"][1]) < 1e-2


def test_state_dict_roundtrip_and_fail_loudly(detector_sd):
    from findtextcenternet_b200.models.detector import TextDetectorModel
    m = TextDetectorModel(pre_weights=False)
    assert list(m.state_dict().keys()) == list(detector_sd.keys())
    with pytest.raises(RuntimeError):
        m.eval().detector(torch.zeros(1, 3, 768, 768))      # CPU tensor: no fallback


def test_text_detector_model_forward_with_fmask(model, golden_detector):
    """TextDetectorModel.forward(x, fmask) = heatmap + SimpleDecoder on the fmask-selected pixels
    (models/detector.py:262-281), fp32, vs the reference golden (every 16th selected row)."""
    from findtextcenternet_b200 import synthetic
    m, _ = model
    m.detector.set_precision("fp32")
    m.decoder.precision = "fp32"
    g = golden_detector
    gen = torch.Generator().manual_seed(7)
    label = torch.rand(1, 5, 192, 192, generator=gen)
    fmask = m.get_fmask(label.cuda(), None)
    assert np.array_equal(torch.nonzero(fmask)[:, 0].cpu().numpy(), g["rand0_fmask_idx"])
    with torch.no_grad():
        heat, dec = m(synthetic.detector_input(1, 0, "rand").cuda(), fmask)
    assert heat.shape == (1, 9, 192, 192)
    for i, d in enumerate(dec):
        assert d.shape == (1024, (1091, 1093, 1097)[i])
        assert rel_l2(d.cpu().numpy()[::16], g[f"rand0_decoder{i}_s16"]) < 1e-3


def test_detect_tiles_end_to_end_matches_reference_decode(detector_sd, golden_detector):
    """The bench's `e2e` path: OCR_b200_Processer.detect_tiles (host NHWC 0..255 tiles -> H2D -> detector -> device peak decode ->
    pinned D2H) against the numpy oracle of process_ocr_base.py:498-538 applied to the reference's own golden heatmap.
    fp32 mode: identical peak set, boxes to 1e-3 (north_star tolerance); features at the peaks vs the golden samples."""
    from findtextcenternet_b200 import synthetic
    from findtextcenternet_b200.process_ocr_b200 import OCR_b200_Processer
    from oracle import detector_oracle as DO
    proc = OCR_b200_Processer(precision="fp32", detector_state_dict=detector_sd)
    x = synthetic.detector_input(1, 0, "rand")                                   # NCHW [0,1)
    tile = (x[0].permute(1, 2, 0) * 255.0).contiguous()[None]                    # backend ABI: NHWC 0..255 float32
    tiles = torch.cat([tile, tile.flip(2)], 0).pin_memory()                      # a second (mirrored) tile: batch > 1
    count, loc, gf = proc.detect_tiles(tiles, [(0, 0), (0, 0)], 768, 768, max_peaks=4096)
    h10 = golden_detector["rand0_heatmap10"]
    n = int(count[0])
    # reference decode needs a feature map only for the gather: use zeros and check features separately
    ref_loc, _ = DO.decode_tile(h10, np.zeros((100, 192, 192), np.float32), 0, 0, 768, 768)
    assert n == len(ref_loc) and n > 10
    got = loc[0, :n].numpy()
    key = lambda l: (int(round(float(l[1]))), int(round(float(l[2]))))
    ref_by = {key(l): i for i, l in enumerate(ref_loc)}
    assert sorted(ref_by) == sorted(key(l) for l in got), "peak set differs from the reference decode"
    perm = np.array([ref_by[key(l)] for l in got])
    np.testing.assert_allclose(got, ref_loc[perm], rtol=1e-3, atol=1e-3)
    # glyph features at the peaks present in the golden sample list
    yx = {(int(y), int(x)): i for i, (y, x) in enumerate(golden_detector["rand0_feat_at_peaks_yx"])}
    hits = 0
    for j, l in enumerate(got):
        k = (int(round(float(l[2]))) // 4, int(round(float(l[1]))) // 4)
        if k in yx:
            assert rel_l2(gf[0, j].numpy(), golden_detector["rand0_feat_at_peaks"][yx[k]]) < 1e-3
            hits += 1
    assert hits > 10
    # the mirrored tile is a different image: it must decode to its own peaks (no cross-tile leakage), same call twice is stable
    count2, loc2, _ = proc.detect_tiles(tiles, [(0, 0), (0, 0)], 768, 768, max_peaks=4096)
    # (the SE squeeze accumulates fc1 shares with fp32 atomics: run-to-run differences of a few ulp are expected)
    assert int(count2[0]) == n and torch.allclose(loc2[0, :n], loc[0, :n], rtol=1e-5, atol=1e-5)
    assert int(count[1]) > 0 and not torch.equal(loc[1, :8], loc[0, :8])


def test_full_size_properties_at_the_benchmarked_configuration(model):
    """BASELINE.json configs[1] at its full size (batch 32, bf16, tcgen05 path), through properties that need no 32-image reference
    run: (1) every image's maps are BIT-identical to its own batch-1 forward (no cross-image coupling, no dependence on how the
    M tiles of a batch are scheduled: the forward is deterministic since round 2); (2) permuting the batch permutes the outputs;
    (3) the peak channel is exactly the 3x3 local-maximum rule of CenterNetDetector.forward (models/detector.py:289-296)
    applied to channel 0, checked against torch's max_pool2d over all 32 x 192 x 192 positions."""
    m, det = model
    m.detector.set_precision("bf16")
    g = torch.Generator().manual_seed(32)
    x = torch.rand(32, 3, 768, 768, generator=g).cuda()
    with torch.no_grad():
        h10, feat = det(x)
        perm = torch.randperm(32, generator=g).cuda()
        h10p, featp = det(x[perm].contiguous())
        for i in (0, 13, 31):
            h1, f1 = det(x[i:i + 1].contiguous())
            assert torch.equal(h1[0], h10[i]) and torch.equal(f1[0], feat[i]), f"image {i}: batch-32 row differs from its batch-1 forward"
    assert torch.equal(h10p, h10[perm]) and torch.equal(featp, feat[perm])
    key = h10[:, 0:1]
    pooled = torch.nn.functional.max_pool2d(torch.nn.functional.pad(key, (1, 1, 1, 1), value=float("-inf")), 3, 1)
    expect = torch.where(key < pooled, torch.full_like(key, float("-inf")), key)
    assert torch.equal(h10[:, 1:2], expect)
    assert torch.isfinite(h10[:, [0] + list(range(2, 10))]).all() and torch.isfinite(feat).all()


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_upload_overlapped_forward_is_bit_identical(model, prec):
    """engine.forward_from_host (the path of detect_tiles for host tiles: the batch crosses PCIe in pieces and
    ftc_detector_forward_part(FTC_PART_EARLY) runs stem + features[1..3] per piece, FTC_PART_REST the rest on the whole batch)
    against the one-piece forward on the same tiles: bit-identical maps for float32 and uint8 host tiles, a ragged split
    (9 images in 4 pieces), more pieces than images, and one piece."""
    m, det = model
    m.detector.set_precision(prec)
    eng = m.detector.engine(torch.device("cuda", torch.cuda.current_device()))
    g = torch.Generator().manual_seed(5)
    b = 9 if prec == "bf16" else 3
    tiles_u8 = torch.randint(0, 256, (b, 768, 768, 3), generator=g, dtype=torch.uint8)
    tiles_f = tiles_u8.float().pin_memory()
    with torch.no_grad():
        ref9, reff, _ = eng.forward(tiles_f.cuda(), False, nhwc255=True)
        ref9, reff = ref9.clone(), reff.clone()
        for tiles, chunks in ((tiles_f, 4), (tiles_u8.pin_memory(), 4), (tiles_f, 16), (tiles_f, 1)):
            h9, ft, _ = eng.forward_from_host(tiles, False, chunks=chunks)
            torch.cuda.synchronize()
            assert torch.equal(h9, ref9) and torch.equal(ft, reff), f"{tiles.dtype} tiles in {chunks} pieces differ from the one-piece forward"
    m.detector.set_precision("bf16")
