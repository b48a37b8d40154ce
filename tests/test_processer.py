"""train1 input pipeline: dataset/processer.pyx on the device.

CPU (-m "not gpu"):
  * the oracle (oracle/processer_oracle.py) reproduces the reference goldens (tests/golden/processer_golden.npz, written by the
    compiled UNMODIFIED reference): bit-exact hashes for image / line maps / id maps / colour images, 1 ulp on the exp / log maps;
  * when /root/reference is present, the oracle is pinned LIVE against the compiled reference over fresh seeds;
  * the product's host-side parameter drawing equals the oracle's (hence the reference's) draw for draw;
  * the real kernel source (csrc/data_ops.cu compiled for host threads, oracle/emu) equals the goldens.
GPU (-m gpu): GpuProcesser.run through the C-ABI equals the goldens / the oracle with the same tolerances.
"""
import ctypes as C
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import processer_oracle as PO  # noqa: E402
from findtextcenternet_b200.dataset import processer as P  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "processer_golden.npz"), allow_pickle=False)
CASES = [str(c) for c in G["cases"]]
# centre map = product of two expf values, each within 1 ulp of any libm: 2.5e-7 relative.  log-size maps = logf(size / 1024) + 3 with
# |logf| < 8: one ulp of the logarithm is 4.8e-7 ABSOLUTE (the +3 cancels most of the magnitude).
def maps_close(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return (np.array_equal(ref == 0, got == 0) and np.all(np.abs(got[0] - ref[0]) <= 2.5e-7 * np.abs(ref[0]) + 1e-37)
            and np.all(np.abs(got[1:3] - ref[1:3]) <= 5e-7))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def case_sample(k):
    return tuple(G[k + n] for n in ("image", "textline", "sepline", "position", "codelist"))


def drawn_params(k, draw):
    s = case_sample(k)
    rand = PO.ListRand(G[k + "rand"])
    p = draw(rand, s[0].shape[0], s[0].shape[1], s[1].shape[0], s[1].shape[1], s[3])
    assert rand.pos == len(rand.values)          # same number of draws as the reference consumed
    return s, p


def case_params(k, draw=None):
    """the parameters the reference run used (stored: cosf / sinf / logf of the box's C library may differ by an ulp)"""
    i, f = G[k + "p_ints"], G[k + "p_floats"]
    p = dict(rot=G[k + "p_rot"], inv=G[k + "p_inv"], inv2=G[k + "p_inv2"], inv_rect=tuple(int(v) for v in i[:4]), cidx=int(i[4]),
             nearest=bool(i[5]), woffset=f[0], hoffset=f[1], startx0=f[2], starty0=f[3])
    return case_sample(k), p


def check_crop(k, image, maps, idmap, minsize):
    assert sha(np.asarray(image, np.float32).reshape(768, 768)) == str(G[k + "sha_image"]), "768x768 image differs from the reference"
    assert sha(np.asarray(maps, np.float32)[3:]) == str(G[k + "sha_lines"]), "textline / separator maps differ"
    assert sha(np.asarray(idmap, np.int32)) == str(G[k + "sha_idmap"]), "id maps differ"
    assert maps_close(np.asarray(maps, np.float32)[:3], G[k + "maps012"]), "centre / log-size maps differ by more than libm ulps"
    assert np.float32(minsize) == G[k + "minsize"]


@pytest.mark.parametrize("k", CASES)
def test_oracle_matches_reference_golden(k):
    s, p = case_params(k)
    assert p["nearest"] == bool(G[k + "nearest"])
    check_crop(k, *PO.transform_crop(*s, p))


def _alpha():
    s, p = case_params("crop2_")
    return PO.transform_crop(*s, p)[0]


def _bgimg(tag):
    v = G["color_bgimg_seed"]
    seed, h, w = (v[0], v[1], v[2]) if tag == "bg_large" else (v[3], v[4], v[5])
    return (np.random.default_rng(int(seed)).random((int(h), int(w), 3)) * 255).astype(np.uint8)


COLOR = ["mono", "single", "double", "bg_large", "bg_small"]


def color_params(tag, mod):
    rand = PO.ListRand(G[f"color_{tag}_rand"])
    if tag.startswith("bg"):
        bg = _bgimg(tag)
        return mod.draw_background(rand, bg), bg
    return getattr(mod, "draw_" + tag)(rand), None


@pytest.mark.parametrize("tag", COLOR)
def test_oracle_colour_matches_reference_golden(tag):
    cp, bg = color_params(tag, PO)
    assert sha(PO.composite(_alpha(), cp, bg)) == str(G[f"color_{tag}_sha"])


def test_oracle_blank_sample():
    out = PO.process(case_sample("crop1_"), PO.ListRand(G["blank_rand"]))
    assert out[4] is None and not out[0].any() and not out[1].any() and not out[2].any()


@pytest.mark.skipif(not os.path.exists("/root/reference/dataset/processer.pyx"), reason="reference tree not present (GPU box)")
def test_oracle_pinned_live_to_compiled_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_processer"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    import make_golden_processer as MG
    ref = build_ref.load()
    for seed in range(40, 52):
        sample = MG.make_sample(seed, [0, 2, 80, 300][seed % 4], shape=(700, 520))
        PO.LibcRand(seed)
        r = ref.transform_crop(*sample)
        p = PO.draw_crop_params(PO.LibcRand(seed), 700, 520, 350, 260, sample[3])
        o = PO.transform_crop(*sample, p)
        assert np.array_equal(r[0], o[0]) and np.array_equal(r[1][3:], o[1][3:]) and np.array_equal(r[2], o[2]) and float(r[3]) == float(o[3])
        assert maps_close(o[1][:3], r[1][:3])
        a = r[0]
        for name in ("mono", "single", "double"):
            PO.LibcRand(seed)
            rc = getattr(ref, "random_" + name)(a)
            assert np.array_equal(rc, PO.composite(a, getattr(PO, "draw_" + name)(PO.LibcRand(seed))))


def test_host_parameters_equal_the_oracles():
    for k in CASES:
        _, a = drawn_params(k, PO.draw_crop_params)
        _, b = drawn_params(k, P.draw_crop_params)
        _, g = case_params(k)
        for key in a:
            assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), (k, key)      # product == oracle on this machine
            if key in ("rot", "inv", "inv2", "woffset", "hoffset", "startx0", "starty0"):  # == the reference run up to libm ulps
                assert np.allclose(np.asarray(a[key], np.float64), np.asarray(g[key], np.float64), rtol=2e-5, atol=1e-6), (k, key)
            else:
                assert a[key] == g[key], (k, key)
    for tag in COLOR:
        a, _ = color_params(tag, PO)
        b, _ = color_params(tag, P)
        for key in ("fg1", "fg2", "bg"):
            assert np.array_equal(np.asarray(a[key], np.float32), np.asarray(b[key], np.float32)), (tag, key)
        assert tuple(a["rect"]) == tuple(b["rect"]) and a.get("bg_start") == b.get("bg_start")
    assert P.draw_process_params(PO.ListRand(G["blank_rand"]), 10, 10, 5, 5, np.zeros((0, 4), np.float32)) == {"blank": True}
    s, p = case_params("crop2_")
    assert np.float32(P.host_minsize(s[3], p)) == G["crop2_minsize"]


def test_product_input_pipeline_has_no_cpu_route():
    """GpuProcesser is the CUDA path or nothing: a CPU device raises (the oracle is never a fallback)."""
    with pytest.raises(RuntimeError, match="CUDA device only"):
        P.GpuProcesser("cpu")
    import inspect
    src = inspect.getsource(P)
    assert "import oracle" not in src and "from oracle" not in src and "/root/reference" not in src


# ---------------------------------------------------------------------------------------------------------------------
# the kernel source on host threads (oracle/emu)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emu():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "emu"))
    import build_emu
    lib = C.CDLL(build_emu.build())
    lib.ftc_crop_scratch_bytes.restype = C.c_size_t
    lib.ftc_last_error.restype = C.c_char_p
    return lib


def run_host(lib, samples, params, colors=None, salts=None, bgimgs=None):
    """ftc_crop_batch on host pointers (the emulated kernels); same descriptor code as the product (fill_descriptor)."""
    B = len(samples)
    desc = (P.CropSample * B)()
    keep = []
    counts = [s[3].reshape(-1, 4).shape[0] for s in samples]
    total = sum(counts)
    pos = np.ascontiguousarray(np.concatenate([s[3].reshape(-1, 4) for s in samples]), np.float32) if total else np.zeros((1, 4), np.float32)
    code = np.ascontiguousarray(np.concatenate([s[4].reshape(-1, 2) for s in samples]), np.int32) if total else np.zeros((1, 2), np.int32)
    begin = 0
    for b, s in enumerate(samples):
        arrs = {"image": np.ascontiguousarray(s[0]), "textline": np.ascontiguousarray(s[1]), "sepline": np.ascontiguousarray(s[2])}
        shapes = {"image": s[0].shape, "textline": s[1].shape}
        color = colors[b] if colors else None
        salt = salts[b] if salts else None
        if color is not None and color["mode"] == 2:
            arrs["bgimg"] = np.ascontiguousarray(bgimgs[b]); shapes["bgimg"] = bgimgs[b].shape[:2]
        if salt is not None:
            arrs["salt"] = np.ascontiguousarray(salt[1]); shapes["salt"] = salt[1].shape
        keep.append(arrs)
        P.fill_descriptor(desc[b], {k: v.ctypes.data for k, v in arrs.items()}, shapes, begin, counts[b], params[b], color, salt)
        begin += counts[b]
    ch = 1 if not colors else 3
    image = np.empty((B, ch, 768, 768), np.float32)
    maps = np.empty((B, 5, 192, 192), np.float32)
    idmap = np.empty((B, 2, 192, 192), np.int32)
    minsize = np.empty(B, np.float32)
    n = lib.ftc_crop_scratch_bytes(B, total)
    scratch = np.empty(n, np.uint8)
    rc = lib.ftc_crop_batch(C.byref(desc), B, C.c_void_p(pos.ctypes.data), C.c_void_p(code.ctypes.data), total, C.c_void_p(image.ctypes.data), ch,
                            C.c_void_p(maps.ctypes.data), C.c_void_p(idmap.ctypes.data), C.c_void_p(minsize.ctypes.data),
                            C.c_void_p(scratch.ctypes.data), C.c_size_t(n), None)
    assert rc == 0, lib.ftc_last_error()
    return image, maps, idmap, minsize


def test_descriptor_layout_matches_the_header(emu):
    assert emu.ftc_crop_sample_bytes() == C.sizeof(P.CropSample)


def test_emu_crop_kernels_equal_reference_golden(emu):
    ks = ["crop0_", "crop2_", "crop4_"]            # no boxes / bilinear with 120 boxes / nearest with 60 boxes, one batch
    sp = [case_params(k, P.draw_crop_params) for k in ks]
    image, maps, idmap, minsize = run_host(emu, [s for s, _ in sp], [p for _, p in sp])
    for b, k in enumerate(ks):
        check_crop(k, image[b, 0], maps[b], idmap[b], minsize[b])


def test_emu_colour_salt_blank(emu):
    s, p = case_params("crop2_", P.draw_crop_params)
    tags = ["double", "bg_small", "single"]
    cps = [color_params(t, P) for t in tags]
    rng = np.random.default_rng(3)
    salt = P.draw_salt(rng, 40.0, 0.3)
    samples, params = [s, s, s, s], [p, p, p, {"blank": True}]
    colors = [c for c, _ in cps] + [cps[2][0]]
    image, maps, idmap, minsize = run_host(emu, samples, params, colors, [None, None, salt, None], [bg for _, bg in cps] + [None])
    for b in range(2):
        assert sha(image[b]) == str(G[f"color_{tags[b]}_sha"]), tags[b]
    # salt: oracle composite of the salted alpha
    a = _alpha()
    cells = np.repeat(np.repeat(salt[1], salt[0], 0), salt[0], 1)[:768, :768]
    a_s = np.where(cells == 0, np.float32(0), np.where(cells == 2, np.float32(1), a)).astype(np.float32)
    assert np.array_equal(image[2], PO.composite(a_s, color_params("single", PO)[0]))
    # blank sample: zero maps, background colour everywhere
    assert not maps[3].any() and not idmap[3].any() and minsize[3] == 0
    assert np.array_equal(image[3], PO.composite(np.zeros((768, 768), np.float32), color_params("single", PO)[0]))
    check_crop("crop2_", a, maps[0], idmap[0], minsize[0])


def test_emu_edge_cases_empty_boxes_tiny_page_and_negative_coordinates(emu):
    """Edge cases the reference handles implicitly: a batch without a single box (crop origin drawn directly, :367-369), a page much
    smaller than the 768 x 768 crop (everything outside reads as 0), boxes far outside the crop (flag 0), a one-pixel page."""
    rng = np.random.default_rng(5)
    tiny = (rng.random((90, 70)) * 255).astype(np.uint8)
    tl, sp = (rng.random((45, 35)) * 255).astype(np.uint8), (rng.random((45, 35)) * 255).astype(np.uint8)
    none4, none2 = np.zeros((0, 4), np.float32), np.zeros((0, 2), np.int32)
    far = np.array([[5000., 5000., 30., 30.], [35., 45., 20., 25.]], np.float32)
    far_code = np.array([[0x3042, 1], [0x3044, 3]], np.int32)
    one = np.array([[200]], np.uint8)
    samples = [(tiny, tl, sp, none4, none2), (tiny, tl, sp, far, far_code), (one, one, one, none4, none2)]
    params = []
    for i, smp in enumerate(samples):
        params.append(PO.draw_crop_params(PO.LibcRand(100 + i), smp[0].shape[0], smp[0].shape[1], smp[1].shape[0], smp[1].shape[1], smp[3]))
    params[1]["cidx"] = 1                        # anchor on the in-page box so that the other one lies far outside the crop
    image, maps, idmap, minsize = run_host(emu, samples, params)
    for b, (smp, p) in enumerate(zip(samples, params)):
        o = PO.transform_crop(*smp, p)
        assert np.array_equal(image[b, 0], o[0]) and np.array_equal(maps[b, 3:], o[1][3:]) and np.array_equal(idmap[b], o[2])
        assert maps_close(maps[b, :3], o[1][:3]) and np.float32(minsize[b]) == o[3]
    assert not idmap[0].any() and not maps[0, :3].any() and minsize[0] == 0
    assert set(np.unique(idmap[1, 0]).tolist()) <= {0, 0x3044}
    # a batch in which NO sample has a box (total_boxes = 0: the label kernel is not launched)
    image2, maps2, idmap2, _ = run_host(emu, [samples[0], samples[2]], [params[0], params[2]])
    assert np.array_equal(image2[0], image[0]) and np.array_equal(image2[1], image[2]) and not idmap2.any()


# ---------------------------------------------------------------------------------------------------------------------
# random_distortion (dataset/data_detector.py:28-42)
# ---------------------------------------------------------------------------------------------------------------------
DSEEDS = [int(v) for v in G["distort_seeds"]]


def _distort_base():
    return PO.composite(_alpha(), color_params("single", PO)[0])       # the image the goldens were made from (random_single, seed 11)


@pytest.mark.parametrize("seed", DSEEDS)
def test_oracle_distortion_matches_reference_golden(seed):
    base = _distort_base()
    d = PO.draw_distortion(np.random.default_rng(seed), 40.0, base.shape)
    assert sha(PO.random_distortion(base, d)) == str(G[f"distort{seed}_sha"])


def test_host_distortion_decisions_follow_the_reference_order():
    for seed in range(60):          # same branch decisions / amplitudes as the oracle when no noise field is drawn in between
        a = PO.draw_distortion(np.random.default_rng(seed), 33.0, None)
        rng = np.random.default_rng(seed)
        b = P.draw_distortion(rng, 33.0)
        if a["noise_on"]:           # the product draws one integer (the device generator's seed) where the reference draws the field
            assert b["noise_on"] and b["alpha"] == a["alpha"]
            continue
        assert (b["noise_on"], b["mode"], b["sigma"], float(b["unsharp_k"])) == (a["noise_on"], a["mode"], a["sigma"], float(a["unsharp_k"]))
    from scipy.ndimage import _filters
    for sigma in (0.3, 1.0, 1.49, 5.0):
        r, w = P.gauss_taps(sigma)
        full = _filters._gaussian_kernel1d(sigma, 0, r)
        assert r == int(4.0 * sigma + 0.5) and np.array_equal(w, full[r::-1])


def run_distort_host(lib, image, dparams, noise):
    B = image.shape[0]
    desc = (P.DistortSample * B)()
    w = np.ascontiguousarray(np.stack([P.fill_distort(desc[b], dparams[b]) for b in range(B)]))
    lib.ftc_distort_scratch_bytes.restype = C.c_size_t
    n = lib.ftc_distort_scratch_bytes(B)
    scratch = np.empty(n, np.uint8)
    img = np.ascontiguousarray(image, np.float32).copy()
    nz = None if noise is None else np.ascontiguousarray(noise, np.float64)
    rc = lib.ftc_distort_batch(C.c_void_p(img.ctypes.data), B, C.byref(desc), C.c_void_p(w.ctypes.data),
                               None if nz is None else C.c_void_p(nz.ctypes.data), C.c_void_p(scratch.ctypes.data), C.c_size_t(n), None)
    assert rc == 0, lib.ftc_last_error()
    return img


def _distort_cases():
    base = _distort_base()
    ds = [PO.draw_distortion(np.random.default_rng(seed), 40.0, base.shape) for seed in DSEEDS]
    noise = np.stack([d["noise"] if d["noise"] is not None else np.zeros(base.shape) for d in ds])
    return base, ds, noise


def test_emu_distortion_kernels_equal_reference_golden(emu):
    base, ds, noise = _distort_cases()
    out = run_distort_host(emu, np.stack([base] * len(ds)), ds, noise)
    for b, seed in enumerate(DSEEDS):
        assert sha(out[b]) == str(G[f"distort{seed}_sha"]), (seed, ds[b]["noise_on"], ds[b]["mode"])


# ---------------------------------------------------------------------------------------------------------------------
# GPU: the product path through the C-ABI
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_crop_batch_equals_reference_golden():
    proc = P.GpuProcesser("cuda:0")
    sp = [case_params(k, P.draw_crop_params) for k in CASES]
    image, maps, idmap, minsize = proc.run([s for s, _ in sp], [p for _, p in sp])
    image, maps, idmap, minsize = image.cpu().numpy(), maps.cpu().numpy(), idmap.cpu().numpy(), minsize.cpu().numpy()
    for b, k in enumerate(CASES):
        check_crop(k, image[b, 0], maps[b], idmap[b], minsize[b])


@pytest.mark.gpu
def test_gpu_colour_salt_blank_equal_reference_golden():
    proc = P.GpuProcesser("cuda:0")
    s, p = case_params("crop2_", P.draw_crop_params)
    cps = [color_params(t, P) for t in COLOR]
    salt = P.draw_salt(np.random.default_rng(3), 40.0, 0.3)
    n = len(COLOR)
    image, maps, idmap, minsize = proc.run([s] * (n + 2), [p] * (n + 1) + [{"blank": True}], [c for c, _ in cps] + [cps[1][0]] * 2,
                                           [None] * n + [salt, None], [bg for _, bg in cps] + [None, None])
    image = image.cpu().numpy()
    for b, tag in enumerate(COLOR):
        assert sha(image[b]) == str(G[f"color_{tag}_sha"]), tag
    a = _alpha()
    cells = np.repeat(np.repeat(salt[1], salt[0], 0), salt[0], 1)[:768, :768]
    a_s = np.where(cells == 0, np.float32(0), np.where(cells == 2, np.float32(1), a)).astype(np.float32)
    assert np.array_equal(image[n], PO.composite(a_s, cps[1][0]))
    assert np.array_equal(image[n + 1], PO.composite(np.zeros((768, 768), np.float32), cps[1][0]))
    assert not maps[n + 1].any().item() and not idmap[n + 1].any().item()


@pytest.mark.gpu
def test_gpu_processer_call_feeds_the_train_step_layout():
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    proc = P.GpuProcesser("cuda:0", rand=PO.ListRand(np.random.default_rng(0).integers(0, 2**31 - 1, 4000)), rng=np.random.default_rng(1))
    samples = [case_sample(k) for k in CASES]
    image, labelmap, idmap, minsize = proc(samples)
    assert image.shape == (len(CASES), 3, 768, 768) and image.dtype == torch.float32 and image.is_cuda
    assert labelmap.shape == (len(CASES), 5, 192, 192) and idmap.shape == (len(CASES), 2, 192, 192) and idmap.dtype == torch.int64
    assert float(image.min()) >= 0.0 and float(image.max()) <= 1.0 and torch.isfinite(labelmap).all()
    assert float(labelmap[:, 0].max()) <= 1.0


@pytest.mark.gpu
def test_gpu_distortion_equals_reference_golden_and_device_noise_is_standard_normal():
    import torch
    proc = P.GpuProcesser("cuda:0")
    base, ds, noise = _distort_cases()
    img = torch.from_numpy(np.stack([base] * len(ds))).cuda()
    out = proc.distort(img, ds, noise=torch.from_numpy(noise)).cpu().numpy()
    for b, seed in enumerate(DSEEDS):
        assert sha(out[b]) == str(G[f"distort{seed}_sha"]), (seed, ds[b]["noise_on"], ds[b]["mode"])
    # device generator: (out - in) / alpha on unclipped pixels of a mid-grey image is N(0, 1)
    grey = torch.full((2, 3, 768, 768), 0.5, device="cuda")
    dp = [dict(noise_on=True, alpha=0.05, mode=0, sigma=0.0, unsharp_k=0.0, noise_seed=s) for s in (123, 456)]
    z = ((proc.distort(grey.clone(), dp) - grey) / 0.05).double()
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3
    assert abs(float((z ** 3).mean())) < 2e-2 and abs(float((z ** 4).mean()) - 3.0) < 5e-2
    assert float((z[0] - z[1]).abs().mean()) > 0.5                    # different seeds: different fields
    zz = z[0].flatten()
    assert abs(float((zz[:-1] * zz[1:]).mean())) < 5e-3               # neighbouring samples uncorrelated


@pytest.mark.gpu
def test_gpu_full_batch_properties():
    """Batch 64 (the bench's size) through size-independent properties: the batch equals its two halves run separately (bitwise: no
    cross-sample coupling through the atomics / scratch), a blank sample composited with any colour is the background colour, and the
    id map only holds codes of in-crop boxes (or 0)."""
    import torch
    proc = P.GpuProcesser("cuda:0")
    base = [case_params(k) for k in CASES]
    sp = [base[i % len(base)] for i in range(64)]
    samples, params = [s for s, _ in sp], [dict(p) for _, p in sp]
    for i, p in enumerate(params):                 # vary the crop anchor per copy so the 64 samples differ
        if p["cidx"] >= 0:
            p["woffset"] = np.float32(float(p["woffset"]) + 3.0 * (i // len(base)))
        else:
            p["startx0"] = np.float32(float(p["startx0"]) + 3.0 * (i // len(base)))
    params[5] = {"blank": True}
    colors = [color_params(COLOR[i % 3], P)[0] for i in range(64)]
    full = [t.cpu() for t in proc.run(samples, params, colors)]
    lo = [t.cpu() for t in proc.run(samples[:32], params[:32], colors[:32])]
    hi = [t.cpu() for t in proc.run(samples[32:], params[32:], colors[32:])]
    for f, a, b in zip(full, lo, hi):
        assert torch.equal(f, torch.cat([a, b]))
    image, maps, idmap, minsize = full
    bgc = torch.tensor(np.asarray(colors[5]["bg"], np.float32)).view(3, 1, 1)
    assert torch.equal(image[5], bgc.expand(3, 768, 768)) and not maps[5].any() and not idmap[5].any() and minsize[5] == 0
    for b in (0, 2, 40):
        codes = set(np.unique(idmap[b, 0].numpy()).tolist()) - {0}
        assert codes <= set(samples[b][4][:, 0].tolist())
        assert float(maps[b, 0].max()) <= 1.0 and float(maps[b, 0].min()) >= 0.0


@pytest.mark.gpu
def test_gpu_processer_batch_drives_a_train1_step():
    """The batch GpuProcesser produces is what the train step consumes: one eager train1 step (fwd + losses + bwd + optimizer) on it."""
    import torch
    from findtextcenternet_b200 import synthetic, train
    from findtextcenternet_b200.loss_func import CoVWeightingLoss
    from findtextcenternet_b200.models.adamw_schedulefree import AdamWScheduleFree
    from findtextcenternet_b200.models.detector import TextDetectorModel
    proc = P.GpuProcesser("cuda:0", rand=PO.ListRand(np.random.default_rng(3).integers(0, 2**31 - 1, 4000)), rng=np.random.default_rng(4))
    samples = [case_sample("crop2_"), case_sample("crop3_")]
    image, labelmap, idmap, _ = proc(samples)
    model = TextDetectorModel(pre_weights=False)
    model.load_state_dict(synthetic.detector_state_dict(0))
    model.set_precision("bf16")
    model = model.cuda().train()
    opt = AdamWScheduleFree([p for p in model.parameters() if p.requires_grad], lr=1e-4)
    opt.train()
    cov = CoVWeightingLoss(device=image.device, losses=train.TRAIN1_LOSSES)
    fmask = model.get_fmask(labelmap, None)
    loss, raw = train.train1_step(model, opt, cov, image, labelmap, idmap, fmask)
    assert torch.isfinite(loss) and all(torch.isfinite(v).all() for v in raw.values())
    assert int(fmask.sum()) == 1024 * len(samples)
