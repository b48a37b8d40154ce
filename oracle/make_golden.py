"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference)
on the synthetic checkpoints.  Runs only in the build container (the GPU box has no
/root/reference); the outputs are committed.  Usage:  python oracle/make_golden.py [detector|page|chunks|transformer|optimizer|radam|loss|all]
"""
import io
import os
import sys
import contextlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
GOLD = os.path.join(ROOT, "tests", "golden")

from findtextcenternet_b200 import arch, synthetic  # noqa: E402


def load_test1_tile():
    """img/test1.png -> the single padded 768x768 tile exactly as call_OCR builds it
    (process_ocr_base.py:58-76).  Returns uint8 [768,768,3]."""
    from PIL import Image
    im0 = np.asarray(Image.open(os.path.join(REF, "img", "test1.png")).convert("RGB"))
    stepx = stepy = int(768 * 0.6)
    padx = max(0, (768 - im0.shape[1]) % stepx, 768 - im0.shape[1])
    pady = max(0, (768 - im0.shape[0]) % stepy, 768 - im0.shape[0])
    im0 = np.pad(im0, [[0, pady], [0, padx], [0, 0]], "constant", constant_values=255)
    assert im0.shape == (768, 768, 3), im0.shape
    return im0


def golden_detector():
    from models.detector import TextDetectorModel, CenterNetDetector
    from process_ocr_base import OCR_Processer

    sd = synthetic.detector_state_dict(0)
    model = TextDetectorModel(pre_weights=False)
    ref_keys = [(k, list(v.shape)) for k, v in model.state_dict().items()]
    missing = model.load_state_dict(sd, strict=True)
    print("load_state_dict:", missing)
    det = CenterNetDetector(model.detector).eval()

    tile = load_test1_tile()
    np.savez_compressed(os.path.join(GOLD, "test1_tile.npz"), tile=tile)
    inputs = {
        "rand0": synthetic.detector_input(1, 0, "rand"),
        "test1": torch.from_numpy(tile.astype(np.float32)[None] / 255.).permute(0, 3, 1, 2).float(),
    }
    out = {}
    with torch.no_grad():
        for name, x in inputs.items():
            h10, feat = det(x)
            h10, feat = h10.numpy(), feat.numpy()
            pk = np.isfinite(h10[0, 1])
            out[name + "_heatmap10"] = h10[0]
            out[name + "_feat_s8"] = feat[0, :, ::8, ::8].copy()
            ys, xs = np.nonzero(pk & (h10[0, 1] > -1.0))
            out[name + "_feat_at_peaks_yx"] = np.stack([ys, xs], 1).astype(np.int32)
            out[name + "_feat_at_peaks"] = feat[0][:, ys, xs].T.copy()
            out[name + "_feat_sum"] = np.array([feat.astype(np.float64).sum(), np.abs(feat).astype(np.float64).sum()])
            print(name, "peaks finite", int(pk.sum()), "kept", len(ys))

    # reference run_detector on the single-tile test1 page and on the rand0 "page"
    class RefProc(OCR_Processer):
        def call_detector(self, image_input):
            images = torch.from_numpy(image_input / 255.).permute(0, 3, 1, 2).float()
            with torch.no_grad():
                h, f = det(images)
            return h.numpy(), f.numpy()

        def call_transformer(self, encoder_input):
            raise NotImplementedError

    proc = RefProc()
    for name, x in inputs.items():
        im = (x[0].permute(1, 2, 0).numpy() * 255.).astype(np.float32)
        ds0 = [{"input": im[None], "offsetx": 0, "offsety": 0}]
        with contextlib.redirect_stdout(io.StringIO()):
            loc, gf, lines, seps = proc.run_detector(ds0, im)
        out[name + "_locations"] = loc
        out[name + "_glyphfeatures"] = gf
        out[name + "_lines_s4"] = lines[::4, ::4].copy()
        out[name + "_seps_s4"] = seps[::4, ::4].copy()
        print(name, "run_detector boxes", loc.shape)

    # train-path pieces: get_fmask + SimpleDecoder (TextDetectorModel.forward), eval mode
    model.eval()
    g = torch.Generator().manual_seed(7)
    label = torch.rand(1, 5, 192, 192, generator=g)
    fmask = model.get_fmask(label, None)
    with torch.no_grad():
        heat, dec = model(inputs["rand0"], fmask)
    out["rand0_fmask_idx"] = torch.nonzero(fmask)[:, 0].numpy().astype(np.int32)
    for i, d in enumerate(dec):
        out[f"rand0_decoder{i}_s16"] = d.numpy()[::16].copy()
    np.savez_compressed(os.path.join(GOLD, "detector_xl_seed0.npz"), **out)
    import json
    with open(os.path.join(GOLD, "detector_state_keys.json"), "w") as f:
        json.dump(ref_keys, f)
    print("detector goldens written")


PAGE4 = dict(seed=0, height=900, width=1000)
DENSE = dict(seed=5, feat_seed=6, height=900, width=1000)


def dense_features(n_tiles):
    g = torch.Generator().manual_seed(DENSE["feat_seed"])
    return torch.randn(n_tiles, 100, 192, 192, generator=g).numpy()


def golden_page():
    """tests/golden/page4_seed0.npz: the UNMODIFIED reference ``run_detector`` (process_ocr_base.py:474-650) over a 4-tile
    synthetic page with the reference detector on CPU -> final boxes / glyph features / page maps, plus the per-tile peak decode
    (oracle.decode_tile applied to the reference's own tile heatmaps) that the device-side ``detect_page`` must reproduce.
    tests/golden/page_dense_seed0.npz: the same reference function over a stub backend whose seeded heatmaps carry ~2 000
    heavily overlapping boxes, so every branch of the greedy selection is exercised."""
    from models.detector import TextDetectorModel, CenterNetDetector
    from process_ocr_base import OCR_Processer
    from findtextcenternet_b200.process_ocr_b200 import page_tiles
    from oracle import detector_oracle as DO

    sd = synthetic.detector_state_dict(0)
    model = TextDetectorModel(pre_weights=False)
    model.load_state_dict(sd, strict=True)
    det = CenterNetDetector(model.detector).eval()
    captured = []

    class RefProc(OCR_Processer):
        def call_detector(self, image_input):
            images = torch.from_numpy(image_input / 255.).permute(0, 3, 1, 2).float()
            with torch.no_grad():
                h, f = det(images)
            captured.append((h.numpy(), f.numpy()))
            return captured[-1]

        def call_transformer(self, encoder_input):
            raise NotImplementedError

    def tiles_of(page):
        im = page.astype(np.float32)
        return im, [{"input": im[None, y:y + 768, x:x + 768, :], "offsetx": x, "offsety": y} for x, y in offsets]

    page, offsets = page_tiles(synthetic.page_image(PAGE4["seed"], PAGE4["height"], PAGE4["width"]))
    im, ds0 = tiles_of(page)
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        loc, gf, lines, seps = RefProc().run_detector(ds0, im)
    pre_loc, pre_gf, counts = [], [], []
    for (h10, feat), (x, y) in zip(captured, offsets):
        l, f = DO.decode_tile(h10[0], feat[0], x, y, page.shape[1], page.shape[0])
        pre_loc.append(l); pre_gf.append(f); counts.append(len(l))
    pre_loc, pre_gf = np.concatenate(pre_loc), np.concatenate(pre_gf)
    heat9 = np.concatenate([np.concatenate([h[:, :1], h[:, 2:]], 1) for h, _ in captured])
    maps7 = DO.page_maps(heat9, offsets, page.shape[1], page.shape[0])
    assert np.array_equal(maps7[1], lines) and np.array_equal(maps7[2], seps), "oracle page maps differ from the reference's"
    print("page4: tiles", len(offsets), "peaks per tile", counts, "-> selected", loc.shape)
    np.savez_compressed(os.path.join(GOLD, "page4_seed0.npz"), seed=np.array(PAGE4["seed"]), image_hw=np.array([PAGE4["height"], PAGE4["width"]]),
                        offsets=np.array(offsets), page_hw=np.array(page.shape[:2]), pre_counts=np.array(counts),
                        pre_locations=pre_loc.astype(np.float32), pre_glyphfeatures=pre_gf.astype(np.float32),
                        locations=loc, glyphfeatures=gf, maps7=maps7)

    # dense stub page
    page, offsets = page_tiles(synthetic.page_image(DENSE["seed"], DENSE["height"], DENSE["width"]))
    heat10 = synthetic.dense_page_heatmaps(DENSE["seed"], len(offsets)).numpy()
    feats = dense_features(len(offsets))

    class Stub(OCR_Processer):
        def __init__(self):
            super().__init__()
            self.n = 0

        def call_detector(self, image_input):
            self.n += 1
            return heat10[self.n - 1:self.n], feats[self.n - 1:self.n]

        def call_transformer(self, encoder_input):
            raise NotImplementedError

    im, ds0 = tiles_of(page)
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        loc, gf, lines, seps = Stub().run_detector(ds0, im)
    n_pre = sum(len(DO.decode_tile(heat10[i], feats[i], x, y, page.shape[1], page.shape[0])[0]) for i, (x, y) in enumerate(offsets))
    print("dense: candidates", n_pre, "-> selected", loc.shape)
    np.savez_compressed(os.path.join(GOLD, "page_dense_seed0.npz"), seed=np.array(DENSE["seed"]), feat_seed=np.array(DENSE["feat_seed"]),
                        image_hw=np.array([DENSE["height"], DENSE["width"]]), offsets=np.array(offsets), page_hw=np.array(page.shape[:2]),
                        n_candidates=np.array(n_pre), locations=loc, glyphfeatures_sum=gf.astype(np.float64).sum(1),
                        lines_all_s3=lines[::3, ::3].copy(), seps_all_s3=seps[::3, ::3].copy())
    print("page goldens written")


def chunk_features(seed: int, n: int) -> np.ndarray:
    """Synthetic ``features`` [n, 106] of call_OCR (process_ocr_base.py:114-171): 100 glyph features + 5 x (vertical, rubybase,
    ruby, space, emphasis, newline) flags, with runs of ruby / ruby base, spaces, single and double newlines and a
    horizontal / vertical change, so that every split rule of the chunk loop (:187-283) fires."""
    rng = np.random.default_rng(seed)
    f = np.zeros((n, 106), dtype=np.float32)
    f[:, :100] = rng.standard_normal((n, 100)).astype(np.float32)
    vertical, i = 0, 0
    while i < n:
        run = int(rng.integers(5, 60))
        kind = rng.random()
        for k in range(i, min(n, i + run)):
            f[k, 100] = 5 * vertical
            if kind < 0.15:
                f[k, 101] = 5                       # ruby base run ...
            elif kind < 0.3:
                f[k, 102] = 5                       # ... ruby text run
            if rng.random() < 0.08:
                f[k, 103] = 5                       # space
        i += run
        if i < n:                                    # newline row(s) between runs
            f[i, :100] = 0
            f[i, 100] = 5 * vertical
            f[i, 105] = 5
            i += 1
            if rng.random() < 0.3 and i < n:
                f[i, :100] = 0
                f[i, 100] = 5 * vertical
                f[i, 105] = 5
                i += 1
            if rng.random() < 0.2:
                vertical ^= 1
    return f


def stub_codes(encoder_input: np.ndarray) -> np.ndarray:
    """Deterministic stand-in for call_transformer in the chunk goldens: one code point per feature row between the SP tokens
    (a function of the row's first feature), SOT first, EOT after the last."""
    x = encoder_input[0]
    sp = np.zeros(106, dtype=np.float32); sp[0:100:2] = 5; sp[1:100:2] = -5
    end = next(k for k in range(1, x.shape[0]) if np.array_equal(x[k], -sp))
    out = np.zeros(x.shape[0], dtype=np.int64)
    out[0] = 1
    for k in range(1, end):
        out[k] = 0x3042 + int(abs(float(x[k, 0])) * 7) % 80 if x[k, 105] == 0 else 0x0A
    out[end] = 2
    return out


def golden_chunks():
    """tests/golden/chunks_seed0.json: the chunk windows, overlaps and assembled text of the UNMODIFIED reference loop
    (process_ocr_base.py:187-283, executed from the reference's own source text with a stub call_transformer)."""
    import json, textwrap
    from const import max_encoderlen, decoder_SOT, decoder_EOT, decoder_PAD
    encoder_dim = 106
    src = open(os.path.join(REF, "process_ocr_base.py")).read().split("\n")
    body = textwrap.dedent("\n".join(src[181:283]))           # cur_i = 0 ... end of the while loop (:182-283)
    assert body.startswith("cur_i = 0") and "linebuf += " in body, body[:80]
    out = {}
    for seed, n in ((0, 37), (1, 420), (2, 1300), (3, 800), (4, 396), (5, 397), (6, 398)):
        features = chunk_features(seed, n)
        calls = []

        class Self:
            def call_transformer(self, encoder_input):
                calls.append(encoder_input.copy())
                return stub_codes(encoder_input)

        SP_token = np.zeros([encoder_dim], dtype=np.float32)
        SP_token[0:100:2] = 5
        SP_token[1:100:2] = -5
        ns = dict(np=np, features=features, self=Self(), max_encoderlen=max_encoderlen, encoder_dim=encoder_dim, SP_token=SP_token,
                  decoder_SOT=decoder_SOT, decoder_EOT=decoder_EOT, decoder_PAD=decoder_PAD)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            exec(compile(body, "process_ocr_base.py:182-283", "exec"), ns)
        windows = [[int(v) for v in ln.split("/")[0].split()] for ln in buf.getvalue().strip().split("\n") if ln.strip()]
        out[f"s{seed}_n{n}"] = dict(seed=seed, n=n, windows=windows, result_txt=ns["result_txt"],
                                    linebuf=[[int(a), int(b), c] for a, b, c in ns["linebuf"]],
                                    input_sums=[float(np.abs(c).sum()) for c in calls])
        print(seed, n, "chunks", len(windows), "text", len(ns["result_txt"]))
    with open(os.path.join(GOLD, "chunks_seed0.json"), "w") as f:
        json.dump(out, f)
    print("chunk goldens written")


TRANSFORMER_CFGS = {
    # name: (dims, batch, predictor max_decoderlen)
    "tiny": (dict(embed_dim=64, head_num=4, enc_block_num=2, dec_block_num=2, max_enc_seq_len=24, max_dec_seq_len=24), 3),
    "cfg4": (dict(embed_dim=512, head_num=16, enc_block_num=16, dec_block_num=16, max_enc_seq_len=100, max_dec_seq_len=100), 2),
    "default": (dict(), 1),
}


def golden_transformer():
    import models.transformer as T
    import json
    out = {}
    keys = {}
    for name, (dims, batch) in TRANSFORMER_CFGS.items():
        cfg = T.ModelDimensions(**dims)
        model = T.Transformer(**cfg.__dict__).eval()
        keys[name] = [(k, list(v.shape)) for k, v in model.state_dict().items()]
        sd = synthetic.transformer_state_dict(0, **cfg.__dict__)
        model.load_state_dict(sd, strict=True)
        le, ld = cfg.max_enc_seq_len, cfg.max_dec_seq_len
        enc, dec, lens = synthetic.transformer_inputs(batch, le, ld, seed=0)
        with torch.no_grad():
            logits = model(enc, dec)
        for i, lg in enumerate(logits):
            lg = lg.numpy()
            out[f"{name}_logits{i}_s"] = lg[:, ::max(1, ld // 8)].copy()      # 8-ish positions, all classes
            out[f"{name}_argmax{i}"] = lg.argmax(-1).astype(np.int32)
            out[f"{name}_lse{i}"] = torch.logsumexp(torch.from_numpy(lg), -1).numpy()
        # predictor (hard-codes const.max_decoderlen at models/transformer.py:278 -> patch the module global)
        old = T.max_decoderlen
        T.max_decoderlen = ld
        pred = T.TransformerPredictor(model.encoder, model.decoder).eval()
        buf = io.StringIO()
        with torch.no_grad(), contextlib.redirect_stdout(buf):
            ids = pred(enc)
        T.max_decoderlen = old
        out[f"{name}_pred_ids"] = ids.numpy()
        out[f"{name}_pred_log"] = np.array(buf.getvalue())
        print(name, "logits ok; predictor:", buf.getvalue().strip() or "(8 passes)", ids[0, :8].tolist())
        if name == "tiny":
            # early-exit variants: bias the three output heads towards the residues of U+3042 so that the
            # candidate probability is ~1 ("early stop") or ~0.95 ("no remask stop"), models/transformer.py:326,356
            for vname, boost in (("peaked", 20.0), ("medium", 10.5)):
                sd2 = dict(sd)
                for i, m in enumerate(arch.MODULO_LIST):
                    bias = sd[f"decoder.out_layers.{i}.bias"].clone()
                    bias[0x3042 % m] += boost
                    sd2[f"decoder.out_layers.{i}.bias"] = bias
                model.load_state_dict(sd2, strict=True)
                T.max_decoderlen = ld
                buf = io.StringIO()
                with torch.no_grad(), contextlib.redirect_stdout(buf):
                    ids = T.TransformerPredictor(model.encoder, model.decoder).eval()(enc)
                T.max_decoderlen = old
                out[f"tiny_{vname}_pred_ids"] = ids.numpy()
                out[f"tiny_{vname}_pred_log"] = np.array(buf.getvalue())
                print("  variant", vname, buf.getvalue().strip(), ids[0, :4].tolist())
    # CRT golden: calc_predid on a seeded batch
    from util_func import calc_predid
    g = torch.Generator().manual_seed(5)
    b = [torch.randint(0, m, (4096,), generator=g) for m in arch.MODULO_LIST]
    out["crt_in"] = torch.stack(b).numpy()
    out["crt_out"] = calc_predid(*b).numpy()
    np.savez_compressed(os.path.join(GOLD, "transformer_seed0.npz"), **out)
    with open(os.path.join(GOLD, "transformer_state_keys.json"), "w") as f:
        json.dump(keys, f)
    print("transformer goldens written")


OPT_SHAPES = [(257,), (64, 33), (8, 4, 3, 3), (10000,)]
OPT_CFG = dict(lr=2.5e-3, weight_decay=1e-2, warmup_steps=3)


def optimizer_inputs(step):
    g = torch.Generator().manual_seed(4242 + step)
    return [torch.randn(s, generator=g) * (0.5 + step) for s in OPT_SHAPES]


def golden_optimizer():
    """models/adamw_schedulefree.py::AdamWScheduleFree: 5 steps (3 warm-up), weight decay, then optimizer.eval()."""
    from models.adamw_schedulefree import AdamWScheduleFree
    params = [torch.nn.Parameter(p.clone()) for p in optimizer_inputs(-1)]
    opt = AdamWScheduleFree(params, **OPT_CFG)
    opt.train()
    out = {}
    for step in range(5):
        for p, g in zip(params, optimizer_inputs(step)):
            p.grad = g.clone()
        opt.step()
        for i, p in enumerate(params):
            out[f"s{step}_y{i}"] = p.detach().numpy().copy()
            out[f"s{step}_z{i}"] = opt.state[p]["z"].numpy().copy()
            out[f"s{step}_v{i}"] = opt.state[p]["exp_avg_sq"].numpy().copy()
    opt.eval()
    for i, p in enumerate(params):
        out[f"eval_x{i}"] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(GOLD, "optimizer_seed0.npz"), **out)
    print("optimizer goldens written")


RADAM_CFGS = {"silent": dict(lr=2.5e-3, weight_decay=1e-2), "sgd": dict(lr=2.5e-3, silent_sgd_phase=False, betas=(0.9, 0.99))}


def golden_radam():
    """models/radam_schedulefree.py::RAdamScheduleFree: 8 steps crossing rho_t = 4 (silent phase, and the SGD-phase variant
    with beta2 = 0.99), then optimizer.eval()."""
    from models.radam_schedulefree import RAdamScheduleFree
    out = {}
    for name, cfg in RADAM_CFGS.items():
        params = [torch.nn.Parameter(p.clone()) for p in optimizer_inputs(-1)]
        opt = RAdamScheduleFree(params, **cfg)
        opt.train()
        for step in range(8):
            for p, g in zip(params, optimizer_inputs(step)):
                p.grad = g.clone()
            opt.step()
            for i, p in enumerate(params[:3]):             # the three small tensors at every step; all four after eval()
                out[f"{name}_s{step}_y{i}"] = p.detach().numpy().copy()
                out[f"{name}_s{step}_z{i}"] = opt.state[p]["z"].numpy().copy()
                out[f"{name}_s{step}_v{i}"] = opt.state[p]["exp_avg_sq"].numpy().copy()
            out[f"{name}_s{step}_lr"] = np.float64(opt.param_groups[0]["scheduled_lr"])
        opt.eval()
        for i, p in enumerate(params):
            out[f"{name}_eval_x{i}"] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(GOLD, "optimizer_radam_seed0.npz"), **out)
    print("radam goldens written")


def golden_loss():
    """loss_func.py: loss_function, loss_function3, CoVWeightingLoss (6 iterations) and autograd d loss / d heatmap."""
    import loss_func as R
    x = synthetic.loss_inputs(0)
    out = {}
    hm = x["heatmap"].clone().requires_grad_(True)
    res = R.loss_function(x["fmask"], x["labelmap"], x["idmap"], hm, [x["dec0"], x["dec1"], x["dec2"]])
    for k, v in res.items():
        out["train1_" + k] = np.asarray(v.detach().numpy() if torch.is_tensor(v) else v)
    names = ["keymap_loss", "size_loss", "textline_loss", "separator_loss", "code1_loss", "code2_loss", "code4_loss", "code8_loss"]
    alphas = torch.tensor([0.3, 0.05, 0.1, 0.15, 0.1, 0.1, 0.1, 0.1])
    sum(a * res[k] for a, k in zip(alphas, names)).backward()
    out["train1_alphas"] = alphas.numpy()
    out["train1_grad_heatmap"] = hm.grad.numpy()
    res3 = R.loss_function3([x["out3_0"], x["out3_1"], x["out3_2"]], x["labelcode"], x["mask3"])
    for k, v in res3.items():
        out["train3_" + k] = np.asarray(v.detach().numpy() if torch.is_tensor(v) else v)
    cov = R.CoVWeightingLoss(losses=names + ["id_loss"])
    g = torch.Generator().manual_seed(99)
    seq, tot, alph = [], [], []
    for it in range(6):
        vals = torch.rand(9, generator=g) * (2.0 / (1 + it)) + 0.1
        seq.append(vals.numpy())
        tot.append(float(cov({k: vals[i] for i, k in enumerate(names + ["id_loss"])})))
        alph.append(cov.alphas.numpy().copy())
    out["cov_inputs"], out["cov_totals"], out["cov_alphas"] = np.stack(seq), np.asarray(tot, np.float32), np.stack(alph)
    np.savez_compressed(os.path.join(GOLD, "loss_seed0.npz"), **out)
    print("loss goldens written:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    if what in ("detector", "all"):
        golden_detector()
    if what in ("page", "all"):
        golden_page()
    if what in ("chunks", "all"):
        golden_chunks()
    if what in ("transformer", "all"):
        golden_transformer()
    if what in ("optimizer", "all"):
        golden_optimizer()
    if what in ("radam", "all"):
        golden_radam()
    if what in ("loss", "all"):
        golden_loss()
