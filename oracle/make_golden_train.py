"""Generate tests/golden/train_xl64_seed0.npz: ONE train-mode forward + backward of the UNMODIFIED reference
``TextDetectorModel`` (imported from /root/reference, torch autograd on CPU, fp32) on the seeded synthetic checkpoint.

Runs only in the build container (the GPU box has no /root/reference); the output is committed.
    python oracle/make_golden_train.py

Contents: input x [2,3,64,64] (the reference network is fully convolutional; 64 px keeps the XL model's 2 444-entry state
dict but makes the CPU run seconds), a fixed fmask, the heatmap and the three SimpleDecoder outputs, the random cotangents
w_heat / w_dec{i} that define the scalar  L = sum(heat * w_heat) + sum_i sum(dec_i * w_dec_i),  and for EVERY parameter the
gradient's L2 norm and its dot product with a deterministic +-1 probe (two fingerprints per tensor instead of 262 M floats),
plus a few small gradients and updated BatchNorm buffers in full.  StochasticDepth is switched off (p = 0) on the
reference's blocks so the run is deterministic; train_ops.py is called with sd_prob = 0 to match.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
GOLD = os.path.join(ROOT, "tests", "golden")

from findtextcenternet_b200 import synthetic  # noqa: E402

FULL = ("detector.backbone.features.0.1.weight", "detector.backbone.features.0.0.weight",
        "detector.backbone.features.4.0.block.1.0.weight", "detector.backbone.features.4.0.block.2.fc1.weight",
        "detector.backbone.features.4.0.block.2.fc2.bias", "detector.backbone.features.7.7.block.3.1.bias",
        "detector.keyheatmap.in_bn.3.weight", "detector.sizes.top_conv.0.bias", "detector.feature.upsamplers.3.1.weight",
        "decoder.blocks.1.4.bias", "decoder.blocks.2.6.bias")
BUFS = ("detector.backbone.features.0.1.running_mean", "detector.backbone.features.0.1.running_var",
        "detector.backbone.features.6.3.block.1.1.running_var", "detector.code4.upsamplers.2.1.running_mean",
        "decoder.blocks.0.1.running_var", "detector.backbone.features.0.1.num_batches_tracked")


def probe(shape, i):
    n = int(np.prod(shape)) if len(shape) else 1
    idx = np.arange(n, dtype=np.int64)
    v = (((idx * 2654435761 + i * 40503) >> 7) & 1).astype(np.float64) * 2 - 1
    return v.reshape(shape)


def run(dtype):
    from models.detector import TextDetectorModel
    from torchvision.ops import StochasticDepth
    torch.manual_seed(0)
    model = TextDetectorModel(pre_weights=False)
    model.load_state_dict(synthetic.detector_state_dict(0), strict=True)
    model = model.to(dtype).train()
    for m in model.modules():
        if isinstance(m, StochasticDepth):
            m.p = 0.0
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(2, 3, 64, 64, generator=g)
    fmask = torch.rand(2 * 16 * 16, generator=g) < 0.25
    heat, dec = model(x.to(dtype), fmask)
    w_heat = torch.randn(heat.shape, generator=g)
    w_dec = [torch.randn(d.shape, generator=g) / 30 for d in dec]
    loss = (heat * w_heat.to(dtype)).sum() + sum((d * w.to(dtype)).sum() for d, w in zip(dec, w_dec))
    loss.backward()
    return model, x, fmask, heat.detach(), [d.detach() for d in dec], w_heat, w_dec, float(loss.detach())


def main():
    # truth = the reference in float64 (same modules, .double()); the fp32 run of the same reference gives, per parameter,
    # the rounding noise a correct fp32 implementation is entitled to (BatchNorm betas / conv biases that feed another
    # batch-statistics layer have near-zero true gradients and fp32 noise far above them)
    model, x, fmask, heat, dec, w_heat, w_dec, loss = run(torch.float64)
    model32, _, _, heat32, dec32, _, _, loss32 = run(torch.float32)
    out = {"x": x.numpy(), "fmask": fmask.numpy(), "heatmap": heat.float().numpy(), "w_heat": w_heat.numpy(),
           "heatmap_fp32_err": np.array(float((heat32.double() - heat).norm() / heat.norm()))}
    for i in range(3):
        out[f"dec{i}"] = dec[i].float().numpy()
        out[f"w_dec{i}"] = w_dec[i].numpy()
    names, norms, dots, errs = [], [], [], []
    p32 = dict(model32.named_parameters())
    for i, (n, p) in enumerate(model.named_parameters()):
        assert p.grad is not None, n
        gd = p.grad.numpy()
        names.append(n)
        norms.append(float(np.linalg.norm(gd)))
        dots.append(float((gd * probe(gd.shape, i)).sum()))
        errs.append(float(np.linalg.norm(p32[n].grad.double().numpy() - gd)))
    out["grad_names"] = np.array(names)
    out["grad_norm"] = np.array(norms)
    out["grad_dot"] = np.array(dots)
    out["grad_fp32_err"] = np.array(errs)
    params = dict(model.named_parameters())
    for k in FULL:
        out["full/" + k] = params[k].grad.float().numpy()
    bufs = dict(model.named_buffers())
    for k in BUFS:
        out["buf/" + k] = bufs[k].double().numpy()
    path = os.path.join(GOLD, "train_xl64_seed0.npz")
    np.savez_compressed(path, **out)
    rel = np.array(errs) / (np.array(norms) + 1e-30)
    print("wrote", path, os.path.getsize(path), "bytes; loss", loss, loss32, "fmask rows", int(fmask.sum()))
    print("reference fp32-vs-fp64 gradient error: median %.2e, p99 %.2e, max %.2e; heatmap %.2e" %
          (np.median(rel), np.quantile(rel, 0.99), rel.max(), float(out["heatmap_fp32_err"])))


TF_DIMS = dict(enc_input_dim=106, embed_dim=64, head_num=4, enc_block_num=2, dec_block_num=2, max_enc_seq_len=24,
               max_dec_seq_len=24, dropout=0.0)   # cross-attn pos_emb_k is sized by the DECODER length (models/transformer.py:186)
TF_FULL = ("encoder.embed.weight", "encoder.pos_emb.encoding", "encoder.blocks.1.mha.pos_emb_q.encoding",
           "encoder.blocks.0.mha.k_proj.weight", "encoder.blocks.1.norm2.weight", "encoder.blocks.0.ff.wg.bias",
           "decoder.embed.1.weight", "decoder.blocks.1.cross_attn.pos_emb_k.encoding", "decoder.blocks.0.cross_attn.v_proj.weight",
           "decoder.blocks.1.norm3.bias", "decoder.blocks.0.ff.w2.weight", "decoder.out_layers.2.bias")


def run_transformer(dtype):
    from models.transformer import Transformer
    model = Transformer(**TF_DIMS)
    dims = {k: v for k, v in TF_DIMS.items() if k != "dropout"}
    model.load_state_dict(synthetic.transformer_state_dict(0, **dims), strict=True)
    model = model.to(dtype).train()
    enc, dec, _ = synthetic.transformer_inputs(3, 24, 16, 0)
    # Transformer.forward (models/transformer.py:248-253) line by line, with the mask cast to the run's dtype: torch's CPU SDPA
    # silently mis-handles a float32 additive mask next to float64 q/k/v (encoder output off by 50 %), which would corrupt
    # the float64 truth; in float32 this is exactly model(enc, dec)
    key_mask = torch.all(enc == 0, dim=-1)
    key_mask = torch.where(key_mask[:, None, None, :], float("-inf"), 0).to(dtype)
    enc_output = model.encoder(enc.to(dtype), key_mask=key_mask)
    outs = model.decoder(dec, enc_output, key_mask=key_mask)
    g = torch.Generator().manual_seed(4321)
    ws = [torch.randn(o.shape, generator=g) / 10 for o in outs]
    sum((o * w.to(dtype)).sum() for o, w in zip(outs, ws)).backward()
    return model, enc, dec, [o.detach() for o in outs], ws


def main_transformer():
    """tests/golden/train_transformer_seed0.npz: Transformer.forward in train mode + backward (train3.py:132-137) of the
    unmodified reference, float64 truth + the fp32 run's per-tensor noise, every gradient fingerprinted."""
    model, enc, dec, outs, ws = run_transformer(torch.float64)
    model32, _, _, outs32, _ = run_transformer(torch.float32)
    out = {"enc": enc.numpy(), "dec": dec.numpy()}
    for i in range(3):
        out[f"out{i}"] = outs[i].float().numpy()
        out[f"w{i}"] = ws[i].numpy()
    names, norms, dots, errs = [], [], [], []
    p32 = dict(model32.named_parameters())
    for i, (n, p) in enumerate(model.named_parameters()):
        names.append(n)
        if p.grad is None:      # self-attention never touches pos_emb_k (models/transformer.py:107-109): no gradient at all
            norms.append(-1.0); dots.append(0.0); errs.append(0.0)
            continue
        gd = p.grad.numpy()
        norms.append(float(np.linalg.norm(gd)))
        dots.append(float((gd * probe(gd.shape, i)).sum()))
        errs.append(float(np.linalg.norm(p32[n].grad.double().numpy() - gd)))
    out["grad_names"], out["grad_norm"], out["grad_dot"], out["grad_fp32_err"] = (np.array(names), np.array(norms), np.array(dots),
                                                                                   np.array(errs))
    params = dict(model.named_parameters())
    for k in TF_FULL:
        out["full/" + k] = params[k].grad.float().numpy()
    # the export-side wrappers (models/transformer.py:362-404) of the same fp32 model in eval mode
    from models.transformer import TransformerEncoderPredictor, TransformerDecoderPredictor
    m32 = model32.eval()
    with torch.no_grad():
        km = torch.where(torch.all(enc == 0, dim=-1)[:, None, None, :], float("-inf"), 0)
        enc_out = TransformerEncoderPredictor(m32.encoder)(enc, km)
        probs = TransformerDecoderPredictor(m32.decoder)(enc_out, dec, km)
    out["pred_enc_out"] = enc_out.numpy()
    for i in range(3):
        out[f"pred_probs{i}_max"] = probs[i].max(-1).values.numpy()
        out[f"pred_probs{i}_argmax"] = probs[i].argmax(-1).numpy()
    path = os.path.join(GOLD, "train_transformer_seed0.npz")
    np.savez_compressed(path, **out)
    rel = np.array(errs)[np.array(norms) > 0] / np.array(norms)[np.array(norms) > 0]
    print("wrote", path, os.path.getsize(path), "bytes;", len(names), "gradients; reference fp32-vs-fp64 error median %.2e max %.2e"
          % (np.median(rel), rel.max()))


def main_pagemaps():
    """tests/golden/page_maps_seed0.npz: lines_all / seps_all as returned by the UNMODIFIED reference ``run_detector``
    (process_ocr_base.py:474-650) for a 900x1000 page whose four tiles get seeded random heatmaps from a stub backend (peak
    channel at -10 so the greedy box selection has nothing to do); stored with the 9-channel heatmaps that produced them."""
    from process_ocr_base import OCR_Processer
    from findtextcenternet_b200.process_ocr_b200 import page_tiles
    g = torch.Generator().manual_seed(77)
    page, offsets = page_tiles(np.full((900, 1000, 3), 255, dtype=np.uint8))
    heat10 = [torch.randn(1, 10, 192, 192, generator=g).mul_(2.0).numpy() for _ in offsets]
    for h in heat10:
        h[0, 1] = -10.0

    class Stub(OCR_Processer):
        def __init__(self):
            super().__init__()
            self.n = 0

        def call_detector(self, image_input):
            h = heat10[self.n]
            self.n += 1
            return h, np.zeros((1, 100, 192, 192), dtype=np.float32)

        def call_transformer(self, encoder_input):
            raise NotImplementedError

    ds = [{"input": None, "offsetx": x, "offsety": y} for x, y in offsets]
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        _, _, lines, seps = Stub().run_detector(ds, page)
    # the heatmaps are regenerated from the seed by the tests (page_maps_inputs below); the maps are kept at every 3rd pixel
    path = os.path.join(GOLD, "page_maps_seed0.npz")
    np.savez_compressed(path, seed=np.array(77), offsets=np.array(offsets), page_hw=np.array(page.shape[:2]),
                        lines_all_s3=lines[::3, ::3].copy(), seps_all_s3=seps[::3, ::3].copy())
    print("wrote", path, os.path.getsize(path), "bytes", lines.shape, float(lines.max()))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("pagemaps", "all"):
        main_pagemaps()
    if what in ("detector", "all"):
        main()
    if what in ("transformer", "all"):
        main_transformer()
