"""Architecture tables for the detector hot path.

Single source of truth for layer shapes and ``state_dict`` key names, shared by the
``nn.Module`` mirror (models/), the synthetic-checkpoint generator and the C++ engine
plan (csrc/engine.cu receives the stage table through the C-ABI).

Reference: /root/reference/models/detector.py:12-28 (``efficientnet_v2_xl`` stage table),
:148-201 (``Leafmap``), :232-254 (``SimpleDecoder``); torchvision
``models/efficientnet.py:105-231`` (MBConv / FusedMBConv block layout).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List, Tuple

# util_func.py:5-9 / const.py:1-15 of the reference
MODULO_LIST = (1091, 1093, 1097)
WIDTH = 768
HEIGHT = 768
SCALE = 4
FEATURE_DIM = 100
ENCODER_ADD_DIM = 6
ENCODER_DIM = FEATURE_DIM + ENCODER_ADD_DIM
MAX_DECODERLEN = 400
MAX_ENCODERLEN = 400
DECODER_PAD, DECODER_SOT, DECODER_EOT, DECODER_MSK = 0, 1, 2, 3

BACKBONE_BN_EPS = 1e-3   # models/detector.py:27
HEAD_BN_EPS = 1e-5       # nn.BatchNorm2d default, models/detector.py:162-187
CONV_DIM = 192           # models/detector.py:159
DECODER_MID_DIM = 2048   # models/detector.py:236

# head name -> out_dim, in heatmap channel order (models/detector.py:207-230);
# "sepatator" is the reference's own spelling and is part of the state_dict contract.
HEADS: Tuple[Tuple[str, int], ...] = (
    ("keyheatmap", 1), ("sizes", 2), ("textline", 1), ("sepatator", 1),
    ("code1", 1), ("code2", 1), ("code4", 1), ("code8", 1), ("feature", FEATURE_DIM),
)


@dataclass(frozen=True)
class StageCfg:
    fused: bool      # FusedMBConv (True) or MBConv (False)
    expand: int
    kernel: int
    stride: int
    cin: int
    cout: int
    layers: int


# (stem_out, stages, last_channel, tap in_dims)
_BACKBONES = {
    # models/detector.py:13-21
    "xl": (32, [StageCfg(True, 1, 3, 1, 32, 32, 4), StageCfg(True, 4, 3, 2, 32, 64, 8),
                StageCfg(True, 4, 3, 2, 64, 96, 8), StageCfg(False, 4, 3, 2, 96, 192, 16),
                StageCfg(False, 6, 3, 1, 192, 256, 24), StageCfg(False, 6, 3, 2, 256, 512, 32),
                StageCfg(False, 6, 3, 1, 512, 640, 8)], 1280),
    # torchvision efficientnet.py _efficientnet_conf("efficientnet_v2_{l,m,s}")
    "l": (32, [StageCfg(True, 1, 3, 1, 32, 32, 4), StageCfg(True, 4, 3, 2, 32, 64, 7),
               StageCfg(True, 4, 3, 2, 64, 96, 7), StageCfg(False, 4, 3, 2, 96, 192, 10),
               StageCfg(False, 6, 3, 1, 192, 224, 19), StageCfg(False, 6, 3, 2, 224, 384, 25),
               StageCfg(False, 6, 3, 1, 384, 640, 7)], 1280),
    "m": (24, [StageCfg(True, 1, 3, 1, 24, 24, 3), StageCfg(True, 4, 3, 2, 24, 48, 5),
               StageCfg(True, 4, 3, 2, 48, 80, 5), StageCfg(False, 4, 3, 2, 80, 160, 7),
               StageCfg(False, 6, 3, 1, 160, 176, 14), StageCfg(False, 6, 3, 2, 176, 304, 18),
               StageCfg(False, 6, 3, 1, 304, 512, 5)], 1280),
    "s": (24, [StageCfg(True, 1, 3, 1, 24, 24, 2), StageCfg(True, 4, 3, 2, 24, 48, 4),
               StageCfg(True, 4, 3, 2, 48, 64, 4), StageCfg(False, 4, 3, 2, 64, 128, 6),
               StageCfg(False, 6, 3, 1, 128, 160, 9), StageCfg(False, 6, 3, 2, 160, 256, 15)], 1280),
}

# backbone.features indices whose outputs are the Leafmap taps (models/detector.py:139-146)
TAP_FEATURE_IDX = (2, 3, 5)


def backbone_cfg(model_size: str = "xl"):
    stem, stages, last = _BACKBONES[model_size]
    return stem, list(stages), last


def tap_dims(model_size: str = "xl") -> List[int]:
    """Channel counts of taps x1..x4 (models/detector.py:151-158)."""
    _, stages, last = backbone_cfg(model_size)
    return [stages[1].cout, stages[2].cout, stages[4].cout, last]


def se_squeeze(cin: int) -> int:
    """torchvision MBConv: squeeze_channels = max(1, input_channels // 4)."""
    return max(1, cin // 4)


@dataclass(frozen=True)
class ParamSpec:
    key: str
    shape: Tuple[int, ...]
    kind: str   # conv | dwconv | bn_w | bn_b | bn_mean | bn_var | bn_count | bias | linear | embed | ln_w | ln_b | posenc
    fan_in: int = 0


def _bn(prefix: str, c: int) -> Iterator[ParamSpec]:
    yield ParamSpec(prefix + ".weight", (c,), "bn_w")
    yield ParamSpec(prefix + ".bias", (c,), "bn_b")
    yield ParamSpec(prefix + ".running_mean", (c,), "bn_mean")
    yield ParamSpec(prefix + ".running_var", (c,), "bn_var")
    yield ParamSpec(prefix + ".num_batches_tracked", (), "bn_count")


def _conv_bn(prefix: str, cin: int, cout: int, k: int, groups: int = 1) -> Iterator[ParamSpec]:
    kind = "dwconv" if groups > 1 else "conv"
    yield ParamSpec(prefix + ".0.weight", (cout, cin // groups, k, k), kind, (cin // groups) * k * k)
    yield from _bn(prefix + ".1", cout)


def backbone_specs(prefix: str, model_size: str = "xl") -> Iterator[ParamSpec]:
    stem, stages, last = backbone_cfg(model_size)
    yield from _conv_bn(f"{prefix}.0", 3, stem, 3)
    for si, st in enumerate(stages, start=1):
        for li in range(st.layers):
            cin = st.cin if li == 0 else st.cout
            exp = cin * st.expand
            p = f"{prefix}.{si}.{li}.block"
            if st.fused:
                if exp != cin:
                    yield from _conv_bn(p + ".0", cin, exp, st.kernel)
                    yield from _conv_bn(p + ".1", exp, st.cout, 1)
                else:
                    yield from _conv_bn(p + ".0", cin, st.cout, st.kernel)
            else:
                sq = se_squeeze(cin)
                yield from _conv_bn(p + ".0", cin, exp, 1)
                yield from _conv_bn(p + ".1", exp, exp, st.kernel, groups=exp)
                yield ParamSpec(p + ".2.fc1.weight", (sq, exp, 1, 1), "conv", exp)
                yield ParamSpec(p + ".2.fc1.bias", (sq,), "bias")
                yield ParamSpec(p + ".2.fc2.weight", (exp, sq, 1, 1), "conv", sq)
                yield ParamSpec(p + ".2.fc2.bias", (exp,), "bias")
                yield from _conv_bn(p + ".3", exp, st.cout, 1)
    yield from _conv_bn(f"{prefix}.{len(stages) + 1}", stages[-1].cout, last, 1)


def leafmap_specs(prefix: str, out_dim: int, model_size: str = "xl") -> Iterator[ParamSpec]:
    dims = tap_dims(model_size)
    for i, d in enumerate(dims):
        yield from _bn(f"{prefix}.in_bn.{i}", d)
    for i, d in enumerate(reversed(dims)):
        cin = d if i == 0 else d + CONV_DIM
        yield from _conv_bn(f"{prefix}.upsamplers.{i}", cin, CONV_DIM, 3)
    yield ParamSpec(f"{prefix}.top_conv.0.weight", (out_dim, CONV_DIM, 3, 3), "conv", CONV_DIM * 9)
    yield ParamSpec(f"{prefix}.top_conv.0.bias", (out_dim,), "bias")


def simple_decoder_specs(prefix: str) -> Iterator[ParamSpec]:
    for i, m in enumerate(MODULO_LIST):
        p = f"{prefix}.blocks.{i}"
        yield ParamSpec(p + ".0.weight", (DECODER_MID_DIM, FEATURE_DIM), "linear", FEATURE_DIM)
        yield from _bn(p + ".1", DECODER_MID_DIM)
        yield ParamSpec(p + ".3.weight", (DECODER_MID_DIM, DECODER_MID_DIM), "linear", DECODER_MID_DIM)
        yield from _bn(p + ".4", DECODER_MID_DIM)
        yield ParamSpec(p + ".6.weight", (m, DECODER_MID_DIM), "linear", DECODER_MID_DIM)
        yield ParamSpec(p + ".6.bias", (m,), "bias")


def detection_specs(prefix: str = "detector", model_size: str = "xl") -> List[ParamSpec]:
    """CenterNetDetection state_dict entries in the reference's registration order."""
    out = list(backbone_specs(f"{prefix}.backbone.features", model_size))
    for name, od in HEADS:
        out.extend(leafmap_specs(f"{prefix}.{name}", od, model_size))
    return out


def text_detector_specs(model_size: str = "xl") -> List[ParamSpec]:
    """TextDetectorModel state_dict entries (2444 for 'xl')."""
    return detection_specs("detector", model_size) + list(simple_decoder_specs("decoder"))


def transformer_specs(enc_input_dim=ENCODER_DIM, embed_dim=768, head_num=12, enc_block_num=10,
                      dec_block_num=10, max_enc_seq_len=MAX_ENCODERLEN, max_dec_seq_len=MAX_DECODERLEN,
                      dropout=0.0) -> List[ParamSpec]:
    """Transformer state_dict entries (models/transformer.py:139-246), registration order."""
    d = embed_dim
    out: List[ParamSpec] = []

    def mha(p, maxlen):
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            out.append(ParamSpec(f"{p}.{n}.weight", (d, d), "linear", d))
        out.append(ParamSpec(f"{p}.pos_emb_q.encoding", (maxlen, d), "posenc"))
        out.append(ParamSpec(f"{p}.pos_emb_k.encoding", (maxlen, d), "posenc"))

    def ln(p):
        out.append(ParamSpec(p + ".weight", (d,), "ln_w"))
        out.append(ParamSpec(p + ".bias", (d,), "ln_b"))

    def ff(p):
        for n, (o, i) in (("w1", (2 * d, d)), ("wg", (2 * d, d)), ("w2", (d, 2 * d))):
            out.append(ParamSpec(f"{p}.{n}.weight", (o, i), "linear", i))
            out.append(ParamSpec(f"{p}.{n}.bias", (o,), "bias"))

    out.append(ParamSpec("encoder.embed.weight", (d, enc_input_dim), "linear", enc_input_dim))
    out.append(ParamSpec("encoder.pos_emb.encoding", (max_enc_seq_len, d), "posenc"))
    ln("encoder.norm")
    for b in range(enc_block_num):
        p = f"encoder.blocks.{b}"
        mha(p + ".mha", max_enc_seq_len)
        ln(p + ".norm1"); ln(p + ".norm2")
        ff(p + ".ff")
    for i, m in enumerate(MODULO_LIST):
        out.append(ParamSpec(f"decoder.embed.{i}.weight", (m, d), "embed"))
    out.append(ParamSpec("decoder.pos_emb.encoding", (max_dec_seq_len, d), "posenc"))
    ln("decoder.norm")
    for b in range(dec_block_num):
        p = f"decoder.blocks.{b}"
        mha(p + ".self_attn", max_dec_seq_len)
        mha(p + ".cross_attn", max_dec_seq_len)
        ln(p + ".norm1"); ln(p + ".norm2"); ln(p + ".norm3")
        ff(p + ".ff")
    for i, m in enumerate(MODULO_LIST):
        out.append(ParamSpec(f"decoder.out_layers.{i}.weight", (m, d), "linear", d))
        out.append(ParamSpec(f"decoder.out_layers.{i}.bias", (m,), "bias"))
    return out
