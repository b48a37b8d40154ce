"""Drop-in mirror of the reference ``models/transformer.py`` module surface.

``ModelDimensions``, ``Transformer(**dims)`` with ``.encoder`` / ``.decoder`` / ``.head_num`` / ``.max_len``,
``TransformerPredictor(encoder, decoder)``; ``state_dict()`` keys/shapes/order equal the reference's
(/root/reference/models/transformer.py:139-246; 416 entries for the default dims), so ``model3.pt`` checkpoints
(``['model_state_dict']`` + ``['config']``, train3.py:205-241) load unchanged.  ``forward`` runs on the sm_100a engine
behind include/ftc_b200.h (ftc_transformer_forward / ftc_transformer_predict); no CPU or eager fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch
from torch import nn

from .. import arch
from ..engine import TransformerEngine, default_precision
from ._tree import Node, populate

encoder_dim = arch.ENCODER_DIM
max_decoderlen = arch.MAX_DECODERLEN   # TransformerPredictor reads this module global, as the reference does (:278)
max_encoderlen = arch.MAX_ENCODERLEN


@dataclass
class ModelDimensions:   # models/transformer.py:255-264
    enc_input_dim: int = encoder_dim
    embed_dim: int = 768
    head_num: int = 12
    enc_block_num: int = 10
    dec_block_num: int = 10
    max_enc_seq_len: int = max_encoderlen
    max_dec_seq_len: int = max_decoderlen
    dropout: float = 0.0


def _no_train(mod: nn.Module):
    if mod.training and torch.is_grad_enabled():
        raise NotImplementedError(
            "findtextcenternet_b200: TransformerPredictor is an inference loop; call .eval() / torch.no_grad()")


class Encoder(nn.Module):
    """Parameter tree of models/transformer.py:162-180 (embed, pos_emb, norm, blocks[*].{mha, norm1, norm2, ff})."""

    def __init__(self, input_dim, embed_dim, head_num, max_seq_len=5000, block_num=6, dropout=0.1):
        super().__init__()
        self.dim, self.head_num, self.input_dim, self.max_seq_len, self.block_num = embed_dim, head_num, input_dim, max_seq_len, block_num
        specs = [s for s in arch.transformer_specs(input_dim, embed_dim, head_num, block_num, 0, max_seq_len, 1)
                 if s.key.startswith("encoder.")]
        populate(self, specs, strip="encoder.", backbone_prefix="\0")


class Decoder(nn.Module):
    """Parameter tree of models/transformer.py:213-238 (embed[3], pos_emb, norm, blocks, out_layers[3])."""

    def __init__(self, embed_dim, head_num, max_seq_len=5000, block_num=6, dropout=0.1):
        super().__init__()
        self.dim, self.head_num, self.max_seq_len, self.block_num = embed_dim, head_num, max_seq_len, block_num
        specs = [s for s in arch.transformer_specs(encoder_dim, embed_dim, head_num, 0, block_num, 1, max_seq_len)
                 if s.key.startswith("decoder.")]
        populate(self, specs, strip="decoder.", backbone_prefix="\0")


class _EngineOwner:
    """Shared engine plumbing of Transformer / TransformerPredictor (both hold .encoder and .decoder)."""

    precision: str
    weights_frozen: bool

    def _init_engine(self):
        self.precision = default_precision()
        self.weights_frozen = False
        self._engine: Optional[TransformerEngine] = None
        self._engine_key = None

    def set_precision(self, precision: str):
        self.precision = precision
        return self

    def _dims(self) -> dict:
        e, d = self.encoder, self.decoder
        return dict(enc_input_dim=e.input_dim, embed_dim=e.dim, head_num=e.head_num, enc_block_num=e.block_num,
                    dec_block_num=d.block_num, max_enc_seq_len=e.max_seq_len, max_dec_seq_len=d.max_seq_len)

    def _tensors(self) -> Dict[str, torch.Tensor]:
        out = {"encoder." + k: v for k, v in self.encoder.state_dict(keep_vars=True).items()}
        out.update({"decoder." + k: v for k, v in self.decoder.state_dict(keep_vars=True).items()})
        return out

    def engine(self, device: torch.device) -> TransformerEngine:
        if (self.weights_frozen and self._engine is not None and self._engine_key is not None
                and self._engine.precision == self.precision and self._engine.device == device):
            return self._engine
        tensors = self._tensors()
        key = (str(device), self.precision, next(iter(tensors.values())).data_ptr(), hash(tuple(t._version for t in tensors.values())))
        if self._engine is None or self._engine.precision != self.precision or self._engine.device != device:
            self._engine = TransformerEngine(self._dims(), self.precision, device)
            self._engine_key = None
        if self._engine_key != key:
            self._engine.pack(tensors)
            self._engine_key = key
        return self._engine


class Transformer(nn.Module, _EngineOwner):
    """models/transformer.py:240-253: forward(enc_input [B,Le,106] f32, dec_input [B,Ld] int64) -> 3 x [B,Ld,m_i]."""

    def __init__(self, enc_input_dim, embed_dim, head_num, enc_block_num=6, dec_block_num=6, max_enc_seq_len=5000,
                 max_dec_seq_len=5000, dropout=0.1):
        super().__init__()
        self.head_num = head_num
        self.dropout = dropout
        self.max_len = max(max_enc_seq_len, max_dec_seq_len)
        self.encoder = Encoder(input_dim=enc_input_dim, embed_dim=embed_dim, head_num=head_num, max_seq_len=max_enc_seq_len,
                               block_num=enc_block_num, dropout=dropout)
        self.decoder = Decoder(embed_dim=embed_dim, head_num=head_num, max_seq_len=max_dec_seq_len, block_num=dec_block_num,
                               dropout=dropout)
        self._init_engine()

    def forward(self, enc_input, dec_input):
        if self.training and torch.is_grad_enabled():
            # train3.py:132-137: train mode with autograd on -> per-layer kernels with a tape (train_ops.py)
            from ..train_ops import transformer_train_forward
            return transformer_train_forward(self, enc_input, dec_input)
        if not enc_input.is_cuda:
            raise RuntimeError("findtextcenternet_b200 transformer: input must be a CUDA tensor (no CPU path)")
        with torch.no_grad():
            return self.engine(enc_input.device).forward(enc_input, dec_input)


class TransformerPredictor(nn.Module, _EngineOwner):
    """models/transformer.py:266-360: encoder once + <= 8 mask-predict decoder passes -> int64 [B, max_decoderlen]."""

    def __init__(self, encoder, decoder):
        super().__init__()
        self.head_num = encoder.head_num
        self.max_len = decoder.max_seq_len
        self.encoder = encoder
        self.decoder = decoder
        self.verbose = True      # the reference prints "[k early stop]" / "[k no remask stop]"
        self._init_engine()

    def forward(self, enc_input):
        _no_train(self)
        if not enc_input.is_cuda:
            raise RuntimeError("findtextcenternet_b200 transformer: input must be a CUDA tensor (no CPU path)")
        with torch.no_grad():
            ids, passes, reason = self.engine(enc_input.device).predict(enc_input, max_decoderlen, 8)
        if self.verbose and reason == 1:
            print(f"[{passes - 1} early stop]")
        elif self.verbose and reason == 2:
            print(f"[{passes - 1} no remask stop]")
        self.last_passes, self.last_stop_reason = passes, reason
        return ids

    def forward_each(self, enc_input):
        """A batch of INDEPENDENT sequences: each one stops by the rules the reference applies to a batch of one, so row i equals
        ``forward(enc_input[i:i+1])`` -- what ``call_OCR`` gets from its chunk-by-chunk ``call_transformer`` loop
        (process_ocr_base.py:235).  Prints the reference's stop line per sequence when ``verbose``."""
        _no_train(self)
        if not enc_input.is_cuda:
            raise RuntimeError("findtextcenternet_b200 transformer: input must be a CUDA tensor (no CPU path)")
        with torch.no_grad():
            ids, state = self.engine(enc_input.device).predict_each(enc_input, max_decoderlen, 8)
        self.last_state = state.cpu()
        if self.verbose:
            for _, passes, reason in self.last_state.tolist():
                if reason == 1:
                    print(f"[{passes - 1} early stop]")
                elif reason == 2:
                    print(f"[{passes - 1} no remask stop]")
        return ids


def _flat_mask(key_mask, batch: int):
    """The reference passes the additive key mask as [B,1,1,Le] (models/transformer.py:250); the kernels take [B, Le] fp32."""
    if key_mask is None:
        return None
    return key_mask.reshape(batch, -1).to(torch.float32).contiguous()


class _PartPredictor(nn.Module):
    """Export-side wrappers (convert3_onnx.py:27-28, convert3_coreml.py:28-29) run layer by layer on the per-layer kernels of
    train_ops.py (no fused engine plan exists for half a model); precision as ``default_precision()`` or ``.precision``."""

    precision: Optional[str] = None

    def _dt(self):
        prec = self.precision or default_precision()
        return torch.float32 if prec == "fp32" else torch.bfloat16


class TransformerEncoderPredictor(_PartPredictor):
    """models/transformer.py:362-370: forward(enc_input [B,Le,106], key_mask [B,1,1,Le]) -> enc_output [B,Le,d]."""

    def __init__(self, encoder):
        super().__init__()
        self.head_num = encoder.head_num
        self.encoder = encoder

    def forward(self, enc_input, key_mask):
        from ..train_ops import _need_cuda, encoder_forward
        _need_cuda(enc_input, "transformer")
        with torch.no_grad():
            return encoder_forward(self.encoder, enc_input, _flat_mask(key_mask, enc_input.shape[0]), self._dt()).float()


class TransformerDecoderPredictor(_PartPredictor):
    """models/transformer.py:385-393: forward(enc_output, decoder_input [B,Ld] int64, key_mask) -> 3 softmaxes [B,Ld,m_i]."""

    def __init__(self, decoder):
        super().__init__()
        self.head_num = decoder.head_num
        self.decoder = decoder

    def forward(self, enc_output, decoder_input, key_mask):
        from ..train_ops import _need_cuda, decoder_forward
        _need_cuda(enc_output, "transformer")
        with torch.no_grad():
            outs = decoder_forward(self.decoder, decoder_input, enc_output.to(self._dt()).contiguous(),
                                   _flat_mask(key_mask, enc_output.shape[0]), self._dt())
        return [torch.softmax(o, dim=-1) for o in outs]


class TransformerDecoderPredictorSplited(TransformerDecoderPredictor):
    """models/transformer.py:395-404: the decoder input arrives as its three residues (x mod 1091, 1093, 1097)."""

    def forward(self, enc_output, decoder_input1, decoder_input2, decoder_input3, key_mask):
        return super().forward(enc_output, [decoder_input1, decoder_input2, decoder_input3], key_mask)
