"""Feature-sequence chunking and batched decode of a page (reference: process_ocr_base.py:182-283).

``call_OCR`` walks the page's feature rows in sliding windows of at most ``max_encoderlen - 3`` rows (fewer when the window
holds spaces or a ruby group), cuts windows at horizontal / vertical changes, at double newlines and between a ruby base and its
ruby text, and re-reads the tail of each window in the next one (``keep_back`` characters of the next prediction are dropped).
It calls the transformer once per window, at batch 1.  None of the window decisions depends on a prediction, so the whole page
can be planned first (``plan_chunks``), decoded in ONE batched predictor call whose sequences stop independently
(``OCR_b200_Processer.call_transformer_batch`` -> ``ftc_transformer_predict_each``) and assembled afterwards
(``assemble_text``) -- same windows, same text, one launch sequence instead of one per window.
Host-side index arithmetic only (numpy); the arithmetic of the hot path stays in the CUDA library.
"""
from __future__ import annotations

from typing import List, NamedTuple, Sequence, Tuple

import numpy as np

from . import arch

PAD, SOT, EOT = 0, 1, 2      # const.py:12-14
# columns of a feature row after the 100 glyph features (process_ocr_base.py:121-126, 166)
VERTICAL, RUBYBASE, RUBY, SPACE, EMPHASIS, NEWLINE = 100, 101, 102, 103, 104, 105


class Chunk(NamedTuple):
    prev_j: int       # end of the previous window (start of the text this window contributes)
    cur_i: int        # first feature row of the window
    cur_j: int        # one past the last feature row
    keep_back: int    # leading characters of the window's prediction that repeat the previous window


def sp_token() -> np.ndarray:
    t = np.zeros(arch.ENCODER_DIM, dtype=np.float32)
    t[0:arch.FEATURE_DIM:2] = 5
    t[1:arch.FEATURE_DIM:2] = -5
    return t


def _window_end(f: np.ndarray, cur_i: int, max_len: int) -> int:
    """End of the window that starts at row cur_i (process_ocr_base.py:193-230)."""
    n = f.shape[0]
    extra, state = 0, 0                       # rows reserved for spaces (1 each) and ruby groups (3 each)
    for k in range(cur_i, min(cur_i + max_len - 3, n)):
        if f[k, SPACE] > 0:
            extra += 1
        if state == 0 and f[k, RUBYBASE] > 0:
            extra += 3
            state = 1
        elif state == 1 and f[k, RUBY] > 0:
            state = 2
        elif state == 2 and f[k, RUBY] == 0:
            state = 0
    cur_j = min(n, cur_i + (max_len - 3 - extra))
    for j in range(cur_i + 1, cur_j):         # horizontal / vertical change
        if f[j, VERTICAL] != f[cur_i, VERTICAL]:
            cur_j = j
            break
    if cur_j < n - 1 and cur_i + 1 < cur_j - 1:   # double newline: a new block starts
        for j in range(cur_i + 1, cur_j - 1):
            if f[j, NEWLINE] > 0 and f[j + 1, NEWLINE] > 0:
                cur_j = j + 2
                break
    if cur_j < n and cur_j > 1 and f[cur_j - 1, NEWLINE] == 0:   # do not cut inside a ruby base / ruby text group
        for j in reversed(range(cur_i + 1, cur_j)):
            if f[j, RUBY] == 0 and f[j, RUBYBASE] == 0:
                cur_j = j + 1
                break
    return cur_j


def _next_start(f: np.ndarray, cur_i: int, cur_j: int) -> Tuple[int, int]:
    """(start row of the next window, its keep_back) after a window [cur_i, cur_j) that did not reach the end (:255-281)."""
    k = cur_j - 1
    keep_back = 0
    while cur_i < k:
        if f[k, VERTICAL] != f[cur_j, VERTICAL]:
            k += 1
            break
        if f[k, RUBYBASE] > 0 or f[k, RUBY] > 0:
            k += 1
            break
        if k < cur_j - 1 and f[k, NEWLINE] > 0:
            k += 1
            break
        if f[k, SPACE] > 0:
            keep_back += 1
        if k > cur_j - 3:
            k -= 1
        else:
            break
    if cur_i < k:
        return k, keep_back + cur_j - k
    return cur_j, 0


def plan_chunks(features: np.ndarray, max_encoderlen: int = arch.MAX_ENCODERLEN) -> List[Chunk]:
    """Every transformer window of a page, in the reference's order, without decoding anything."""
    f = np.asarray(features, dtype=np.float32)
    n = f.shape[0]
    chunks: List[Chunk] = []
    cur_i = prev_j = keep_back = 0
    while cur_i < n:
        cur_j = _window_end(f, cur_i, max_encoderlen)
        if prev_j == cur_j:                    # the window adds nothing new: restart right behind it (:232-235)
            keep_back = 0
            cur_i = cur_j
            continue
        chunks.append(Chunk(prev_j, cur_i, cur_j, keep_back))
        if cur_j >= n:
            break
        prev_j = cur_j
        cur_i, keep_back = _next_start(f, cur_i, cur_j)
    return chunks


def chunk_inputs(features: np.ndarray, chunks: Sequence[Chunk], max_encoderlen: int = arch.MAX_ENCODERLEN) -> np.ndarray:
    """float32 [len(chunks), max_encoderlen, 106]: SP token, the window's rows, -SP token, zero padding (:237-240)."""
    f = np.asarray(features, dtype=np.float32)
    sp = sp_token()
    out = np.zeros((len(chunks), max_encoderlen, arch.ENCODER_DIM), dtype=np.float32)
    for n, c in enumerate(chunks):
        rows = c.cur_j - c.cur_i
        out[n, 0] = sp
        out[n, 1:1 + rows] = f[c.cur_i:c.cur_j]
        out[n, 1 + rows] = -sp
    return out


def codes_to_text(pred: Sequence[int]) -> str:
    """One window's code points -> text (:243-254): skip SOT, stop at PAD / EOT, U+FFFD for surrogates and out-of-range codes."""
    s = []
    for p in pred:
        p = int(p)
        if p == SOT:
            continue
        if p == PAD or p == EOT:
            break
        if 0xD800 <= p <= 0xDFFF:
            s.append("�")
        elif p < 0x3FFFF:
            s.append(chr(p))
        else:
            s.append("�")
    return "".join(s)


def assemble_text(chunks: Sequence[Chunk], preds: np.ndarray):
    """(result_txt, linebuf) of call_OCR (:256-258) from the per-window predictions int64 [len(chunks), max_decoderlen]."""
    result_txt, linebuf = "", []
    for c, pred in zip(chunks, preds):
        s = codes_to_text(pred)[c.keep_back:]
        result_txt += s
        linebuf.append((c.prev_j, c.cur_j, s))
    return result_txt, linebuf
