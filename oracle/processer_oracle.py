"""TEST INFRASTRUCTURE (never imported by the product): CPU restatement of the train1 input pipeline of the reference,
``dataset/processer.pyx``: augmentation parameters, affine crop (bilinear / nearest), label-map
rasterisation and the colour compositing functions.

Pinned to the UNMODIFIED reference: ``oracle/ref_processer/build_ref.py`` compiles ``/root/reference/dataset/processer.pyx`` where
it lies (Cython -> C++, no FMA contraction) and ``tests/test_processer_oracle.py`` runs both on the same inputs with the same
libc ``srand`` seed (the reference draws every parameter from ``rand()``; this file draws through the same libc in the same
order) -> bit-identical images / maps / id maps; committed goldens: ``tests/golden/processer_*.npz`` (``oracle/make_golden_processer.py``).

Everything is float32 arithmetic rounded operation by operation, written with numpy float32 arrays / scalars (numpy never fuses
a multiply-add).  Where the Cython source mixes in a C double literal (``-0.5 * ax``, ``rx + 0.5``, ``logf(..) + 3.0``) the same
promotion is made here.  libm ``expf / logf / cosf / sinf`` are evaluated in float64 and rounded (equal to glibc's float results
except for vanishingly rare double-rounding cases; the pin test would show one).

Reference lines (dataset/processer.pyx): random_uniform / random_gaussian :22-38, gkern / gaussian_kernel :40-60, matrix_dot /
vector_dot / GetMatrix :62-122, inverse_partial :124-135, center_map :137-163, box_map :165-186, id_map :188-206, getpixel :208-211,
transform_crop :260-454, process :655-673, random_background :675-743, random_mono :745-764, random_single :766-807,
random_double :809-887.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math

import numpy as np

F = np.float32
WIDTH, HEIGHT, SCALE = 768, 768, 4          # util_func.py:6-8

# the parameter stage calls the C library's float functions exactly as the compiled reference does (cosf / sinf / logf are not
# correctly rounded, so "float64 then round" differs from them by an ulp now and then; seed 42 of the live pin test showed it)
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("cosf", "sinf", "logf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]


def _cosf(x):
    return F(_libm.cosf(float(x)))


def _sinf(x):
    return F(_libm.sinf(float(x)))


def _logf(x):
    return F(_libm.logf(float(x)))
W4, H4 = WIDTH // SCALE, HEIGHT // SCALE
RAND_MAX = 2147483647


# ---------------------------------------------------------------------------------------------------------------------
# random numbers: the reference uses libc rand(); srand(seed) through ctypes seeds the same generator
# ---------------------------------------------------------------------------------------------------------------------
class LibcRand:
    """rand() of the process's libc (the generator the compiled reference uses)."""

    def __init__(self, seed=None):
        self.libc = ctypes.CDLL("libc.so.6")
        self.libc.rand.restype = ctypes.c_int
        if seed is not None:
            self.libc.srand(ctypes.c_uint(seed))

    def __call__(self) -> int:
        return int(self.libc.rand())


class ListRand:
    """Replays recorded rand() values (golden fixtures carry the draws, the GPU box needs no particular libc)."""

    def __init__(self, values):
        self.values, self.pos = [int(v) for v in values], 0

    def __call__(self) -> int:
        v = self.values[self.pos]
        self.pos += 1
        return v


class RecordingRand:
    def __init__(self, inner):
        self.inner, self.values = inner, []

    def __call__(self) -> int:
        v = self.inner()
        self.values.append(v)
        return v


def random_uniform(rand) -> np.float32:                       # :22-25
    return F(F(rand()) / F(RAND_MAX))


def random_gaussian(rand) -> np.float32:                      # :27-38 (polar Box-Muller, first variate only)
    w = F(2.0)
    x1 = x2 = F(0)
    while w >= F(1.0):
        x1 = F(2.0 * float(random_uniform(rand)) - 1.0)
        x2 = F(2.0 * float(random_uniform(rand)) - 1.0)
        w = F(F(x1 * x1) + F(x2 * x2))
    logw = float(_logf(w))
    w = F(math.pow((-2.0 * logw) / float(w), 0.5))
    return F(x1 * w)


# ---------------------------------------------------------------------------------------------------------------------
# matrices
# ---------------------------------------------------------------------------------------------------------------------
def matrix_dot(a, b):                                         # :62-73
    out = np.zeros(9, F)
    for j in range(3):
        for i in range(3):
            v = F(0)
            for k in range(3):
                v = F(v + F(a[j * 3 + k] * b[k * 3 + i]))
            out[j * 3 + i] = v
    return out


def vector_dot(a, x1, y1):                                    # :75-86; x1, y1 float32 scalars or arrays
    x1 = np.asarray(x1, F)
    y1 = np.asarray(y1, F)
    one = F(1)
    rx = ((F(0) + a[0] * x1).astype(F) + (a[1] * y1).astype(F)).astype(F)
    rx = (rx + F(a[2] * one)).astype(F)
    ry = ((F(0) + a[3] * x1).astype(F) + (a[4] * y1).astype(F)).astype(F)
    ry = (ry + F(a[5] * one)).astype(F)
    return rx, ry


def get_matrix(x, y, angle, size_x, size_y, sh_x, sh_y):      # :88-122
    x, y, angle = F(x), F(y), F(angle)
    c, s = _cosf(angle), _sinf(angle)
    shear = np.array([1, sh_y, 0, sh_x, 1, 0, 0, 0, 1], F)
    resize = np.array([size_x, 0, 0, 0, size_y, 0, 0, 0, 1], F)
    move = np.array([1, 0, x, 0, 1, y, 0, 0, 1], F)
    rot = np.array([c, -s, 0, s, c, 0, 0, 0, 1], F)
    back = np.array([1, 0, -x, 0, 1, -y, 0, 0, 1], F)
    r = matrix_dot(shear, resize)
    r = matrix_dot(r, move)
    r = matrix_dot(r, rot)
    r = matrix_dot(r, back)
    return r


# ---------------------------------------------------------------------------------------------------------------------
# parameters of one transform_crop call, drawn in the reference's order
# ---------------------------------------------------------------------------------------------------------------------
def draw_crop_params(rand, im_h, im_w, im_h2, im_w2, position):
    """-> dict(rot, inv, inv2 [9] float32; inverse rect (i, j, h, w); cidx, woffset, hoffset; startx0, starty0; nearest)"""
    position = np.asarray(position, F).reshape(-1, 4)
    n = position.shape[0]
    minsize = F(0)                                            # :283-289
    for i in range(n):
        minsize = F(minsize + max(position[i, 2], position[i, 3]))
    if minsize <= 0:
        minsize = F(10)
    else:
        minsize = F(minsize / F(n))
    rotation_angle = F(np.deg2rad(float(random_gaussian(rand)) * 5.0))     # :292
    size_x = F(1.0 * float(random_gaussian(rand)) + 1.0)
    aspect_ratio = F(float(abs(random_gaussian(rand))) + 1.0)
    sh_x = F(float(random_gaussian(rand)) * 0.01)
    sh_y = F(float(random_gaussian(rand)) * 0.01)
    if float(size_x) < 0.8:
        size_x = F(0.8 - float(size_x) + 0.8)
    if float(size_x) < 1.0 and F(size_x * minsize) < 10:
        size_x = F(10.0 / float(minsize))
        aspect_ratio = F(1)
    if float(random_uniform(rand)) < 0.5:
        size_y = F(size_x * aspect_ratio)
    else:
        size_y = F(size_x / aspect_ratio)
    rot = get_matrix(im_w // 2, im_h // 2, rotation_angle, size_x, size_y, sh_x, sh_y)       # :308-309 (C integer division)
    rot2 = get_matrix(im_w2 // 2, im_h2 // 2, rotation_angle, size_x, size_y, sh_x, sh_y)
    inv = np.linalg.inv(rot.reshape(3, 3)).astype(F).reshape(9)                              # :314-326 (float32 LAPACK)
    inv2 = np.linalg.inv(rot2.reshape(3, 3)).astype(F).reshape(9)
    # inverse_partial :124-128
    h = int(F(random_uniform(rand) * F(im_h - 1)))
    w = int(F(random_uniform(rand) * F(im_w - 1)))
    i = int(F(random_uniform(rand) * F(im_h - h + 1)))
    j = int(F(random_uniform(rand) * F(im_w - w + 1)))
    p = {"rot": rot, "inv": inv, "inv2": inv2, "inv_rect": (i, j, h, w), "cidx": -1, "woffset": F(0), "hoffset": F(0),
         "startx0": F(0), "starty0": F(0)}
    if n > 0:                                                 # :358-366
        p["cidx"] = int(F(random_uniform(rand) * F(n)))
        p["woffset"] = F(float(F(random_uniform(rand) * F(WIDTH))) * 0.75 + float(F(WIDTH)) / 8.0)
        p["hoffset"] = F(float(F(random_uniform(rand) * F(HEIGHT))) * 0.75 + float(F(HEIGHT)) / 8.0)
    else:
        p["startx0"] = F(random_uniform(rand) * F(WIDTH))
        p["starty0"] = F(random_uniform(rand) * F(HEIGHT))
    p["nearest"] = bool(float(random_uniform(rand)) < 0.05)   # :389
    return p


# ---------------------------------------------------------------------------------------------------------------------
# label rasterisation
# ---------------------------------------------------------------------------------------------------------------------
def _expf(x32):
    return np.exp(np.asarray(x32, F).astype(np.float64)).astype(F)


def gkern(l, sig):                                            # :40-47 (the exponent is formed in double: "-0.5 * ax")
    i = np.arange(l, dtype=np.float64)
    ax = (i.astype(F).astype(np.float64) - float(F(l - 1)) / 2.0).astype(F)
    arg = ((-0.5 * ax.astype(np.float64)) * ax.astype(np.float64)) / float(F(F(sig) * F(sig)))
    return _expf(arg.astype(F))


def center_map(cx, cy, w, h, center):                         # :137-163; center: float32 [H4, W4], in place
    cx, cy, w, h = F(cx / F(SCALE)), F(cy / F(SCALE)), F(w / F(SCALE)), F(h / F(SCALE))
    fix_w = F(max(float(F(w / F(2))), 1.0))
    fix_h = F(max(float(F(h / F(2))), 1.0))
    kernel_size = int(max(float(fix_w) * 1.5, float(fix_h) * 1.5))
    std_x, std_y = F(fix_w / F(4)), F(fix_h / F(4))
    L = kernel_size * 2 + 1
    gx, gy = gkern(L, std_x), gkern(L, std_y)
    k2d = (gy[:, None] * gx[None, :]).astype(F)
    xi, yi = int(_roundf(cx)), int(_roundf(cy))
    y0, x0 = yi - kernel_size, xi - kernel_size
    ya, yb = max(y0, 0), min(y0 + L, H4)
    xa, xb = max(x0, 0), min(x0 + L, W4)
    if ya >= yb or xa >= xb:
        return
    sub = k2d[ya - y0:yb - y0, xa - x0:xb - x0]
    np.maximum(center[ya:yb, xa:xb], sub, out=center[ya:yb, xa:xb])


def _roundf(v):
    v = float(v)
    return math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5)        # half away from zero


def _ellipse(cx, cy, w, h):
    """pixel window and membership mask shared by box_map :165-186 and id_map :188-206"""
    fix_w = F(max(float(F(w / F(10))), float(SCALE)))
    fix_h = F(max(float(F(h / F(10))), float(SCALE)))
    xmin = max(0, int(F(F(cx - fix_w) / F(SCALE))) - 2)
    xmax = min(W4, int(F(F(cx + fix_w) / F(SCALE))) + 2)
    ymin = max(0, int(F(F(cy - fix_h) / F(SCALE))) - 2)
    ymax = min(H4, int(F(F(cy + fix_h) / F(SCALE))) + 2)
    if xmin >= xmax or ymin >= ymax:
        return None
    xs = (np.arange(xmin, xmax) * SCALE).astype(F) - F(cx)
    ys = (np.arange(ymin, ymax) * SCALE).astype(F) - F(cy)
    qx = (xs / fix_w).astype(F)
    qy = (ys / fix_h).astype(F)
    inside = ((qy * qy).astype(F)[:, None] + (qx * qx).astype(F)[None, :]).astype(F) < F(1)
    return xmin, xmax, ymin, ymax, inside


def box_map(cx, cy, w, h, boxmap):                            # boxmap float32 [2, H4, W4] initialised to +inf
    cx, cy, w, h = F(cx), F(cy), F(w), F(h)
    with np.errstate(invalid="ignore", divide="ignore"):
        sizex = F(float(F(np.log(np.float64(F(w / F(1024)))))) + 3.0)
        sizey = F(float(F(np.log(np.float64(F(h / F(1024)))))) + 3.0)
    e = _ellipse(cx, cy, w, h)
    if e is None:
        return
    xmin, xmax, ymin, ymax, inside = e
    for ch, v in ((0, sizex), (1, sizey)):
        win = boxmap[ch, ymin:ymax, xmin:xmax]
        # Cython min(sizex, cur) = cur if cur < sizex else sizex
        win[inside] = np.where(win[inside] < v, win[inside], v)


def id_map(cx, cy, w, h, code1, code2, indexmap):             # indexmap int32 [2, H4, W4] initialised to 0
    e = _ellipse(F(cx), F(cy), F(w), F(h))
    if e is None:
        return
    xmin, xmax, ymin, ymax, inside = e
    for ch, v in ((0, code1), (1, code2)):
        win = indexmap[ch, ymin:ymax, xmin:xmax]
        win[inside] = np.maximum(win[inside], np.int32(v))


# ---------------------------------------------------------------------------------------------------------------------
# affine warps
# ---------------------------------------------------------------------------------------------------------------------
def _getpixel(img, x, y):                                     # :208-211 (arrays of int coordinates)
    im_h, im_w = img.shape
    ok = (x >= 0) & (x < im_w) & (y >= 0) & (y < im_h)
    v = img[np.clip(y, 0, im_h - 1), np.clip(x, 0, im_w - 1)].astype(F) / F(255)
    return np.where(ok, v, F(0)).astype(F)


def _trunc(a):
    return np.trunc(a).astype(np.int64)                       # C float -> int conversion


def _bilinear(img, rx, ry):
    dx = (rx - np.floor(rx)).astype(F)
    dy = (ry - np.floor(ry)).astype(F)
    dxd, dyd = dx.astype(np.float64), dy.astype(np.float64)   # "(1 - dx)" is emitted as the double expression (1.0 - dx)
    w11 = ((1.0 - dxd) * (1.0 - dyd)).astype(F)
    w21 = (dxd * (1.0 - dyd)).astype(F)
    w12 = ((1.0 - dxd) * dyd).astype(F)
    w22 = (dx * dy).astype(F)
    ix, iy = _trunc(rx), _trunc(ry)
    out = (w11 * _getpixel(img, ix, iy)).astype(F)
    out = (out + (w21 * _getpixel(img, ix + 1, iy)).astype(F)).astype(F)
    out = (out + (w12 * _getpixel(img, ix, iy + 1)).astype(F)).astype(F)
    out = (out + (w22 * _getpixel(img, ix + 1, iy + 1)).astype(F)).astype(F)
    return out


def rotate_positions(position, rot):                          # :345-355
    pos = np.asarray(position, F).reshape(-1, 4)
    half_w, half_h = (pos[:, 2] / F(2)).astype(F), (pos[:, 3] / F(2)).astype(F)
    x1, y1 = (pos[:, 0] - half_w).astype(F), (pos[:, 1] - half_h).astype(F)
    x2, y2 = (pos[:, 0] + half_w).astype(F), (pos[:, 1] + half_h).astype(F)
    xr1, yr1 = vector_dot(rot, x1, y1)
    xr2, yr2 = vector_dot(rot, x2, y2)
    out = np.empty_like(pos)
    out[:, 0] = ((xr1 + xr2).astype(F) / F(2)).astype(F)
    out[:, 1] = ((yr1 + yr2).astype(F) / F(2)).astype(F)
    out[:, 2] = (xr2 - xr1).astype(F)
    out[:, 3] = (yr2 - yr1).astype(F)
    return out


def transform_crop(image, textline, sepline, position, codelist, p):
    """dataset/processer.pyx:260-454 with the random draws replaced by the parameter dict p (draw_crop_params).
    -> outimage float32 [768,768], mapimage float32 [5,192,192], indexmap int32 [2,192,192], minsize float32"""
    image = np.array(image, np.uint8, copy=True)
    textline = np.asarray(textline, np.uint8)
    sepline = np.asarray(sepline, np.uint8)
    codelist = np.asarray(codelist, np.int32).reshape(-1, 2)
    i, j, h, w = p["inv_rect"]
    image[i:i + h, j:j + w] = 255 - image[i:i + h, j:j + w]   # inverse_partial :129-135
    pos = rotate_positions(position, p["rot"])
    n = pos.shape[0]
    if n > 0:
        startx = F(pos[p["cidx"], 0] - p["woffset"])
        starty = F(pos[p["cidx"], 1] - p["hoffset"])
    else:
        startx, starty = F(p["startx0"]), F(p["starty0"])
    center = np.zeros((H4, W4), F)
    boxmap = np.full((2, H4, W4), np.inf, F)
    indexmap = np.zeros((2, H4, W4), np.int32)
    minsize = F(0)
    for k in range(n):                                        # :371-385
        cx, cy = F(pos[k, 0] - startx), F(pos[k, 1] - starty)
        bw, bh = pos[k, 2], pos[k, 3]
        if cx > 0 and cx < WIDTH and cy > 0 and cy < HEIGHT:
            center_map(cx, cy, bw, bh, center)
            box_map(cx, cy, bw, bh, boxmap)
            id_map(cx, cy, bw, bh, codelist[k, 0], codelist[k, 1], indexmap)
            m = max(bw, bh)
            minsize = m if minsize <= 0 else min(minsize, m)
    xs = (np.arange(WIDTH, dtype=np.int64).astype(F) + startx).astype(F)
    ys = (np.arange(HEIGHT, dtype=np.int64).astype(F) + starty).astype(F)
    rx, ry = vector_dot(p["inv"], xs[None, :], ys[:, None])
    if p["nearest"]:                                          # :390-394 ("rx + 0.5" is a double addition)
        ix = np.trunc(rx.astype(np.float64) + 0.5).astype(np.int64)
        iy = np.trunc(ry.astype(np.float64) + 0.5).astype(np.int64)
        outimage = _getpixel(image, ix, iy)
    else:
        outimage = _bilinear(image, rx, ry)
    half = F(SCALE // 2)
    xs2 = ((np.arange(W4).astype(F) * half).astype(F) + F(startx / F(2))).astype(F)
    ys2 = ((np.arange(H4).astype(F) * half).astype(F) + F(starty / F(2))).astype(F)
    rx2, ry2 = vector_dot(p["inv2"], xs2[None, :], ys2[:, None])
    mapimage = np.empty((5, H4, W4), F)
    mapimage[0] = center
    mapimage[1] = np.where(np.isfinite(boxmap[0]), boxmap[0], F(0))
    mapimage[2] = np.where(np.isfinite(boxmap[1]), boxmap[1], F(0))
    mapimage[3] = _bilinear(textline, rx2, ry2)
    mapimage[4] = _bilinear(sepline, rx2, ry2)
    return outimage, mapimage, indexmap, F(minsize)


def process(sample, rand):
    """dataset/processer.pyx:655-673: 1 % blank samples, else transform_crop with freshly drawn parameters.
    -> (outimage, mapimage, indexmap, minsize, params or None)"""
    if float(random_uniform(rand)) < 0.01:
        return np.zeros((HEIGHT, WIDTH), F), np.zeros((5, H4, W4), F), np.zeros((2, H4, W4), np.int32), F(0), None
    image, textline, sepline, position, codelist = sample
    p = draw_crop_params(rand, image.shape[0], image.shape[1], textline.shape[0], textline.shape[1], position)
    return (*transform_crop(image, textline, sepline, position, codelist, p), p)


# ---------------------------------------------------------------------------------------------------------------------
# colour compositing (dataset/processer.pyx:675-887): parameters drawn first, then out = a * fg + (1 - a) * bg per channel
# ---------------------------------------------------------------------------------------------------------------------
def _contrast_bg(fg, u):
    """bg from fg (:752-757 and the per-channel copies): fg > 0.5 -> u * (fg - 0.5) else 1 - u * (1 - (fg + 0.5))"""
    fg, u = F(fg), F(u)
    hi, lo = F(float(fg) + 0.5), F(float(fg) - 0.5)
    if float(fg) > 0.5:
        return F(u * lo)
    return F(1.0 - float(u) * (1.0 - float(hi)))              # double expression (1.0 - (u * (1.0 - hi)))


def draw_mono(rand):                                          # :748-757
    fg = random_uniform(rand)
    u = random_uniform(rand)
    bg = _contrast_bg(fg, u)
    return {"mode": "mono", "fg1": np.array([fg, fg, fg], F), "fg2": np.array([fg, fg, fg], F), "bg": np.array([bg, bg, bg], F),
            "rect": (0, 0, 0, 0)}


def draw_single(rand):                                        # :769-796
    fg = [random_uniform(rand) for _ in range(3)]
    bg = []
    for c in range(3):
        bg.append(_contrast_bg(fg[c], random_uniform(rand)))
    return {"mode": "single", "fg1": np.array(fg, F), "fg2": np.array(fg, F), "bg": np.array(bg, F), "rect": (0, 0, 0, 0)}


def draw_double(rand):                                        # :812-871
    fg1 = [random_uniform(rand) for _ in range(3)]
    fg2 = [random_uniform(rand) for _ in range(3)]
    for c in range(3):
        if float(fg1[c]) > 0.5:
            fg2[c] = F(float(fg2[c]) * 0.5 + 0.5)
        else:
            fg2[c] = F(float(fg2[c]) * 0.5)
    hi = [F(float(max(fg1[c], fg2[c])) + 0.5) for c in range(3)]
    lo = [F(float(min(fg1[c], fg2[c])) - 0.5) for c in range(3)]
    bgu = [random_uniform(rand) for _ in range(3)]
    bg = []
    for c in range(3):
        if float(fg1[c]) > 0.5:
            bg.append(F(bgu[c] * lo[c]))
        else:
            bg.append(F(1.0 - float(bgu[c]) * (1.0 - float(hi[c]))))
    top = int(F(random_uniform(rand) * F(HEIGHT - 1)))
    bottom = int(F(random_uniform(rand) * F(HEIGHT - top))) + top
    left = int(F(random_uniform(rand) * F(WIDTH - 1)))
    right = int(F(random_uniform(rand) * F(WIDTH - left))) + left
    return {"mode": "double", "fg1": np.array(fg1, F), "fg2": np.array(fg2, F), "bg": np.array(bg, F),
            "rect": (top, bottom, left, right)}


def draw_background(rand, bgimg):                             # :675-731
    bgimg = np.asarray(bgimg, np.uint8)
    bh, bw = bgimg.shape[:2]
    startx = int(F(random_uniform(rand) * F(bw - WIDTH))) if bw > WIDTH else 0
    starty = int(F(random_uniform(rand) * F(bh - HEIGHT))) if bh > HEIGHT else 0
    crop = crop_background(bgimg, startx, starty)
    fg = []
    for c in range(3):
        bgm = F(np.mean(crop[c]))
        hi, lo = F(float(bgm) + 0.5), F(float(bgm) - 0.5)
        u = random_uniform(rand)
        if float(bgm) > 0.5:
            fg.append(F(u * lo))
        else:
            fg.append(F(1.0 - float(u) * (1.0 - float(hi))))
    return {"mode": "background", "fg1": np.array(fg, F), "fg2": np.array(fg, F), "bg": np.zeros(3, F), "rect": (0, 0, 0, 0),
            "bg_start": (startx, starty)}


def crop_background(bgimg, startx, starty):                   # :690-703
    bh, bw = bgimg.shape[:2]
    crop = np.zeros((3, HEIGHT, WIDTH), F)
    ye, xe = min(HEIGHT, bh - starty), min(WIDTH, bw - startx)
    if ye > 0 and xe > 0:
        crop[:, :ye, :xe] = (bgimg[starty:starty + ye, startx:startx + xe, :3].astype(F) / F(255)).transpose(2, 0, 1)
    return crop


def composite(a, cp, bgimg=None):
    """a float32 [768,768] (alpha = ink) -> float32 [3,768,768]; cp from draw_mono / single / double / background"""
    a = np.asarray(a, F)
    na = 1.0 - a.astype(np.float64)                           # "(1 - a)" is the double expression (1.0 - a)

    def blend(fg, bg):                                        # (a * fg) is a float product, the rest double, result -> float
        return (a * F(fg)).astype(F).astype(np.float64) + na * np.asarray(bg, F).astype(np.float64)

    out = np.empty((3, HEIGHT, WIDTH), F)
    if cp["mode"] == "background":
        crop = crop_background(np.asarray(bgimg, np.uint8), *cp["bg_start"])
        for c in range(3):
            out[c] = np.maximum(0.0, np.minimum(1.0, blend(cp["fg1"][c], crop[c]))).astype(F)    # :739-741
        return out
    top, bottom, left, right = cp["rect"]
    yy, xx = np.mgrid[0:HEIGHT, 0:WIDTH]
    inner = (xx > left) & (xx < right) & (yy > top) & (yy < bottom) if cp["mode"] == "double" else np.zeros_like(a, bool)
    for c in range(3):
        v1 = blend(cp["fg1"][c], cp["bg"][c]).astype(F)
        v2 = blend(cp["fg2"][c], cp["bg"][c]).astype(F)
        out[c] = np.where(inner, v2, v1)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# random_distortion (dataset/data_detector.py:28-42): numpy Generator decisions, scipy.ndimage.gaussian_filter pixels
# ---------------------------------------------------------------------------------------------------------------------
def draw_distortion(rng, s, shape=None):
    """The decisions of random_distortion in its draw order.  With ``shape`` the noise field is drawn from ``rng`` exactly where the
    reference draws it (pin test: the generator streams stay aligned); without it the noise is left to the caller."""
    d = {"noise_on": False, "alpha": 0.0, "mode": 0, "sigma": 0.0, "unsharp_k": F(0), "noise": None}
    if rng.random() < 0.3:
        d["noise_on"] = True
        d["alpha"] = min(0.4 * rng.random(), 20 / max(1, s))
        if shape is not None:
            d["noise"] = rng.normal(size=shape)
    if rng.random() < 0.3:
        d["mode"], d["sigma"] = 1, min(s / 8, 1.5 * rng.random())
    elif rng.random() < 0.3:
        d["mode"], d["sigma"], d["unsharp_k"] = 2, 5.0, 10. * rng.random()
    return d


def random_distortion(im, d, noise=None):
    """im float32 [3,768,768] -> float32 [3,768,768]; d from draw_distortion; noise: float64 standard normals (else d['noise'])"""
    from scipy.ndimage import gaussian_filter
    im = np.array(im, F, copy=True)
    if d["noise_on"]:
        z = noise if noise is not None else d["noise"]
        im += d["alpha"] * z
        im = np.clip(im, 0, 1)
    if d["mode"] == 1:
        im = gaussian_filter(im, sigma=d["sigma"])
        im = np.clip(im, 0, 1)
    elif d["mode"] == 2:
        blurred = gaussian_filter(im, sigma=5.)
        im = im + float(d["unsharp_k"]) * (im - blurred)
        im = np.clip(im, 0, 1)
    return im
