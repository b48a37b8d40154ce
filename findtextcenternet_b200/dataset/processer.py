"""Device version of the reference's train1 input pipeline.

Reference: ``dataset/processer.pyx`` (``process`` :655-673 -> ``transform_crop`` :260-454; ``random_background`` / ``random_mono`` /
``random_single`` / ``random_double`` :675-887) and ``dataset/data_detector.py`` (``random_salt`` :17-26, ``transforms3`` :44-59).
There every sample goes through one Python / Cython call on a DataLoader worker; here a whole batch is four kernel launches
(``ftc_crop_batch``, csrc/data_ops.cu) and the result is born in HBM in the layout the train step reads:
image float32 [B,3,768,768] in [0,1], labelmap float32 [B,5,192,192], idmap [B,2,192,192].

Split of the work: every RANDOM DECISION stays on the host, drawn in the reference's order from the same generator the
reference uses (libc ``rand()``; ``srand(seed)`` therefore reproduces the reference's augmentation stream draw for draw), a few
dozen scalars per sample; every PIXEL is computed on the device.  The same parameters give the reference's arrays bit for bit
(tests/test_processer.py).  ``random_distortion`` (data_detector.py:28-42) is ``ftc_distort_batch``: its decisions are drawn from the
numpy Generator in the reference's order, blur / unsharp mask reproduce scipy's gaussian_filter (all three axes, reflect, double
accumulation in correlate1d's order); only the NOISE FIELD differs in kind: 1.8 M normals per image come from a counter-based
generator on the device instead of numpy's stream (same distribution, not the same numbers).

No CPU fallback: without the CUDA library / a CUDA device the calls raise.
"""
from __future__ import annotations

import ctypes as C
import ctypes.util
import math
from typing import Callable, Optional, Sequence

import numpy as np

F = np.float32
WIDTH, HEIGHT, SCALE = 768, 768, 4
RAND_MAX = 2147483647


# cosf / sinf / logf of the C library, the calls the reference's parameter stage makes (float results are not correctly rounded:
# going through float64 would differ by an ulp now and then, and an ulp in the affine matrix moves every pixel)
_libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("cosf", "sinf", "logf"):
    getattr(_libm, _n).restype = C.c_float
    getattr(_libm, _n).argtypes = [C.c_float]


class CropSample(C.Structure):
    """ftc_crop_sample of include/ftc_b200.h"""
    _fields_ = [
        ("image", C.c_void_p), ("textline", C.c_void_p), ("sepline", C.c_void_p), ("bgimg", C.c_void_p), ("salt", C.c_void_p),
        ("im_h", C.c_int), ("im_w", C.c_int), ("im_h2", C.c_int), ("im_w2", C.c_int),
        ("box_begin", C.c_int), ("box_count", C.c_int),
        ("rot", C.c_float * 9), ("inv", C.c_float * 9), ("inv2", C.c_float * 9),
        ("inv_i", C.c_int), ("inv_j", C.c_int), ("inv_h", C.c_int), ("inv_w", C.c_int),
        ("cidx", C.c_int),
        ("woffset", C.c_float), ("hoffset", C.c_float), ("startx0", C.c_float), ("starty0", C.c_float),
        ("nearest", C.c_int), ("blank", C.c_int), ("color_mode", C.c_int),
        ("fg1", C.c_float * 3), ("fg2", C.c_float * 3), ("bg", C.c_float * 3),
        ("rect_top", C.c_int), ("rect_bottom", C.c_int), ("rect_left", C.c_int), ("rect_right", C.c_int),
        ("bg_h", C.c_int), ("bg_w", C.c_int), ("bg_startx", C.c_int), ("bg_starty", C.c_int),
        ("salt_s", C.c_int), ("salt_h", C.c_int), ("salt_w", C.c_int),
    ]


class DistortSample(C.Structure):
    """ftc_distort_sample of include/ftc_b200.h"""
    _fields_ = [("noise_on", C.c_int), ("mode", C.c_int), ("radius", C.c_int), ("pad_", C.c_int), ("alpha", C.c_double),
                ("noise_seed", C.c_ulonglong), ("unsharp_k", C.c_float), ("pad2_", C.c_float)]


class LibcRand:
    """``rand()`` of the C library -- the generator dataset/processer.pyx draws from (:22-25, seeded at import :887)."""

    def __init__(self, seed: Optional[int] = None):
        self._libc = C.CDLL(None)
        self._libc.rand.restype = C.c_int
        if seed is not None:
            self._libc.srand(C.c_uint(seed))

    def __call__(self) -> int:
        return int(self._libc.rand())


def _uniform(rand) -> np.float32:
    return F(F(rand()) / F(RAND_MAX))


def _gaussian(rand) -> np.float32:
    # polar method, first variate (processer.pyx:27-38); the square root is taken in double
    while True:
        x1 = F(2.0 * float(_uniform(rand)) - 1.0)
        x2 = F(2.0 * float(_uniform(rand)) - 1.0)
        w = F(F(x1 * x1) + F(x2 * x2))
        if w < F(1.0):
            break
    lw = float(F(_libm.logf(float(w))))
    return F(x1 * F(math.pow((-2.0 * lw) / float(w), 0.5)))


def _mat3(a, b):
    out = np.zeros(9, F)
    for j in range(3):
        for i in range(3):
            v = F(0)
            for k in range(3):
                v = F(v + F(a[j * 3 + k] * b[k * 3 + i]))
            out[j * 3 + i] = v
    return out


def _affine(cx, cy, angle, size_x, size_y, sh_x, sh_y):
    """GetMatrix (processer.pyx:88-122): shear . resize . move . rotate . move-back in float32"""
    cx, cy = F(cx), F(cy)
    c, s = F(_libm.cosf(float(angle))), F(_libm.sinf(float(angle)))
    m = _mat3(np.array([1, sh_y, 0, sh_x, 1, 0, 0, 0, 1], F), np.array([size_x, 0, 0, 0, size_y, 0, 0, 0, 1], F))
    m = _mat3(m, np.array([1, 0, cx, 0, 1, cy, 0, 0, 1], F))
    m = _mat3(m, np.array([c, -s, 0, s, c, 0, 0, 0, 1], F))
    return _mat3(m, np.array([1, 0, -cx, 0, 1, -cy, 0, 0, 1], F))


def draw_crop_params(rand: Callable[[], int], im_h: int, im_w: int, im_h2: int, im_w2: int, position) -> dict:
    """The random decisions of one transform_crop call, in the reference's order (processer.pyx:283-366, :389)."""
    position = np.asarray(position, F).reshape(-1, 4)
    n = position.shape[0]
    # running float32 sum in box order (:283-289); ufunc.accumulate is strictly sequential, so the rounding sequence is the loop's
    mean_size = F(np.add.accumulate(np.maximum(position[:, 2], position[:, 3]), dtype=F)[-1]) if n else F(0)
    mean_size = F(10) if mean_size <= 0 else F(mean_size / F(n))
    angle = F(np.deg2rad(float(_gaussian(rand)) * 5.0))
    size_x = F(float(_gaussian(rand)) + 1.0)
    aspect = F(float(abs(_gaussian(rand))) + 1.0)
    sh_x = F(float(_gaussian(rand)) * 0.01)
    sh_y = F(float(_gaussian(rand)) * 0.01)
    if float(size_x) < 0.8:
        size_x = F(0.8 - float(size_x) + 0.8)
    if float(size_x) < 1.0 and F(size_x * mean_size) < 10:
        size_x, aspect = F(10.0 / float(mean_size)), F(1)
    size_y = F(size_x * aspect) if float(_uniform(rand)) < 0.5 else F(size_x / aspect)
    rot = _affine(im_w // 2, im_h // 2, angle, size_x, size_y, sh_x, sh_y)
    rot2 = _affine(im_w2 // 2, im_h2 // 2, angle, size_x, size_y, sh_x, sh_y)
    inv = np.linalg.inv(rot.reshape(3, 3)).astype(F).reshape(9)
    inv2 = np.linalg.inv(rot2.reshape(3, 3)).astype(F).reshape(9)
    h = int(F(_uniform(rand) * F(im_h - 1)))                  # inverse_partial :124-128
    w = int(F(_uniform(rand) * F(im_w - 1)))
    i = int(F(_uniform(rand) * F(im_h - h + 1)))
    j = int(F(_uniform(rand) * F(im_w - w + 1)))
    p = dict(rot=rot, inv=inv, inv2=inv2, inv_rect=(i, j, h, w), cidx=-1, woffset=F(0), hoffset=F(0), startx0=F(0), starty0=F(0))
    if n > 0:
        p["cidx"] = int(F(_uniform(rand) * F(n)))
        p["woffset"] = F(float(F(_uniform(rand) * F(WIDTH))) * 0.75 + WIDTH / 8.0)
        p["hoffset"] = F(float(F(_uniform(rand) * F(HEIGHT))) * 0.75 + HEIGHT / 8.0)
    else:
        p["startx0"] = F(_uniform(rand) * F(WIDTH))
        p["starty0"] = F(_uniform(rand) * F(HEIGHT))
    p["nearest"] = bool(float(_uniform(rand)) < 0.05)
    return p


def draw_process_params(rand, im_h, im_w, im_h2, im_w2, position) -> dict:
    """process() (:655-673): 1 % of the samples are blank, the others get transform_crop parameters"""
    if float(_uniform(rand)) < 0.01:
        return dict(blank=True)
    return draw_crop_params(rand, im_h, im_w, im_h2, im_w2, position)


def _contrast(v, u):
    """the partner colour at least 0.5 away from v (:752-757 and its per-channel copies); double expression rounded once"""
    v, u = F(v), F(u)
    if float(v) > 0.5:
        return F(u * F(float(v) - 0.5))
    return F(1.0 - float(u) * (1.0 - float(F(float(v) + 0.5))))


def draw_mono(rand) -> dict:
    fg = _uniform(rand)
    bg = _contrast(fg, _uniform(rand))
    return dict(mode=1, fg1=[fg] * 3, fg2=[fg] * 3, bg=[bg] * 3, rect=(0, 0, 0, 0))


def draw_single(rand) -> dict:
    fg = [_uniform(rand) for _ in range(3)]
    bg = [_contrast(fg[c], _uniform(rand)) for c in range(3)]
    return dict(mode=1, fg1=fg, fg2=list(fg), bg=bg, rect=(0, 0, 0, 0))


def draw_double(rand) -> dict:
    fg1 = [_uniform(rand) for _ in range(3)]
    fg2 = [_uniform(rand) for _ in range(3)]
    fg2 = [F(float(fg2[c]) * 0.5 + 0.5) if float(fg1[c]) > 0.5 else F(float(fg2[c]) * 0.5) for c in range(3)]
    hi = [F(float(max(fg1[c], fg2[c])) + 0.5) for c in range(3)]
    lo = [F(float(min(fg1[c], fg2[c])) - 0.5) for c in range(3)]
    u = [_uniform(rand) for _ in range(3)]
    bg = [F(u[c] * lo[c]) if float(fg1[c]) > 0.5 else F(1.0 - float(u[c]) * (1.0 - float(hi[c]))) for c in range(3)]
    top = int(F(_uniform(rand) * F(HEIGHT - 1)))
    bottom = int(F(_uniform(rand) * F(HEIGHT - top))) + top
    left = int(F(_uniform(rand) * F(WIDTH - 1)))
    right = int(F(_uniform(rand) * F(WIDTH - left))) + left
    return dict(mode=1, fg1=fg1, fg2=fg2, bg=bg, rect=(top, bottom, left, right))


def draw_background(rand, bgimg: np.ndarray) -> dict:
    """random_background (:675-731): crop origin, then one foreground colour per channel contrasting with the crop's mean"""
    bh, bw = bgimg.shape[:2]
    sx = int(F(_uniform(rand) * F(bw - WIDTH))) if bw > WIDTH else 0
    sy = int(F(_uniform(rand) * F(bh - HEIGHT))) if bh > HEIGHT else 0
    crop = np.zeros((3, HEIGHT, WIDTH), F)
    ye, xe = min(HEIGHT, bh - sy), min(WIDTH, bw - sx)
    crop[:, :ye, :xe] = (bgimg[sy:sy + ye, sx:sx + xe, :3].astype(F) / F(255)).transpose(2, 0, 1)
    fg = []
    for c in range(3):
        m = F(np.mean(crop[c]))
        u = _uniform(rand)
        fg.append(F(u * F(float(m) - 0.5)) if float(m) > 0.5 else F(1.0 - float(u) * (1.0 - float(F(float(m) + 0.5)))))
    return dict(mode=2, fg1=fg, fg2=list(fg), bg=[F(0)] * 3, rect=(0, 0, 0, 0), bg_start=(sx, sy))


def draw_salt(rng: np.random.Generator, minsize: float, prob: float):
    """random_salt (data_detector.py:17-26) as a cell grid: (cell size, uint8 cells with 0 -> ink 0, 1 -> keep, 2 -> ink 1)"""
    s = min(max(1, int(minsize / 4)), int(rng.integers(1, 16)))
    shape = ((HEIGHT + s) // s, (WIDTH + s) // s)
    cells = rng.choice(np.array([0, 1, 2], np.uint8), p=[prob / 2, 1 - prob, prob / 2], size=shape).astype(np.uint8)
    return s, cells


def draw_distortion(rng: np.random.Generator, minsize: float) -> dict:
    """random_distortion's decisions (data_detector.py:28-42) in its draw order; the noise field itself is generated on the device"""
    d = dict(noise_on=False, alpha=0.0, mode=0, sigma=0.0, unsharp_k=0.0, noise_seed=0)
    if rng.random() < 0.3:
        d["noise_on"], d["alpha"] = True, min(0.4 * rng.random(), 20 / max(1, minsize))
        d["noise_seed"] = int(rng.integers(0, 2 ** 63 - 1))
    if rng.random() < 0.3:
        d["mode"], d["sigma"] = 1, min(minsize / 8, 1.5 * rng.random())
    elif rng.random() < 0.3:
        d["mode"], d["sigma"], d["unsharp_k"] = 2, 5.0, 10. * rng.random()
    return d


def gauss_taps(sigma: float, truncate: float = 4.0):
    """(radius, taps w[0..radius]) of scipy.ndimage.gaussian_filter1d: radius int(truncate sigma + 0.5), exp(-x^2 / 2 sigma^2) normalised
    over the full kernel in float64 (scipy's _gaussian_kernel1d); sigma <= 1e-15 is scipy's "skip this axis" = identity."""
    if not sigma > 1e-15:
        return 0, np.ones(1)
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    phi = phi / phi.sum()
    return radius, phi[radius::-1].copy()          # w[j] = tap at distance j (left half, as correlate1d indexes it)


def fill_distort(d: DistortSample, p: dict) -> np.ndarray:
    d.noise_on, d.mode = int(bool(p["noise_on"])), int(p["mode"])
    d.alpha, d.noise_seed = float(p["alpha"]), int(p.get("noise_seed", 0))
    d.unsharp_k = float(np.float32(p["unsharp_k"]))
    radius, taps = gauss_taps(p["sigma"]) if p["mode"] else (0, np.ones(1))
    if radius > 63:
        raise ValueError("random_distortion: blur radius above 63 taps")
    d.radius = radius
    w = np.zeros(64)
    w[:radius + 1] = taps
    return w


def fill_descriptor(d: CropSample, ptrs: dict, shapes: dict, box_begin: int, box_count: int, p: dict, color: Optional[dict] = None,
                    salt: Optional[tuple] = None) -> None:
    """ptrs: addresses of image / textline / sepline (/ bgimg / salt cells); shapes: their (h, w)"""
    d.image, d.textline, d.sepline = ptrs["image"], ptrs["textline"], ptrs["sepline"]
    d.im_h, d.im_w = shapes["image"]
    d.im_h2, d.im_w2 = shapes["textline"]
    d.box_begin, d.box_count = box_begin, box_count
    d.blank = 1 if p.get("blank") else 0
    if not d.blank:
        for k in ("rot", "inv", "inv2"):
            getattr(d, k)[:] = [float(v) for v in p[k]]
        d.inv_i, d.inv_j, d.inv_h, d.inv_w = p["inv_rect"]
        d.cidx = max(int(p["cidx"]), 0)
        d.woffset, d.hoffset, d.startx0, d.starty0 = float(p["woffset"]), float(p["hoffset"]), float(p["startx0"]), float(p["starty0"])
        d.nearest = 1 if p["nearest"] else 0
    d.color_mode = 0
    d.bgimg = None
    if color is not None:
        d.color_mode = color["mode"]
        d.fg1[:] = [float(v) for v in color["fg1"]]
        d.fg2[:] = [float(v) for v in color["fg2"]]
        d.bg[:] = [float(v) for v in color["bg"]]
        d.rect_top, d.rect_bottom, d.rect_left, d.rect_right = color["rect"]
        if color["mode"] == 2:
            d.bgimg = ptrs["bgimg"]
            d.bg_h, d.bg_w = shapes["bgimg"]
            d.bg_startx, d.bg_starty = color["bg_start"]
    d.salt = None
    if salt is not None:
        d.salt = ptrs["salt"]
        d.salt_s = int(salt[0])
        d.salt_h, d.salt_w = shapes["salt"]


class GpuProcesser:
    """Batch version of ``dataset.map(process).map(transforms3)`` (dataset/data_detector.py:96-97).

    samples: sequence of (image uint8 [H,W], textline uint8 [H/2,W/2], sepline uint8 [H/2,W/2], position float32 [n,4],
    codelist int32 [n,2]) exactly as the reference's WebDataset decoder yields them (numpy, host).  ``__call__`` returns device
    tensors (image [B,3,768,768] float32, labelmap [B,5,192,192] float32, idmap [B,2,192,192] int64, minsize [B] float32)."""

    def __init__(self, device="cuda", rand: Optional[Callable[[], int]] = None, rng: Optional[np.random.Generator] = None,
                 backgrounds: Sequence[np.ndarray] = (), distortion: bool = True):
        import torch
        from .. import _lib
        self.torch, self.lib = torch, _lib.load()
        self._check = _lib.check
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GpuProcesser runs on a CUDA device only (no CPU fallback)")
        self.rand = rand if rand is not None else LibcRand()
        self.rng = rng if rng is not None else np.random.default_rng()
        self.backgrounds = list(backgrounds)
        self.distortion = distortion

    # -- parameter drawing: the reference's control flow (process :655-673, transforms3 :44-59) --
    def draw(self, sample) -> tuple:
        image, textline, sepline, position, codelist = sample
        p = draw_process_params(self.rand, image.shape[0], image.shape[1], textline.shape[0], textline.shape[1], position)
        return p

    def draw_color(self, minsize: float):
        """transforms3 (:44-59): optional salt, then background photograph / mono / single / double colouring"""
        rng = self.rng
        salt = None
        if rng.random() < 0.2:
            salt = draw_salt(rng, minsize, 0.2 * rng.random())
        bgimg = None
        if rng.random() < 0.3 and self.backgrounds:
            bgimg = self.backgrounds[int(rng.integers(len(self.backgrounds)))]
            color = draw_background(self.rand, bgimg)
        elif rng.random() < 0.5:
            color = draw_mono(self.rand)
        elif rng.random() < 0.5:
            color = draw_single(self.rand)
        else:
            color = draw_double(self.rand)
        return color, salt, bgimg

    def _arena(self, nbytes: int):
        """Two rotating (pinned host, device) byte arenas: one memcpy per array into pinned memory, ONE H2D copy per batch.  A set is
        reused only after the copy that read its host half has completed (event), and device-side reuse is ordered by the stream."""
        torch = self.torch
        if not hasattr(self, "_arenas"):
            self._arenas, self._turn = [None, None], 0
        self._turn ^= 1
        a = self._arenas[self._turn]
        if a is None or a["host"].numel() < nbytes:
            cap = max(nbytes, 1 << 20) * 5 // 4
            a = {"host": torch.empty(cap, dtype=torch.uint8).pin_memory(), "dev": torch.empty(cap, dtype=torch.uint8, device=self.device),
                 "event": None}
            self._arenas[self._turn] = a
        if a["event"] is not None:
            a["event"].synchronize()
        return a

    def stage(self, samples, params, colors=None, salts=None, bgimgs=None) -> dict:
        """Host -> device: pages, masks, boxes and the descriptor table of one batch through one pinned arena and ONE asynchronous
        H2D copy (the per-array pin_memory() + copy of the first version cost more than the kernels)."""
        torch = self.torch
        B = len(samples)
        counts = [int(np.asarray(s[3]).reshape(-1, 4).shape[0]) for s in samples]
        total = int(sum(counts))
        items = []                      # (key, sample index or -1, contiguous numpy array)

        def add(key, b, a, dtype):
            items.append((key, b, np.ascontiguousarray(a, dtype)))

        add("pos", -1, np.concatenate([np.asarray(s[3], F).reshape(-1, 4) for s in samples]) if total else np.zeros((1, 4), F), F)
        add("code", -1, np.concatenate([np.asarray(s[4], np.int32).reshape(-1, 2) for s in samples]) if total else np.zeros((1, 2), np.int32), np.int32)
        shapes = []
        for b, s in enumerate(samples):
            add("image", b, s[0], np.uint8); add("textline", b, s[1], np.uint8); add("sepline", b, s[2], np.uint8)
            sh = {"image": s[0].shape[:2], "textline": s[1].shape[:2]}
            color = colors[b] if colors is not None else None
            salt = salts[b] if salts is not None else None
            if color is not None and color["mode"] == 2:
                add("bgimg", b, bgimgs[b][:, :, :3], np.uint8)
                sh["bgimg"] = bgimgs[b].shape[:2]
            if salt is not None:
                add("salt", b, salt[1], np.uint8)
                sh["salt"] = salt[1].shape
            shapes.append(sh)
        offs, n = [], 0
        for _, _, a in items:
            offs.append(n)
            n += (a.nbytes + 255) // 256 * 256
        desc_off = n
        n += (C.sizeof(CropSample) * B + 255) // 256 * 256
        arena = self._arena(n)
        host = arena["host"].numpy()
        host_t = arena["host"]
        base = arena["dev"].data_ptr()
        ptrs = [dict() for _ in range(B)]
        glob = {}
        for (key, b, a), o in zip(items, offs):
            flat = a.reshape(-1).view(np.uint8)
            if flat.size >= (1 << 20) and flat.flags.writeable:        # page-sized arrays: torch's threaded copy
                host_t[o:o + flat.size].copy_(torch.from_numpy(flat))
            else:
                host[o:o + flat.size] = flat
            if b < 0:
                glob[key] = base + o
            else:
                ptrs[b][key] = base + o
        desc = (CropSample * B)()
        begin = 0
        for b in range(B):
            fill_descriptor(desc[b], ptrs[b], shapes[b], begin, counts[b], params[b], colors[b] if colors is not None else None,
                            salts[b] if salts is not None else None)
            begin += counts[b]
        raw = np.frombuffer(bytes(desc), np.uint8)
        host[desc_off:desc_off + raw.size] = raw
        cs = torch.cuda.current_stream(self.device)
        arena["dev"][:n].copy_(arena["host"][:n], non_blocking=True)
        arena["event"] = torch.cuda.Event()
        arena["event"].record(cs)
        nbytes = int(self.lib.ftc_crop_scratch_bytes(B, total))
        return dict(batch=B, total=total, desc_ptr=base + desc_off, pos_ptr=glob["pos"], code_ptr=glob["code"], arena=arena,
                    channels=1 if colors is None else 3, scratch=torch.empty(nbytes, dtype=torch.uint8, device=self.device),
                    scratch_bytes=nbytes, h2d_bytes=n)

    def launch(self, st: dict, stream=None, out=None):
        """ftc_crop_batch on a staged batch: four kernel launches, outputs born on the device."""
        torch, dev, B = self.torch, self.device, st["batch"]
        if out is None:
            out = (torch.empty(B, st["channels"], HEIGHT, WIDTH, dtype=torch.float32, device=dev),
                   torch.empty(B, 5, HEIGHT // SCALE, WIDTH // SCALE, dtype=torch.float32, device=dev),
                   torch.empty(B, 2, HEIGHT // SCALE, WIDTH // SCALE, dtype=torch.int32, device=dev),
                   torch.empty(B, dtype=torch.float32, device=dev))
        image, labelmap, idmap, minsize = out
        cs = stream if stream is not None else torch.cuda.current_stream(dev)
        total = st["total"]
        self._check(self.lib.ftc_crop_batch(st["desc_ptr"], B, st["pos_ptr"] if total else None,
                                            st["code_ptr"] if total else None, total, image.data_ptr(), st["channels"],
                                            labelmap.data_ptr(), idmap.data_ptr(), minsize.data_ptr(), st["scratch"].data_ptr(),
                                            st["scratch_bytes"], cs.cuda_stream), "ftc_crop_batch")
        return image, labelmap, idmap, minsize

    def run(self, samples, params, colors=None, salts=None, bgimgs=None, stream=None):
        """stage + launch with explicit parameters (what the parity tests drive).  colors None -> gray [B,1,768,768]."""
        st = self.stage(samples, params, colors, salts, bgimgs)
        if stream is not None:              # the arena copy was enqueued on the current stream
            stream.wait_stream(self.torch.cuda.current_stream(self.device))
        out = self.launch(st, stream)
        if stream is not None:
            st["scratch"].record_stream(stream)
            self.torch.cuda.current_stream(self.device).wait_stream(stream)   # the arena is reused in stream order
        return out

    def distort(self, image, dparams, noise=None, stream=None):
        """random_distortion (data_detector.py:28-42) in place on a device batch [B,3,768,768] float32 (ftc_distort_batch)."""
        torch = self.torch
        B = image.shape[0]
        assert image.is_cuda and image.dtype == torch.float32 and image.is_contiguous() and tuple(image.shape[1:]) == (3, HEIGHT, WIDTH)
        desc = (DistortSample * B)()
        w = np.stack([fill_distort(desc[b], dparams[b]) for b in range(B)])
        if not any(d.noise_on or d.mode for d in desc):
            return image
        desc_dev = torch.from_numpy(np.frombuffer(bytes(desc), np.uint8).copy()).to(self.device)
        w_dev = torch.from_numpy(np.ascontiguousarray(w)).to(self.device)
        nz = None if noise is None else noise.to(self.device, torch.float64).contiguous()
        nbytes = int(self.lib.ftc_distort_scratch_bytes(B))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        cs = stream if stream is not None else torch.cuda.current_stream(self.device)
        self._check(self.lib.ftc_distort_batch(image.data_ptr(), B, desc_dev.data_ptr(), w_dev.data_ptr(), None if nz is None else nz.data_ptr(),
                                               scratch.data_ptr(), nbytes, cs.cuda_stream), "ftc_distort_batch")
        for t in (desc_dev, w_dev, scratch) + (() if nz is None else (nz,)):
            t.record_stream(cs)
        return image

    def __call__(self, samples):
        """process + transforms3 for a batch: parameters drawn per sample in order, pixels on the device.  The colour stage needs
        minsize (salt cell size) only when salt is drawn; it is recomputed on the host from the rotated boxes in that case."""
        params = [self.draw(s) for s in samples]
        colors, salts, bgs = [], [], []
        host_ms = [host_minsize(s[3], p) for s, p in zip(samples, params)]
        for m in host_ms:
            c, salt, bg = self.draw_color(m)
            colors.append(c); salts.append(salt); bgs.append(bg)
        image, labelmap, idmap, minsize = self.run(samples, params, colors, salts, bgs)
        if self.distortion:
            self.distort(image, [draw_distortion(self.rng, m) for m in host_ms])
        return image, labelmap, idmap.long(), minsize


def host_minsize(position, p) -> float:
    """minsize of transform_crop (:371-385) from the parameters alone (a few hundred boxes: host arithmetic)"""
    if p.get("blank"):
        return 0.0
    pos = np.asarray(position, F).reshape(-1, 4)
    if pos.shape[0] == 0:
        return 0.0
    a = p["rot"]

    def vd(x, y):
        rx = ((a[0] * x).astype(F) + (a[1] * y).astype(F)).astype(F) + a[2]
        ry = ((a[3] * x).astype(F) + (a[4] * y).astype(F)).astype(F) + a[5]
        return rx.astype(F), ry.astype(F)

    hw, hh = (pos[:, 2] / F(2)).astype(F), (pos[:, 3] / F(2)).astype(F)
    x1, y1 = vd((pos[:, 0] - hw).astype(F), (pos[:, 1] - hh).astype(F))
    x2, y2 = vd((pos[:, 0] + hw).astype(F), (pos[:, 1] + hh).astype(F))
    cx, cy = ((x1 + x2).astype(F) / F(2)).astype(F), ((y1 + y2).astype(F) / F(2)).astype(F)
    w, h = (x2 - x1).astype(F), (y2 - y1).astype(F)
    sx, sy = F(cx[p["cidx"]] - p["woffset"]), F(cy[p["cidx"]] - p["hoffset"])
    rx, ry = (cx - sx).astype(F), (cy - sy).astype(F)
    m = 0.0
    for k in range(pos.shape[0]):
        if 0 < rx[k] < WIDTH and 0 < ry[k] < HEIGHT:
            v = float(max(w[k], h[k]))
            m = v if m <= 0 else min(m, v)
    return m
