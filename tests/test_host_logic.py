"""CPU: host-side logic, C-ABI surface, world_size-2 gloo sharding."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_loads_and_exports_every_declared_symbol():
    from findtextcenternet_b200 import _lib, build
    build.build()
    lib = _lib.load()
    with open(os.path.join(ROOT, "include", "ftc_b200.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(ftc_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ftc_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.ftc_version() >= 100
    # struct layout agreement between ctypes and the header (no GPU needed: create/destroy are host-only)
    cfg = _lib.make_detector_config("xl", _lib.PREC_BF16, _lib.GEMM_TCGEN05)
    h = C.c_void_p()
    _lib.check(lib.ftc_detector_create(C.byref(cfg), C.byref(h)))
    wb = lib.ftc_detector_weight_bytes(h)
    assert 4.0e8 < wb < 9.0e8, wb            # ~242 M conv weights in bf16 + tables
    assert lib.ftc_detector_workspace_bytes(h, 32) > lib.ftc_detector_workspace_bytes(h, 1) > 0
    assert lib.ftc_detector_num_ops(h) > 280
    lib.ftc_detector_destroy(h)
    tcfg = _lib.TransformerConfig(106, 512, 16, 16, 16, 100, 100, _lib.PREC_BF16, _lib.GEMM_TCGEN05)
    _lib.check(lib.ftc_transformer_create(C.byref(tcfg), C.byref(h)))
    assert lib.ftc_transformer_weight_bytes(h) > 2.0e8
    assert lib.ftc_transformer_workspace_bytes(h, 256, 100, 100) > 0
    lib.ftc_transformer_destroy(h)
    # bad configs fail loudly with a message
    bad = _lib.TransformerConfig(106, 500, 16, 1, 1, 100, 100, 0, 0)
    assert lib.ftc_transformer_create(C.byref(bad), C.byref(h)) != 0
    assert b"embed_dim" in lib.ftc_last_error()


def test_product_path_has_no_cpu_fallback():
    from findtextcenternet_b200.models.detector import TextDetectorModel
    from findtextcenternet_b200.models.transformer import ModelDimensions, Transformer
    m = TextDetectorModel(pre_weights=False).eval()
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            m.detector(torch.zeros(1, 3, 768, 768))
    t = Transformer(**ModelDimensions(embed_dim=64, head_num=4, enc_block_num=1, dec_block_num=1).__dict__).eval()
    with pytest.raises(RuntimeError):
        t(torch.zeros(1, 8, 106), torch.zeros(1, 8, dtype=torch.long))
    with pytest.raises(RuntimeError):     # the train-mode (autograd) path has no CPU route either
        m.train().detector(torch.zeros(1, 3, 768, 768))
    # nothing under the package imports the oracle
    pkg = os.path.join(ROOT, "findtextcenternet_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, fn


def test_tile_meta_matches_oracle_mask():
    from findtextcenternet_b200.process_ocr_b200 import tile_meta
    from oracle import detector_oracle as DO
    for (x_i, y_i, w, h) in [(0, 0, 768, 768), (0, 0, 2148, 2148), (460, 0, 2148, 2148), (1380, 920, 2148, 2148), (460, 460, 1228, 1688)]:
        ox, oy, x0, x1, y0, y1 = tile_meta(x_i, y_i, w, h)
        mask = np.zeros((192, 192), bool)
        mask[y0:y1, x0:x1] = True
        assert (ox, oy) == (x_i, y_i)
        assert np.array_equal(mask, DO.tile_mask(x_i, y_i, w, h))


def test_shard_range_partitions():
    from findtextcenternet_b200.shard import shard_range
    for n in (0, 1, 7, 16, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["FTC_ROOT"])
from findtextcenternet_b200.shard import shard_range, gather_ragged, allreduce_gradients
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# inference sharding: 16 tiles -> ranks, ragged per-rank results gathered everywhere
a, b = shard_range(16, rank, world)
local = torch.arange(a, b, dtype=torch.float32).repeat_interleave(rank + 1).unsqueeze(1)   # rank-dependent row count
parts = gather_ragged(local)
assert len(parts) == world
for r, p in enumerate(parts):
    ra, rb = shard_range(16, r, world)
    assert torch.equal(p[:, 0], torch.arange(ra, rb, dtype=torch.float32).repeat_interleave(r + 1)), (r, p)
# training: bucketed gradient averaging equals the mean of per-rank gradients
torch.manual_seed(0)
model = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.Linear(53, 11))
g = torch.Generator().manual_seed(100 + rank)
for p in model.parameters():
    p.grad = torch.randn(p.shape, generator=g)
expect = []
for p in model.parameters():
    acc = torch.zeros_like(p)
    for r in range(world):
        gr = torch.Generator().manual_seed(100 + r)
        # regenerate rank r's gradients in registration order
    expect.append(None)
grads_all = []
for r in range(world):
    gr = torch.Generator().manual_seed(100 + r)
    grads_all.append([torch.randn(p.shape, generator=gr) for p in model.parameters()])
calls = allreduce_gradients(model.parameters(), bucket_bytes=4096)
assert calls >= 2
for i, p in enumerate(model.parameters()):
    mean = sum(grads_all[r][i] for r in range(world)) / world
    assert torch.allclose(p.grad, mean, atol=1e-6), i
# overlapped exchange: buckets launched from inside backward by post-accumulate hooks give the same means
from findtextcenternet_b200.shard import GradientBuckets
torch.manual_seed(1)
net = torch.nn.Sequential(torch.nn.Linear(19, 31), torch.nn.Tanh(), torch.nn.Linear(31, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
unused = torch.nn.Parameter(torch.zeros(5))            # never receives a gradient (like the self-attention pos_emb_k tables)
params = list(net.parameters()) + [unused]
gb = GradientBuckets(params, bucket_bytes=1024)
assert len(gb.buckets) >= 3
for step in range(2):
    for p in params:
        p.grad = None
    gx = torch.Generator().manual_seed(500 + 10 * step + rank)
    net(torch.randn(8, 19, generator=gx)).square().sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    calls = gb.finish()
    assert calls >= 3 and gb.launched_during_backward >= 2 * (step + 1), (calls, gb.launched_during_backward)
    gathered = [torch.zeros_like(torch.cat([g.reshape(-1) for g in local])) for _ in range(world)]
    dist.all_gather(gathered, torch.cat([g.reshape(-1) for g in local]))
    mean = sum(gathered) / world
    assert torch.allclose(torch.cat([p.grad.reshape(-1) for p in net.parameters()]), mean, atol=1e-6)
    assert unused.grad is None
gb.remove()
# gradients that LIVE in flat bucket storage (shard.FlatGradients): in-place bucket all-reduce launched from inside backward,
# views stay attached across steps, zero() replaces zero_grad(), a parameter without a gradient contributes zeros
from findtextcenternet_b200.shard import FlatGradients
torch.manual_seed(2)
net2 = torch.nn.Sequential(torch.nn.Linear(19, 31), torch.nn.Tanh(), torch.nn.Linear(31, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
unused2 = torch.nn.Parameter(torch.zeros(5))
params2 = list(net2.parameters()) + [unused2]
fg = FlatGradients(params2, bucket_bytes=1024)
assert len(fg.buckets) >= 3 and all(p.grad is not None for p in params2)
ptrs = [p.grad.data_ptr() for p in params2]
for step in range(3):
    fg.zero()
    gx = torch.Generator().manual_seed(700 + 10 * step + rank)
    # the same un-averaged local gradients, computed on a detached copy
    ref_net = torch.nn.Sequential(torch.nn.Linear(19, 31), torch.nn.Tanh(), torch.nn.Linear(31, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    ref_net.load_state_dict(net2.state_dict())
    xin = torch.randn(8, 19, generator=gx)
    ref_net(xin).square().sum().backward()
    local = torch.cat([p.grad.reshape(-1) for p in ref_net.parameters()])
    net2(xin).square().sum().backward()
    assert fg.finish() == len(fg.buckets) and fg.launched_during_backward >= 2 * (step + 1)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    mean = sum(gathered) / world
    assert torch.allclose(torch.cat([p.grad.reshape(-1) for p in net2.parameters()]), mean, atol=1e-6), step
    assert float(unused2.grad.abs().max()) == 0.0
    assert [p.grad.data_ptr() for p in params2] == ptrs           # static addresses (CUDA-graph / fused-optimizer tables)
    fg.check_views()
net2[0].weight.grad = None
try:
    fg.check_views()
    raise SystemExit("check_views did not notice the detached gradient view")
except RuntimeError:
    pass
fg.remove()
# train1_step: the nine raw losses that drive the CoV weights take their cross-rank mean VALUE but keep the local gradient path
from findtextcenternet_b200.train import TRAIN1_LOSSES, _sync_loss_values
leaf = torch.full((len(TRAIN1_LOSSES),), float(rank + 1), requires_grad=True)
raw = {k: leaf[i] * (i + 1) for i, k in enumerate(TRAIN1_LOSSES)}
raw["correct"] = torch.tensor(3)
synced = _sync_loss_values(raw)
mean_rank = sum(r + 1 for r in range(world)) / world
for i, k in enumerate(TRAIN1_LOSSES):
    assert abs(float(synced[k]) - mean_rank * (i + 1)) < 1e-6, (k, float(synced[k]))
assert int(synced["correct"]) == 3
sum(synced[k] for k in TRAIN1_LOSSES).backward()
assert torch.allclose(leaf.grad, torch.arange(1, len(TRAIN1_LOSSES) + 1, dtype=torch.float32))
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gloo_world2_sharding_and_gradient_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, FTC_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") >= 2   # the two ranks interleave their prints


@pytest.mark.parametrize("h,w", [(640, 533), (768, 768), (769, 1000), (2048, 2048), (300, 2000)])
def test_page_tiles_match_the_reference_tiling(h, w):
    """process_ocr_base.py:62-76 restated line by line: pad amounts, padded size, tile offsets and their order."""
    from findtextcenternet_b200.process_ocr_b200 import page_tiles
    width = height = 768
    stepx = stepy = int(768 * 0.6)
    im0 = np.full((h, w, 3), 7, dtype=np.uint8)
    padx = max(0, (width - im0.shape[1]) % stepx, width - im0.shape[1])
    pady = max(0, (height - im0.shape[0]) % stepy, height - im0.shape[0])
    ref = np.pad(im0, [[0, pady], [0, padx], [0, 0]], "constant", constant_values=((255, 255), (255, 255), (255, 255)))
    ref_offsets = []
    for y in range(0, ref.shape[0] - height + 1, stepy):
        for x in range(0, ref.shape[1] - width + 1, stepx):
            ref_offsets.append((x, y))
    page, offsets = page_tiles(im0)
    assert page.shape == ref.shape and np.array_equal(page, ref)
    assert offsets == ref_offsets
    if (h, w) == (2048, 2048):
        assert page.shape[:2] == (2148, 2148) and len(offsets) == 16


def test_chunk_planner_matches_the_reference_loop():
    """ocr_text.plan_chunks / chunk_inputs / assemble_text == the window loop of call_OCR (process_ocr_base.py:182-283), executed
    from the UNMODIFIED reference source with a stub transformer (tests/golden/chunks_seed0.json, oracle/make_golden.py::
    golden_chunks): same windows, same encoder inputs, same overlaps, same assembled text -- for pages that need 1 ... 13 windows
    and lengths right at the 397-row window limit."""
    import json
    from conftest import GOLDEN
    from findtextcenternet_b200 import ocr_text
    from oracle.make_golden import chunk_features, stub_codes
    with open(os.path.join(GOLDEN, "chunks_seed0.json")) as f:
        gold = json.load(f)
    assert len(gold) == 7
    for name, g in gold.items():
        feats = chunk_features(g["seed"], g["n"])
        chunks = ocr_text.plan_chunks(feats)
        assert [[c.prev_j, c.cur_i, c.cur_j] for c in chunks] == g["windows"], name
        x = ocr_text.chunk_inputs(feats, chunks)
        assert x.shape == (len(chunks), 400, 106)
        assert [float(np.abs(x[i]).sum()) for i in range(len(chunks))] == g["input_sums"], name
        preds = np.stack([stub_codes(x[i:i + 1]) for i in range(len(chunks))])
        txt, linebuf = ocr_text.assemble_text(chunks, preds)
        assert txt == g["result_txt"], name
        assert [[a, b, c] for a, b, c in linebuf] == g["linebuf"], name


def test_forward_part_argument_checks_need_no_gpu():
    """ftc_detector_forward_part (include/ftc_b200.h: the forward in two parts so that the upload of host tiles overlaps the
    first layers) rejects bad parts / unpacked plans with an error code and a message before any CUDA call."""
    import ctypes as C
    from findtextcenternet_b200 import _lib
    lib = _lib.load()
    cfg = _lib.make_detector_config("xl", 1, 1, 768, 768)      # bf16, tcgen05: the plan is host arithmetic only
    h = C.c_void_p()
    assert lib.ftc_detector_create(C.byref(cfg), C.byref(h)) == 0
    try:
        fake = C.c_void_p(256)
        ws = lib.ftc_detector_workspace_bytes(h, 4)
        assert ws > 0
        # unknown part
        assert lib.ftc_detector_forward_part(h, fake, 4, 7, 0, 4, fake, fake, None, fake, ws, None) != 0
        # null images
        assert lib.ftc_detector_forward_part(h, None, 4, _lib.PART_EARLY, 0, 4, None, None, None, fake, ws, None) != 0
        # weights not packed: refused before anything is launched
        assert lib.ftc_detector_forward_part(h, fake, 4, _lib.PART_EARLY, 0, 2, None, None, None, fake, ws, None) != 0
        assert b"pack_weights" in lib.ftc_last_error()
        # the second part needs the output maps
        assert lib.ftc_detector_forward_part(h, fake, 4, _lib.PART_REST, 0, 4, None, None, None, fake, ws, None) != 0
    finally:
        lib.ftc_detector_destroy(h)
