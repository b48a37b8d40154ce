"""Schedule-free AdamW: oracle vs reference golden (CPU) and fused CUDA step vs the same golden (GPU)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
import os


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "optimizer_seed0.npz"))


def _inputs(step):
    from oracle.make_golden import optimizer_inputs
    return optimizer_inputs(step)


def test_optimizer_oracle_matches_reference(gold):
    from oracle.make_golden import OPT_CFG
    from oracle.optimizer_oracle import AdamWScheduleFreeOracle
    # optimizer.train() on fresh params is a no-op (no 'z' yet): y starts at the initial parameters
    o = AdamWScheduleFreeOracle([p.numpy() for p in _inputs(-1)], **OPT_CFG)
    for step in range(5):
        o.step([g.numpy() for g in _inputs(step)])
        for i in range(len(o.y)):
            np.testing.assert_allclose(o.y[i], gold[f"s{step}_y{i}"], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(o.z[i], gold[f"s{step}_z{i}"], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(o.v[i], gold[f"s{step}_v{i}"], rtol=2e-6, atol=1e-12)
    for i, x in enumerate(o.eval_params()):
        np.testing.assert_allclose(x, gold[f"eval_x{i}"], rtol=2e-5, atol=1e-6)


@pytest.mark.gpu
def test_fused_adamw_schedulefree_matches_reference(gold):
    """fp32, tolerance 2e-6 relative (fused multiply-add contraction differs from the foreach sequence by <= 1 ulp)."""
    from oracle.make_golden import OPT_CFG
    from findtextcenternet_b200 import _lib
    from findtextcenternet_b200.models.adamw_schedulefree import AdamWScheduleFree
    params = [torch.nn.Parameter(p.clone().cuda()) for p in _inputs(-1)]
    opt = AdamWScheduleFree(params, **OPT_CFG)
    with pytest.raises(Exception):
        opt.step()                       # not in train mode: same error as the reference
    opt.train()
    l0 = _lib.launch_count()
    for step in range(5):
        for p, g in zip(params, _inputs(step)):
            p.grad = g.clone().cuda()
        opt.step()
        for i, p in enumerate(params):
            np.testing.assert_allclose(p.detach().cpu().numpy(), gold[f"s{step}_y{i}"], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(opt.state[p]["z"].cpu().numpy(), gold[f"s{step}_z{i}"], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(opt.state[p]["exp_avg_sq"].cpu().numpy(), gold[f"s{step}_v{i}"], rtol=2e-6, atol=1e-12)
    assert _lib.launch_count() - l0 == 5       # one fused launch per step for all tensors
    opt.eval()
    for i, p in enumerate(params):
        np.testing.assert_allclose(p.detach().cpu().numpy(), gold[f"eval_x{i}"], rtol=2e-5, atol=1e-6)
    assert opt.param_groups[0]["k"] == 5 and set(opt.state[params[0]].keys()) == {"z", "exp_avg_sq"}


# ---- RAdamScheduleFree (reference models/radam_schedulefree.py, train3.py:121) ----
@pytest.fixture(scope="module")
def gold_radam():
    return np.load(os.path.join(GOLDEN, "optimizer_radam_seed0.npz"))


@pytest.mark.parametrize("name", ["silent", "sgd"])
def test_radam_oracle_matches_reference(gold_radam, name):
    from oracle.make_golden import RADAM_CFGS
    from oracle.optimizer_oracle import RAdamScheduleFreeOracle
    o = RAdamScheduleFreeOracle([p.numpy() for p in _inputs(-1)], **RADAM_CFGS[name])
    for step in range(8):
        o.step([g.numpy() for g in _inputs(step)])
        for i in range(3):
            np.testing.assert_allclose(o.y[i], gold_radam[f"{name}_s{step}_y{i}"], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(o.z[i], gold_radam[f"{name}_s{step}_z{i}"], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(o.v[i], gold_radam[f"{name}_s{step}_v{i}"], rtol=2e-6, atol=1e-12)
    for i, x in enumerate(o.eval_params()):
        np.testing.assert_allclose(x, gold_radam[f"{name}_eval_x{i}"], rtol=2e-5, atol=1e-6)
    # the fixture crosses rho_t = 4: the scheduled lr is 0 (silent) / lr (SGD phase) first and rectified afterwards
    lrs = [float(gold_radam[f"{name}_s{s}_lr"]) for s in range(8)]
    assert lrs[-1] > 0 and lrs[-1] != lrs[0]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["silent", "sgd"])
def test_fused_radam_schedulefree_matches_reference(gold_radam, name):
    from oracle.make_golden import RADAM_CFGS
    from findtextcenternet_b200 import _lib
    from findtextcenternet_b200.models.radam_schedulefree import RAdamScheduleFree
    params = [torch.nn.Parameter(p.clone().cuda()) for p in _inputs(-1)]
    opt = RAdamScheduleFree(params, **RADAM_CFGS[name])
    with pytest.raises(Exception):
        opt.step()
    opt.train()
    l0 = _lib.launch_count()
    for step in range(8):
        for p, g in zip(params, _inputs(step)):
            p.grad = g.clone().cuda()
        opt.step()
        assert abs(opt.param_groups[0]["scheduled_lr"] - float(gold_radam[f"{name}_s{step}_lr"])) < 1e-15
        for i, p in enumerate(params[:3]):
            np.testing.assert_allclose(p.detach().cpu().numpy(), gold_radam[f"{name}_s{step}_y{i}"], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(opt.state[p]["z"].cpu().numpy(), gold_radam[f"{name}_s{step}_z{i}"], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(opt.state[p]["exp_avg_sq"].cpu().numpy(), gold_radam[f"{name}_s{step}_v{i}"], rtol=2e-6, atol=1e-12)
    assert _lib.launch_count() - l0 == 8
    opt.eval()
    for i, p in enumerate(params):
        np.testing.assert_allclose(p.detach().cpu().numpy(), gold_radam[f"{name}_eval_x{i}"], rtol=2e-5, atol=1e-6)
    assert set(opt.param_groups[0].keys()) >= {"silent_sgd_phase", "scheduled_lr", "weight_sum", "lr_max", "k", "train_mode"}


@pytest.mark.gpu
def test_graph_replayed_adamw_step_matches_reference(gold):
    """The CUDA-graph form of the step (ftc_adamw_sf_step_dev: schedule state k / lr_max / weight_sum on the device, advanced by
    each replay) against the SAME reference golden: one eager step, then the remaining four as replays of one captured graph
    whose gradients are refreshed in place (static gradient storage, as shard.FlatGradients provides)."""
    from oracle.make_golden import OPT_CFG
    from findtextcenternet_b200.models.adamw_schedulefree import AdamWScheduleFree
    params = [torch.nn.Parameter(p.clone().cuda()) for p in _inputs(-1)]
    opt = AdamWScheduleFree(params, **OPT_CFG)
    opt.train()
    for p, g in zip(params, _inputs(0)):
        p.grad = g.clone().cuda()
    opt.step()                                            # eager: creates z / exp_avg_sq
    opt.prepare_graph()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt.step()
    for step in range(1, 5):
        for p, g in zip(params, _inputs(step)):
            p.grad.copy_(g.cuda())                        # same storage: the graph's pointers stay valid
        graph.replay()
        torch.cuda.synchronize()
        for i, p in enumerate(params):
            np.testing.assert_allclose(p.detach().cpu().numpy(), gold[f"s{step}_y{i}"], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(opt.state[p]["z"].cpu().numpy(), gold[f"s{step}_z{i}"], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(opt.state[p]["exp_avg_sq"].cpu().numpy(), gold[f"s{step}_v{i}"], rtol=2e-6, atol=1e-12)
    opt.sync_from_graph()
    assert opt.param_groups[0]["k"] == 5
    opt.eval()
    for i, p in enumerate(params):
        np.testing.assert_allclose(p.detach().cpu().numpy(), gold[f"eval_x{i}"], rtol=2e-5, atol=1e-6)
