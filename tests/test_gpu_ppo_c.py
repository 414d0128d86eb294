"""GPU: uhc_ppo_update (include/uhc_ppo.h, the whole of agent.update behind one C-ABI call) against the Python-orchestrated sequence of the
same kernels (nn.ppo_epochs_tc, itself pinned to khrylib's AgentPPO.update_policy by tests/test_gpu_product_paths.py) and against
nn.gae (pinned to estimate_advantages by tests/test_gpu_nn.py)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _nets(nn, D, A, hs, dev):
    pol = nn.MLPNet(D, hs, A, "gelu", device=dev, head_name="action_mean", seed=3)
    val = nn.MLPNet(D, hs, 1, "gelu", device=dev, head_name="value_head", seed=4)
    return pol, val, nn.Adam(pol.params(), 5e-5, net=pol), nn.Adam(val.params(), 3e-4, net=val)


@pytest.mark.parametrize("shape", [(8, 256, (256, 128)), (5, 200, (320, 192, 64))])
def test_c_update_matches_python_orchestration(shape):
    import torch
    from uhc_b200 import nn
    T, E, hs = shape
    D, A, M = 657, 105, shape[0] * shape[1]
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    f = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32).to(dev)
    states, last_states, actions = f(M, D).clamp(-5, 5), f(E, D).clamp(-5, 5), 0.3 * f(M, A)
    rewards = torch.rand(T, E, generator=g, dtype=torch.float32).to(dev)
    masks = (torch.rand(T, E, generator=g) > 0.06).float().to(dev)
    exps = (torch.rand(M, generator=g) > 0.1).float().to(dev)
    log_std = torch.full((A,), -2.3, device=dev)
    gamma, tau, eps, epochs, clipn = 0.95, 0.95, 0.2, 3, 40.0

    # ---- the Python-orchestrated sequence (what round 1 ran)
    pol1, val1, op1, ov1 = _nets(nn, D, A, hs, dev)
    p0, v0w = pol1.flat.clone(), val1.flat.clone()
    tp, tv = nn.TCTrainer(pol1), nn.TCTrainer(val1)
    pol1._tc_trainer, val1._tc_trainer = tp, tv
    xb, xT = tp.prepare_input(states)
    tv.cache["xb"], tv.cache["xT"] = xb, xT
    v0, ctx0 = tv.forward(xb)
    last_v = val1.forward_tc(last_states).reshape(E).clone()
    adv, ret = nn.gae(rewards, masks, v0.reshape(T, E), last_v, gamma, tau, normalize=True)
    adv, ret = adv.reshape(-1).clone(), ret.reshape(-1).clone()
    l1 = nn.ppo_epochs_tc(pol1, val1, log_std, op1, ov1, xb, xT, actions, ret, adv, exps, eps, epochs, clipn, first_value=(v0, ctx0)).clone()
    torch.cuda.synchronize()

    # ---- one C-ABI call
    pol2, val2, op2, ov2 = _nets(nn, D, A, hs, dev)
    assert torch.equal(pol2.flat, p0) and torch.equal(val2.flat, v0w)
    tr = nn.CPpoTrainer(pol2, val2, op2, ov2, M, E, dev)
    l2 = torch.zeros(2, device=dev)
    tr.update(states, last_states, actions, rewards, masks, exps, log_std, T, E, gamma, tau, eps, epochs, clipn, l2)
    torch.cuda.synchronize()
    assert op2.step_n == epochs and ov2.step_n == epochs and op2._clip_consumed

    a2 = tr.advantages(M)
    assert (a2 - adv).abs().max().item() < 2e-4 * max(1.0, adv.abs().max().item())          # same kernels, same inputs
    for name, w1, w2, w0, lr in (("policy", pol1.flat, pol2.flat, p0, 5e-5), ("value", val1.flat, val2.flat, v0w, 3e-4)):
        d1, d2 = (w1 - w0), (w2 - w0)
        assert d1.abs().max().item() > 0.5 * lr                                             # the nets did move
        # identical kernels in identical order; what differs is the order of fp32 atomic accumulations inside them (split-K dW, bias sums)
        assert (d1 - d2).abs().mean().item() < 0.02 * d1.abs().mean().item(), name
        assert (d1 - d2).abs().max().item() < 0.6 * epochs * lr, name
    assert abs(l1[0].item() - l2[0].item()) < 1e-3 * max(1.0, abs(l1[0].item())) and abs(l1[1].item() - l2[1].item()) < 1e-3 * max(1.0, abs(l1[1].item()))
    # the bf16 weight copies the rollout graph reads were refreshed in place
    for i, (w, wb) in enumerate(zip(pol2.W, pol2._bf16_store)):
        assert torch.equal(wb[:, :w.shape[1]], w.to(torch.bfloat16)), i
    # second call: Adam state and the clip flag carry over (no clip on later policy steps, as the reference's consumed generator)
    tr.update(states, last_states, actions, rewards, masks, exps, log_std, T, E, gamma, tau, eps, 1, clipn, l2)
    torch.cuda.synchronize()
    assert op2.step_n == epochs + 1 and torch.isfinite(pol2.flat).all() and torch.isfinite(val2.flat).all()
    tr.close()


def test_c_update_rejects_bad_arguments():
    import torch
    from uhc_b200 import nn
    dev = torch.device("cuda", 0)
    pol, val, op, ov = _nets(nn, 657, 105, (128,), dev)
    tr = nn.CPpoTrainer(pol, val, op, ov, 512, 64, dev)
    z = torch.zeros(1024 * 657, device=dev)
    with pytest.raises(RuntimeError, match="capacity"):
        tr.update(z, z, z, z, z, z, z, 16, 64, 0.95, 0.95, 0.2, 1, 40.0, torch.zeros(2, device=dev))
    with pytest.raises(RuntimeError, match="ncclComm_t"):
        tr.update(z, z, z, z, z, z, z, 4, 64, 0.95, 0.95, 0.2, 1, 40.0, torch.zeros(2, device=dev), world=2)
    tr.close()
