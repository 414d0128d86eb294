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


def _mcp_from_golden(nn, g, dev):
    hs, P, cdim = tuple(int(x) for x in g["hsize"]), int(g["nprim"]), tuple(int(x) for x in g["composer_dim"])
    S, A = g["states"].shape[1], g["actions"].shape[1]
    pol = nn.MCPNet(S, hs, A, "relu", num_primitive=P, composer_dim=cdim, device=dev, seed=1)
    pol.load_state_dict({k[3:]: g[k] for k in g.files if k.startswith("p0.")})
    val = nn.MLPNet(S, hs, 1, "relu", device=dev, head_name="value_head", seed=2)
    val.load_state_dict({k[3:]: g[k] for k in g.files if k.startswith("v0.")})
    return pol, val


def test_policy_mcp_forward_matches_reference(golden_dir):
    """PolicyMCP.forward (uhc/models/policy_mcp.py:28-36) of the unmodified reference in fp64 (tools/make_golden.py gen_mcp): primitives, composer with its
    activated last layer, softmax, weighted sum -- on the fp32 SIMT GEMMs (2e-6) and on the tensor-core path the rollout uses (bf16 operands)."""
    import os
    import torch
    from uhc_b200 import nn
    g = np.load(os.path.join(golden_dir, "mcp_ppo.npz"))
    dev = torch.device("cuda", 0)
    pol, val = _mcp_from_golden(nn, g, dev)
    x = torch.tensor(g["states"], device=dev)
    scale = np.abs(g["mean"]).mean()
    m32 = pol.forward(x).cpu().numpy()
    assert np.abs(m32 - g["mean"]).max() < 2e-6 + 1e-4 * scale
    mtc = pol.forward_tc(x).cpu().numpy()
    assert np.abs(mtc - g["mean"]).mean() < 0.02 * scale and np.abs(mtc - g["mean"]).max() < 0.25 * np.abs(g["mean"]).max()
    # checkpoint keys of the reference's module tree
    sd = pol.state_dict()
    assert "nets.0.0.affine_layers.0.weight" in sd and "nets.3.1.bias" in sd and "composer.0.affine_layers.2.weight" in sd
    assert len(sd) == len([k for k in g.files if k.startswith("p0.") and k != "p0.action_log_std"])


def test_policy_mcp_update_matches_reference(golden_dir):
    """AgentPPO.update_policy with a PolicyMCP actor (2 epochs, Adam, first-step grad clip) against uhc_ppo_update_policy on an MCP trainer
    (uhc_ppo_trainer_create_mcp: mixture backward into the primitives and through the softmax into the composer, ONE flat parameter tensor)."""
    import os
    import torch
    from uhc_b200 import nn
    g = np.load(os.path.join(golden_dir, "mcp_ppo.npz"))
    dev = torch.device("cuda", 0)
    pol, val = _mcp_from_golden(nn, g, dev)
    op, ov = nn.Adam(pol.params(), 5e-5, net=pol), nn.Adam(val.params(), 3e-4, net=val)
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float32).reshape(len(a), -1).squeeze(-1) if np.asarray(a).ndim == 2 and np.asarray(a).shape[1] == 1 else np.asarray(a, dtype=np.float32), device=dev)
    M = g["states"].shape[0]
    tr = nn.CPpoTrainer(pol, val, op, ov, M, 1, dev)
    log_std = torch.full((g["actions"].shape[1],), -2.3, device=dev)
    w0 = pol.flat.clone()
    tr.update_policy(t(g["states"]), t(g["actions"]), t(g["returns"]), t(g["advantages"]), t(g["exps"]), log_std, 0.2, int(g["epochs"]), 40.0, torch.zeros(2, device=dev))
    torch.cuda.synchronize()
    assert op.step_n == 2 and ov.step_n == 2 and not torch.equal(w0, pol.flat)
    sd = pol.state_dict()
    rels = []
    for k in [k for k in g.files if k.startswith("p1.")]:
        name = k[3:]
        if name == "action_log_std":
            continue
        d_ref = g[k].astype(np.float64) - g["p0." + name].astype(np.float64)
        d_our = sd[name].numpy().astype(np.float64) - g["p0." + name].astype(np.float64)
        if np.abs(d_ref).mean() > 1e-9:
            rels.append((np.abs(d_our - d_ref).mean() / np.abs(d_ref).mean(), name))
    worst = max(rels)
    # same bound as the single-MLP tensor-core update (bf16 operands; Adam's first steps are ~ lr sign(g)): mean deviation below 10 % of the mean update
    assert np.mean([r for r, _ in rels]) < 0.10 and worst[0] < 0.35, (np.mean([r for r, _ in rels]), worst)
    m1 = pol.forward(torch.tensor(g["states"][:64], device=dev)).cpu().numpy()
    dref = g["mean_after"] - g["mean"][:64]
    assert np.abs((m1 - g["mean"][:64]) - dref).mean() < 0.10 * np.abs(dref).mean()
    tr.close()


def test_policy_mcp_rollout_and_train_iteration(golden_dir):
    """config/release/uhc_implicit.yml shape end to end on the C paths: obs v1 (784), 75-wide actions (no meta-PD), relu PolicyMCP through uhc_rollout_mcp
    (CUDA graph) and uhc_ppo_update; uhc_policy_forward_mcp reproduces MCPNet.forward_tc."""
    import os
    import ctypes as C
    import torch
    from uhc_b200 import nn
    from uhc_b200.agent import BatchedAgent
    z = np.load(os.path.join(golden_dir, "expert_sway.npz"))
    ex = {k: z[k] for k in z.files}
    so = np.concatenate([ex["beta"][0], [ex["gender"][0]]])
    ag = BatchedAgent(64, [ex], [so], policy_hsize=(128, 64), value_hsize=(128, 64), htype="relu", num_optim_epoch=2, t_min=15, t_max=60, actor_type="mcp",
                      num_primitive=4, composer_dim=(48, 32), obs_v=1, meta_pd=0)
    assert ag.obs_dim == 784 and ag.act_dim == 75 and isinstance(ag.policy, nn.MCPNet)
    ag.reset_envs()
    # uhc_policy_forward_mcp (mean action: no noise) against the Python composition of the same kernels
    L, st = ag.engine.lib, C.c_void_p(torch.cuda.current_stream().cuda_stream)
    mc = nn.mcp_struct(ag.policy)
    act = torch.empty(64, 75, device=ag.dev); state = torch.empty(64, 784, device=ag.dev)
    ones = torch.ones(64, dtype=torch.uint8, device=ag.dev)
    L.uhc_rollout_last_error.restype = C.c_char_p
    rc = L.uhc_policy_forward_mcp(ag.engine.h, C.c_void_p(ag.obs.data_ptr()), C.byref(mc), C.c_void_p(ag.log_std.data_ptr()), C.c_void_p(ag.running_state.stats.data_ptr()),
                                  C.c_float(5.0), C.c_int(1), C.c_ulonglong(3), C.c_void_p(ones.data_ptr()), C.c_void_p(state.data_ptr()), C.c_void_p(act.data_ptr()), None, st)
    assert rc == 0, L.uhc_rollout_last_error()
    torch.cuda.synchronize()
    ref = ag.policy.forward_tc(state)
    assert (act - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())
    w0 = ag.policy.flat.clone()
    log = ag.optimize_policy(8)
    torch.cuda.synchronize()
    assert log["num_steps"] == 8 * 64 and np.isfinite(log["avg_reward"]) and np.isfinite(log["surr_loss"]) and np.isfinite(log["value_loss"])
    assert not torch.equal(w0, ag.policy.flat) and torch.isfinite(ag.policy.flat).all()
    assert ag.engine.counters["invalid_env_steps"] == 0
