"""GPU: rollout + PPO update driver (uhc_b200/agent.py) end-to-end sanity and buffer semantics."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _clips(golden_dir):
    out, shapes = [], []
    for tag in ("sway", "kick"):
        z = np.load(os.path.join(golden_dir, f"expert_{tag}.npz"))
        out.append({k: z[k] for k in z.files})
        shapes.append(np.concatenate([z["beta"][0], [z["gender"][0]]]))
    return out, shapes


def test_rollout_buffer_semantics_and_update(golden_dir):
    import torch
    from uhc_b200.agent import BatchedAgent
    clips, shapes = _clips(golden_dir)
    ag = BatchedAgent(96, clips, shapes, policy_hsize=(128, 64), value_hsize=(128, 64), num_optim_epoch=2, t_max=40)
    buf, log = ag.sample(24)
    assert log["num_steps"] == 24 * 96
    assert torch.isfinite(buf.states).all() and torch.isfinite(buf.actions).all() and torch.isfinite(buf.rewards).all()
    assert float(buf.states.abs().max()) <= 5.0 + 1e-6                     # ZFilter clip (agent_copycat.py:147)
    r = buf.rewards.cpu().numpy()
    assert (r >= 0).all() and (r <= 1.0 + 1e-6).all()
    m = buf.masks.cpu().numpy()
    assert set(np.unique(m)) <= {0.0, 1.0} and (m == 0).sum() == log["num_episodes"] > 0
    # an untrained policy falls: episodes end by failure or by clip end within t_max
    assert log["avg_episode_len"] <= 40
    w0 = [p.clone() for p in ag.policy.params()]
    out = ag.update_params(buf)
    assert np.isfinite(out["surr_loss"]) and np.isfinite(out["value_loss"])
    assert any((a - b).abs().max().item() > 0 for a, b in zip(w0, ag.policy.params()))
    # checkpoint wire format round trip (agent_copycat.py:190-201 keys)
    cp = ag.state_dicts()
    assert "net.affine_layers.0.weight" in cp["policy_dict"] and "action_mean.bias" in cp["policy_dict"] and "action_log_std" in cp["policy_dict"]
    assert "value_head.weight" in cp["value_dict"]
    ag2 = BatchedAgent(8, clips, shapes, policy_hsize=(128, 64), value_hsize=(128, 64))
    ag2.load_state_dicts(cp)
    x = torch.randn(8, 657, device="cuda")
    assert torch.allclose(ag.policy.forward(x), ag2.policy.forward(x))


def test_deterministic_rollout_given_seed(golden_dir):
    from uhc_b200.agent import BatchedAgent
    clips, shapes = _clips(golden_dir)
    a = BatchedAgent(32, clips, shapes, policy_hsize=(64,), value_hsize=(64,), seed=3)
    b = BatchedAgent(32, clips, shapes, policy_hsize=(64,), value_hsize=(64,), seed=3)
    ba, _ = a.sample(6)
    bb, _ = b.sample(6)
    assert (ba.actions == bb.actions).all() and (ba.rewards == bb.rewards).all()


def test_explicit_residual_force_train_iteration(golden_dir):
    """uhc_explicit shape end to end: 315-wide actions through uhc_rollout (CUDA graph) and uhc_ppo_update."""
    import torch
    from uhc_b200.agent import BatchedAgent
    z = np.load(os.path.join(golden_dir, "expert_sway.npz"))
    ex = {k: z[k] for k in z.files}
    so = np.concatenate([ex["beta"][0], [ex["gender"][0]]])
    ag = BatchedAgent(64, [ex], [so], policy_hsize=(128, 64), value_hsize=(128, 64), num_optim_epoch=2, t_min=15, t_max=60, rfc_mode="explicit")
    assert ag.act_dim == 315 and ag.policy.dims[-1] == 315
    w0 = ag.policy.flat.clone()
    log = ag.optimize_policy(8)
    torch.cuda.synchronize()
    assert log["num_steps"] == 8 * 64 and np.isfinite(log["avg_reward"]) and np.isfinite(log["surr_loss"]) and np.isfinite(log["value_loss"])
    assert not torch.equal(w0, ag.policy.flat) and torch.isfinite(ag.policy.flat).all()
    assert ag.engine.counters["invalid_env_steps"] == 0


def test_obs_v3_train_iteration(golden_dir):
    """obs_v 3 (3 future frames -> 1920-wide observations without the shape vector), no residual force (99-wide actions) through uhc_rollout / uhc_ppo_update"""
    import torch
    from uhc_b200.agent import BatchedAgent
    z = np.load(os.path.join(golden_dir, "expert_sway.npz"))
    ex = {k: z[k] for k in z.files}
    so = np.concatenate([ex["beta"][0], [ex["gender"][0]]])
    ag = BatchedAgent(64, [ex], [so], policy_hsize=(128, 64), value_hsize=(128, 64), htype="relu", num_optim_epoch=2, t_min=15, t_max=60, obs_v=3, fut_frames=3,
                      fut_skip=10, has_shape=False, rfc_mode="none")
    assert ag.obs_dim == 1920 and ag.act_dim == 99
    w0 = ag.policy.flat.clone()
    log = ag.optimize_policy(8)
    torch.cuda.synchronize()
    assert log["num_steps"] == 8 * 64 and np.isfinite(log["avg_reward"]) and np.isfinite(log["surr_loss"]) and np.isfinite(log["value_loss"])
    assert not torch.equal(w0, ag.policy.flat) and torch.isfinite(ag.policy.flat).all()
