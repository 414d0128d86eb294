/* uhc_rollout.h -- C ABI of the sampling loop (part of libuhc_b200.so): policy forward and the fused T-step rollout.
 *
 * Reference interface replaced (SURVEY.md section 8b lists these as the surface a C-ABI replacement must export):
 *   uhc_policy_forward   Agent.trans_policy + running_state + PolicyGaussian.select_action
 *                        (uhc/agents/agent_copycat.py:521-531, khrylib/rl/core/policy.py:12-15, policy_gaussian.py:26-31, utils/zfilter.py:59-73)
 *   uhc_rollout          AgentCopycat.sample_worker's loop body over all environments (uhc/agents/agent_copycat.py:496-571): per control step
 *                        normalise -> policy -> sample -> env.step -> custom_reward -> push(state, action, mask, reward, exp); finished
 *                        episodes are re-seeded inside the step kernel (UhcEnvCfg.auto_reset, :503-517 + dataset_amass_single.py:172-253).
 * All pointers are CUDA device pointers owned by the caller (PyTorch tensors: weights, ZFilter statistics, rollout buffer) unless
 * suffixed _host.  Calls are stream-ordered on `stream`; return 0 on success, <0 on error (uhc_rollout_last_error()).
 */
#ifndef UHC_ROLLOUT_H
#define UHC_ROLLOUT_H
#include "uhc_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* MLP trunk + head in the tensor-core layout (khrylib/models/mlp.py:5-27 + PolicyGaussian.action_mean): layer i maps dims[i] -> dims[i+1];
 * W_bf16[i] = [dims[i+1]][kp[i]] bf16, kp[i] = dims[i] rounded up to 64 and zero padded; bias[i] fp32; act = UHC_ACT_* of the trunk. */
typedef struct {
    int nlayers, act;
    int dims[10];
    int kp[8];
    const void *W_bf16[8];
    const float *bias[8];
} UhcMlp;

/* PolicyMCP (uhc/models/policy_mcp.py:9-37, actor_type "mcp" of config/release/uhc_implicit.yml): nprim primitive MLPs (same widths as a
 * PolicyGaussian trunk + action_mean head) and a composer MLP obs -> composer_dim -> nprim whose EVERY layer is followed by the activation
 * (mlp.py:24-27) before the softmax; action_mean = sum_k softmax(composer(x))_k * prim_k(x). */
#define UHC_MCP_MAX_PRIM 8
typedef struct {
    int nprim, reserved;
    UhcMlp prim[UHC_MCP_MAX_PRIM];
    UhcMlp composer;
} UhcMcp;

/* Time-major rollout buffer [T_cap][E][...] (the device-resident TrajBatch, khrylib/rl/core/trajbatch.py:4-15) and the current
 * observation of every env (written by uhc_env_reset / the step kernel). logp / fails may be NULL. */
typedef struct {
    float *states, *actions, *rewards, *masks, *exps, *logp;
    int *fails;
    float *obs_cur;       /* [E][uhc_engine_obs_dim] raw observation before the step; overwritten with the next one */
    int *ep_clip;         /* optional [T_cap][E]: clip index of the episode that ended at this step (-1: none) ... */
    float *ep_pct;        /* ... and its completed fraction: the per-clip success history of agent_copycat.py:561 */
    int T_cap, reserved;
} UhcRolloutBuf;

const char *uhc_rollout_last_error(void);

/* RNG stream position of the action noise: element (step, env, dim) of the stream `seed`; advanced by one per rollout step. */
int uhc_rollout_set_step(UhcEngine *e, unsigned long long step);
int uhc_rollout_get_step(UhcEngine *e, unsigned long long *step);

/* obs [E][657] -> ZFilter (update_filter != 0 merges the batch into zfilter_stats first) -> MLP on tensor cores -> Gaussian head:
 * action = mean + exp(log_std) eps (the mean where mean_action[e] != 0), logp = summed Normal log-prob.  state_out = the normalised
 * observation (what the reference pushes into its memory), may be NULL.  Does not advance the step counter. */
int uhc_policy_forward(UhcEngine *e, const float *obs_dev, const UhcMlp *mlp, const float *log_std, double *zfilter_stats, float zclip, int update_filter,
                       unsigned long long seed, const unsigned char *mean_action_or_null, float *state_out_or_null, float *action_out, float *logp_out_or_null,
                       void *stream);

/* the same with a PolicyMCP mixture in place of the single MLP */
int uhc_policy_forward_mcp(UhcEngine *e, const float *obs_dev, const UhcMcp *mcp, const float *log_std, double *zfilter_stats, float zclip, int update_filter,
                           unsigned long long seed, const unsigned char *mean_action_or_null, float *state_out_or_null, float *action_out, float *logp_out_or_null,
                           void *stream);

/* T lock-step control steps of every env into rows row0 .. row0+T-1 of the buffer.  noise_rate: P(sampled action) per env and step
 * (agent_copycat.py:530; exp = 1 for sampled rows).  use_graph != 0: the kernel sequence is captured once per argument set into a
 * CUDA graph and replayed (no host work between kernels); 0: plain stream launches of the same kernels (bit-identical results). */
int uhc_rollout(UhcEngine *e, int T, int row0, const UhcMlp *mlp, const float *log_std, double *zfilter_stats, float zclip, int update_filter,
                unsigned long long seed, float noise_rate, const UhcRolloutBuf *buf, int use_graph, void *stream);
int uhc_rollout_mcp(UhcEngine *e, int T, int row0, const UhcMcp *mcp, const float *log_std, double *zfilter_stats, float zclip, int update_filter,
                    unsigned long long seed, float noise_rate, const UhcRolloutBuf *buf, int use_graph, void *stream);
/* measurement hook: CUDA events around the env-step kernel of buffer rows 0 .. nrows-1, recorded on the launching stream (also inside
 * graph replays); uhc_rollout_env_step_ms returns the duration of the last step written to `row`.  nrows = 0 disables. */
int uhc_rollout_time_env_step(UhcEngine *e, int nrows);
int uhc_rollout_env_step_ms(UhcEngine *e, int row, float *ms);
int uhc_rollout_launches_per_step(UhcEngine *e);   /* kernels per control step of the last uhc_rollout (bench `gpu_launches`) */
void uhc_rollout_release(UhcEngine *e);            /* frees graphs / scratch of this engine; call before uhc_engine_destroy */

#ifdef __cplusplus
}
#endif
#endif
