"""evaluate_two_policies_in_batch on the device (competitive_rl/pong/evaluate.py:53-88).

Same contract as the reference: two action functions play each other on a cPongDouble vec-env until `num_episodes`
episodes have finished; returns ([win, draw, lose, cumulative reward] of agent 0, the same for agent 1).  The per-env
bookkeeping (episode returns, win / draw / lose classification on done, zeroing finished envs) is tensor arithmetic on
the device instead of a Python loop over envs; the only host traffic per step is the finished-episode count that the
stopping rule needs."""
import torch


def evaluate_two_policies_in_batch(compute_action0, compute_action1, envs, num_episodes):
    dev = envs.device
    n = envs.num_envs
    episode_rewards = torch.zeros((n, 2), dtype=torch.float64, device=dev)
    tally = torch.zeros(5, dtype=torch.float64, device=dev)     # wins0, draws, losses0, sum reward0, sum reward1
    total_episodes = 0
    obs = envs.reset()
    while True:
        a0 = torch.as_tensor(compute_action0(obs[0])).to(dev, dtype=torch.int32).reshape(-1)
        a1 = torch.as_tensor(compute_action1(obs[1])).to(dev, dtype=torch.int32).reshape(-1)
        obs, reward, done, info = envs.step(torch.stack([a0, a1], dim=1))
        done = done.bool()
        if done.ndim == 2:
            done = done.all(dim=1)
        episode_rewards += reward.to(torch.float64).reshape(n, 2)
        r0 = episode_rewards[:, 0]
        tally += torch.stack([(done & (r0 > 0)).sum(), (done & (r0 == 0)).sum(), (done & (r0 < 0)).sum(),
                              (r0 * done).sum(), (episode_rewards[:, 1] * done).sum()]).to(torch.float64)
        episode_rewards *= (~done).to(torch.float64).reshape(-1, 1)
        total_episodes += int(done.sum())                       # the one host sync of the step
        if total_episodes >= num_episodes:
            break
    w, d, l, s0, s1 = [float(x) for x in tally.tolist()]
    return [int(w), int(d), int(l), s0], [int(l), int(d), int(w), s1]
