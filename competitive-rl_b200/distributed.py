"""Multi-GPU plumbing: environments are independent, so a job of `total_envs` environments on
`world` GPUs is `world` shards stepped with NO collective on the step path.
The only communication is the optional end-of-run reduction of the episode statistics vector."""
import torch

STAT_FIELDS = ["episodes", "sum_length", "left_wins", "right_wins", "draws", "sum_margin_biased"]


def shard_range(total_envs, rank, world):
    """Contiguous shard [first, first + count) of `rank`; shards differ by at most one env.
    RNG streams are keyed by GLOBAL env index, so any sharding reproduces the same envs."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(int(total_envs), int(world))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def stats_from_raw(raw):
    """Decode the uint64[8] vector of crl_pong_get_stats."""
    ep = int(raw[0])
    d = {k: int(raw[i]) for i, k in enumerate(STAT_FIELDS)}
    d["mean_length"] = d["sum_length"] / ep if ep else 0.0
    d["mean_margin"] = (d["sum_margin_biased"] - 64 * ep) / ep if ep else 0.0   # mean(score_left - score_right)
    return d


def raw_from_stats(d):
    return [int(d[k]) for k in STAT_FIELDS] + [0, 0]


def gather_episode_stats(local_stats, group=None, device=None):
    """Sum the per-shard statistics over all ranks (one all-reduce of 8 int64; NCCL when `device`
    is a CUDA device, gloo otherwise).  Returns the decoded whole-job dict on every rank."""
    import torch.distributed as dist
    t = torch.tensor(raw_from_stats(local_stats), dtype=torch.int64, device=device or "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return stats_from_raw(t.tolist())
