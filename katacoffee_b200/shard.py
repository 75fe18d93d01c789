"""Multi-GPU plumbing for the hot path: games are independent, so ranks only agree on disjoint game-id
ranges and sum their statistics once at the end (the reference's analogue is one server thread +
ComputeHandle per GPU over a shared LoadedModel, cpp/program/setup.cpp:190-229).  No data-path collective."""

GAME_ID_STRIDE = 1 << 40   # ids of rank r live in [r * 2^40, (r+1) * 2^40): never collide, refills included

STAT_FIELDS = ("steps", "evals", "gamesFinished", "blackWins", "whiteWins", "draws")


def first_game_id(rank):
    return rank * GAME_ID_STRIDE


def reduce_stats(values, device="cpu"):
    """End-of-run reduce(sum) of the statistics counters to rank 0 (NCCL on GPUs, gloo in the CPU tests).
    `values`: sequence of ints in STAT_FIELDS order.  Returns the totals on rank 0 (local values elsewhere)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    return [int(x) for x in t.tolist()]


def max_over_ranks(value, device="cpu"):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
