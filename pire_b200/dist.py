"""Sharding of a batch across the GPUs of one box (one process per GPU).

The path partitions by string: every Run() depends only on its own bytes and the
replicated scanner tables, so there is no data-path collective.
The single exchange is at the end, on the match bitmap: shards are contiguous
string ranges whose boundaries are multiples of 32 strings, so bitmap words never
straddle ranks and an all-reduce(SUM) of the zero-initialised full-length bitmap
equals a bitwise OR.  ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests)
is plumbing only.
"""
import os


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_bounds(n_strings, rank, world):
    """[lo, hi) of this rank: equal shares rounded up to 32 strings."""
    per = (n_strings + world - 1) // world
    per = (per + 31) // 32 * 32
    # lo stays on the 32-string grid even for a rank whose shard is empty (n_strings < 32 * world):
    # merge_match_bits places shards by lo // 32
    lo = rank * per
    hi = max(lo, min(n_strings, lo + per))
    return lo, hi


def merge_match_bits(local_bits, lo, n_total, group=None):
    """All-reduce the packed match bitmap.

    local_bits: int32 tensor with ceil((hi-lo)/32) words for strings [lo, hi).
    Returns the full ceil(n_total/32)-word bitmap, identical on every rank.
    """
    import torch
    import torch.distributed as dist
    assert lo % 32 == 0
    words = (n_total + 31) // 32
    full = torch.zeros(words, dtype=torch.int32, device=local_bits.device)
    if local_bits.numel():                      # an empty shard contributes nothing but still joins the collective
        full[lo // 32: lo // 32 + local_bits.numel()] = local_bits
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return full


def popcount_bits(bits):
    """Number of set bits of an int32 bitmap tensor (matches)."""
    import torch
    v = bits.view(torch.uint8)
    # 8-bit popcount table
    table = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device=bits.device)
    return int(table[v.long()].sum().item())
