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


class Comm:
    """One rank's communicator for pire_gpu_run_sharded (include/pire_b200.h): NCCL behind the C ABI.

    ``torch.distributed`` (any backend; gloo works) only ships rank 0's 128-byte NCCL id to the other ranks --
    the scan and the exchange themselves never touch PyTorch."""

    def __init__(self, device, rank=None, world=None, group=None):
        import ctypes as C
        import torch.distributed as dist
        from . import _native as N
        self._N = N
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        ident = (C.c_uint8 * N.COMM_ID_BYTES)()
        if self.rank == 0:
            N.check(N.lib.pire_gpu_comm_get_id(ident), "pire_gpu_comm_get_id")
        box = [bytes(ident)]
        if self.world > 1:
            dist.broadcast_object_list(box, src=0, group=group)
        ident = (C.c_uint8 * N.COMM_ID_BYTES).from_buffer_copy(box[0])
        self._h = C.c_void_p()
        N.check(N.lib.pire_gpu_comm_create(ident, self.world, self.rank, self.device, C.byref(self._h)), "pire_gpu_comm_create")

    def close(self):
        if getattr(self, "_h", None):
            self._N.lib.pire_gpu_comm_destroy(self._h)
            self._h = None

    __del__ = close

    def bounds(self, n_global):
        return native_shard_bounds(n_global, self.rank, self.world)

    def words(self, n_global):
        return int(self._N.lib.pire_gpu_sharded_words(n_global, self.world))

    def run_sharded(self, sc, batch, n_global, flags, bits_all, masks=None, states=None, async_exchange=False, stream=None):
        """Scan this rank's shard (`batch`) and gather every rank's slot of the int32 CUDA tensor `bits_all`
        (self.words(n_global) words)."""
        import torch
        N = self._N
        if stream is None:
            stream = torch.cuda.current_stream(batch.device).cuda_stream
        ptr = lambda t: None if t is None else t.data_ptr()
        N.check(N.lib.pire_gpu_run_sharded(sc._h, self._h, batch.corpus.data_ptr(), ptr(batch.offsets), batch.fixed_len, n_global,
                                           flags | (N.RUN_ASYNC_EXCHANGE if async_exchange else 0), bits_all.data_ptr(),
                                           ptr(masks), ptr(states), stream), "pire_gpu_run_sharded")

    def gather_bits(self, n_global, bits_all, async_exchange=False, stream=None):
        import torch
        N = self._N
        if stream is None:
            stream = torch.cuda.current_stream(bits_all.device).cuda_stream
        N.check(N.lib.pire_gpu_comm_gather_bits(self._h, n_global, bits_all.data_ptr(), N.RUN_ASYNC_EXCHANGE if async_exchange else 0,
                                                stream), "pire_gpu_comm_gather_bits")

    def wait(self, device=None, stream=None):
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(device).cuda_stream
        self._N.check(self._N.lib.pire_gpu_comm_wait(self._h, stream), "pire_gpu_comm_wait")


def native_shard_bounds(n_strings, rank, world):
    """pire_gpu_shard_bounds of the C ABI (the same split as shard_bounds above)."""
    import ctypes as C
    from . import _native as N
    lo, hi = C.c_uint64(0), C.c_uint64(0)
    N.lib.pire_gpu_shard_bounds(n_strings, world, rank, C.byref(lo), C.byref(hi))
    return int(lo.value), int(hi.value)
