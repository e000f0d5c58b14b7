"""Sample-index partition of one frame across the GPUs of a box.

One process per GPU (torchrun / torch.distributed). Every rank holds a scene replica and renders ITS samples of EVERY
pixel into its own fp32 accumulation buffer; Philox is keyed (seed, pixel, sample), so the union of the ranks' sample
ranges is the same set of paths a single GPU would trace. The only exchange step of the path is one sum-reduce of the
W*H*3 accumulation buffers onto rank 0 (NCCL over NVLink on GPUs; gloo in the CPU tests), after which rank 0 resolves
(divide by the total sample count, tone-map) exactly as the single-GPU path does. No scene sharding, no ray forwarding.
"""


def sample_range(rank, world, spp):
    """Contiguous block of sample indices [begin, end) of `rank`; the remainder goes to the first ranks, so the blocks
    tile [0, spp) exactly and differ in size by at most one sample."""
    if world <= 0 or not 0 <= rank < world or spp < 0:
        raise ValueError(f"bad partition rank={rank} world={world} spp={spp}")
    base, rem = divmod(spp, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class PartitionedFrame:
    """Drives one context per rank. `ctx` needs clear(), render(W,H,begin,end,bounces,seed=,flags=),
    set_samples_accumulated(n) (narvalengine_b200.engine.Context has them); `accum` is the rank's accumulation buffer
    as a torch tensor aliasing the library's device memory (see `alias_accum`); `dist` is torch.distributed or None."""

    def __init__(self, ctx, accum, rank=0, world=1, dist=None, group=None):
        self.ctx, self.accum, self.rank, self.world, self.dist, self.group = ctx, accum, rank, world, dist, group

    def render(self, W, H, spp_total, bounces, seed=1, flags=0):
        begin, end = sample_range(self.rank, self.world, spp_total)
        self.ctx.clear()
        self.ctx.render(W, H, begin, end, bounces, seed=seed, flags=flags)
        if self.world > 1:
            self.dist.reduce(self.accum, dst=0, op=self.dist.ReduceOp.SUM, group=self.group)
        if self.rank == 0:
            self.ctx.set_samples_accumulated(spp_total)
        return begin, end


def alias_accum(ctx, device_index):
    """The context's accumulation buffer (ne_b200_accum_buffer) as a torch tensor, without a copy."""
    import torch
    ptr, nfloat, _ = ctx.accum_buffer()

    class _Alias:
        __cuda_array_interface__ = {"shape": (nfloat,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Alias(), device=f"cuda:{device_index}")
