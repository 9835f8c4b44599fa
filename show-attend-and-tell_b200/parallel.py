"""Multi-GPU host logic of the decode path: one process per GPU, images (batch rows) are the
independent units so inference shards the batch with NO data-path collective;
torch.distributed (nccl on GPUs, gloo in CPU tests) only carries the timing max and result
gathers.  The reference's own "distributed" mode (main_distributed.py, async parameter server)
is non-functional and is not mirrored."""
import os


def shard_range(n, rank, world):
    """Contiguous shard [lo, hi) of n images for `rank` of `world` (sizes differ by at most 1)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (RANK/WORLD_SIZE/MASTER_*)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_rank_world()
    if world == 1 or dist.is_initialized():
        return rank, local_rank, world
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend, rank=rank, world_size=world,
                                device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (timing of a sharded step = slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_shards(local, n_total, device=None):
    """Reassemble per-rank result rows (e.g. tokens [n_local, T]) in image order on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)


class StepCollective(object):
    """The collective protocol of one data-parallel training step (model.train_step), on any torch device.

    flat = [gradients (n_grad floats) | ce, accuracy, attention sums | sum of the NEXT batch's masks | pad]: ONE
    all-reduce per step.  The whole-batch mask sum a step needs BEFORE its backward pass (model.py:316-318 divides by
    it) therefore arrived with the previous step's collective; only the first step of a run, or a batch whose masks
    were not announced, pays a separate 8-byte all-reduce.  Everything stays on the device and in stream order."""

    def __init__(self, flat, n_grad):
        import torch
        self.flat, self.n_grad = flat, int(n_grad)
        self.tail = flat[self.n_grad:]
        assert self.tail.numel() >= 4
        self.mask_sum = torch.zeros(1, dtype=torch.float64, device=flat.device)   # global sum for the batch `key` names
        self.key = None
        self.collectives = 0

    @staticmethod
    def _key(masks):
        return (masks.data_ptr(), masks._version, tuple(masks.shape))

    def global_mask_sum(self, masks):
        """Device scalar (float64 [1]): sum of `masks` over all ranks.  No collective if the previous reduce() carried
        it: either announced explicitly (`next_masks`, a promise about this call's masks: only the shape is checked)
        or the very same tensor is fed again."""
        import torch
        import torch.distributed as dist
        if self.key not in (self._key(masks), ("announced", tuple(masks.shape))):
            self.mask_sum.copy_(masks.sum(dtype=torch.float64).reshape(1))
            dist.all_reduce(self.mask_sum)
            self.collectives += 1
        self.key = self._key(masks)
        return self.mask_sum

    def reduce(self, shard_losses, next_masks, announced=False):
        """After the backward pass filled flat[:n_grad]: sums gradients and the three shard-additive losses over the
        ranks and hands over the next batch's mask sum (announced=True: `next_masks` are the masks the NEXT step will
        be called with, whatever tensor object they arrive in).  Returns the device tensor [ce, accuracy, attention]."""
        import torch.distributed as dist
        self.tail[:3].copy_(shard_losses[:3])
        self.tail[3:4].copy_(next_masks.sum().reshape(1))    # (exact in fp32: at most batch x steps ones per rank)
        dist.all_reduce(self.flat)
        self.collectives += 1
        self.mask_sum.copy_(self.tail[3:4])
        self.key = ("announced", tuple(next_masks.shape)) if announced else self._key(next_masks)
        return self.tail[:3]
