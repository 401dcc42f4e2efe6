"""Data-parallel sharding of a parameter batch over the GPUs of one node.

Samples are independent, so the batch is split into contiguous slices, one per rank, with no
data-path collective; the only collective is the final gather of the [n, nb, 3] results
(NCCL over NVLink when the tensors are on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world_size, rank):
    """Contiguous slice [lo, hi) of rank `rank`: sizes differ by at most one sample."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(local, n_total, group=None, dst=None):
    """Concatenate per-rank [n_r, nb, 3] results in rank order.

    dst=None: every rank gets the full [n_total, nb, 3] tensor (all_gather);
    dst=r:    only rank r gets it (others get None).
    Shard sizes differ by at most one sample, so each rank contributes a buffer padded to the
    largest shard and the padding row is dropped after the collective."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    tail = tuple(local.shape[1:])
    bounds = [shard_bounds(n_total, world, r) for r in range(world)]
    mine = bounds[rank][1] - bounds[rank][0]
    if local.shape[0] != mine:
        raise ValueError("local shard has the wrong number of samples for this rank")
    cap = max(hi - lo for lo, hi in bounds)
    if mine == cap:
        send = local.contiguous()
    else:
        send = torch.zeros((cap,) + tail, dtype=local.dtype, device=local.device)
        send[:mine] = local
    even = all(hi - lo == cap for lo, hi in bounds)
    want = dst is None or rank == dst
    recv = torch.empty((world * cap,) + tail, dtype=local.dtype, device=local.device) if want else None
    if dst is None:
        dist.all_gather_into_tensor(recv, send, group=group)
    else:
        parts = list(recv.chunk(world, dim=0)) if want else None
        dist.gather(send, gather_list=parts, dst=dst, group=group)
    if not want:
        return None
    if even:
        return recv
    return torch.cat([recv[r * cap:r * cap + (hi - lo)] for r, (lo, hi) in enumerate(bounds)], dim=0)


def run_batch_sharded(params, sensor, precision="fp64", group=None, dst=None, compute=None):
    """Every rank holds (or can index) the full [27, n] block; each computes its slice and
    the results are gathered.  `compute(params_slice, sensor, precision)` defaults to the
    CUDA path; tests inject a stand-in to exercise the sharding/gather logic without a GPU."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = params.shape[1]
    lo, hi = shard_bounds(n, world, rank)
    if compute is None:
        from .batch import run_batch_params
        compute = lambda p, s, pr: run_batch_params(p, s, pr)
    local = compute(params[:, lo:hi].contiguous(), sensor, precision)
    return gather_results(local, n, group=group, dst=dst)
