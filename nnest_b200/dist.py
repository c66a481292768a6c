"""One-process-per-GPU plumbing (torch.distributed; NCCL over NVLink on the GPU box, gloo in CPU tests).

Replaces the reference's mpi4py pickle collectives (nnest/sampler.py:165-177; nnest/nested.py:199-226,416-427):
  * chains are sharded by rank -- rank r owns global chain ids [r*n, (r+1)*n), which key the Philox streams, so with a
    FIXED step size a chain's trajectory does not depend on how many GPUs the batch is split over (with
    mcmc_dynamic_step_size=True the scale adapts on each rank's own accept counts, as under the reference's MPI mode,
    so trajectories then depend on the split);
  * after a refill only the end states (start point, end point, end loglike: all that nested.py:432-437 reads)
    are all-gathered, in rank order = the reference's np.concatenate order;
  * flow weights are broadcast from rank 0 as one flat buffer after every (re)training.
There is no data-path collective inside the MCMC steps themselves.
"""
import numpy as np
import torch
import torch.distributed as dist


def ensure_initialized():
    """The reference probes MPI by itself when a sampler is built (nnest/sampler.py:165-177).  The equivalent here: a
    process started by torchrun (RANK / WORLD_SIZE / LOCAL_RANK in the environment, world size > 1) joins the default
    process group -- NCCL with this rank's GPU, gloo without CUDA -- unless the caller has already done so."""
    import os
    if not dist.is_available() or dist.is_initialized():
        return
    try:
        world = int(os.environ.get('WORLD_SIZE', '1'))
    except ValueError:
        world = 1
    if world <= 1 or 'RANK' not in os.environ:
        return
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29500')
    if torch.cuda.is_available():
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    else:
        dist.init_process_group('gloo')
    import atexit
    atexit.register(_shutdown)      # the group was created here, so it is torn down here


def _shutdown():
    try:
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def chain_offset(n_chains_per_rank):
    return rank_world()[0] * n_chains_per_rank


def allgather_rows(t):
    """Concatenate every rank's tensor along dim 0 in rank order (same shape on every rank)."""
    if not is_distributed():
        return t
    t = t.contiguous()
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    return torch.cat(parts, dim=0)


def shard_bounds(n, rank, world):
    """Contiguous share [lo, hi) of n items for `rank`: the first n % world ranks get one extra."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allgather_ragged(a, n_total, device):
    """Rank-order concatenation of 1-D float64 numpy shares cut with shard_bounds (lengths differ by at most one)."""
    if not is_distributed():
        return a
    world = dist.get_world_size()
    width = (int(n_total) + world - 1) // world
    buf = torch.zeros((width,), dtype=torch.float64, device=device)
    buf[:len(a)] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        out.append(parts[r][:hi - lo].cpu().numpy())
    return np.concatenate(out)


def broadcast_array(a, device, src=0):
    """Broadcast a numpy array from `src`; every rank passes an array of the right shape/dtype."""
    if not is_distributed():
        return a
    t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def broadcast_parameters(module, src=0):
    """Replicate `module`'s parameters from `src` with ONE collective on a flat buffer."""
    if not is_distributed():
        return
    params = list(module.parameters())
    flat = torch.nn.utils.parameters_to_vector(params).detach().clone()
    dist.broadcast(flat, src=src)
    off = 0
    with torch.no_grad():      # in place: captured CUDA graphs hold the parameter addresses
        for p in params:
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))
            off += n


def allreduce_sum_int(v, device):
    if not is_distributed():
        return int(v)
    t = torch.tensor([int(v)], dtype=torch.int64, device=device)
    dist.all_reduce(t)
    return int(t.item())
