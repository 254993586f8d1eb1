"""Channel sharding of the ADMM iteration over GPUs (one process per GPU).

The y-update of UniRes is an independent SPD system per channel
(unires/_update.py:122-150: the `lhs` closure captures x[c], y[c] only), so
channels shard one-per-GPU with NO collective inside the CG loop.  Channels
couple only through the joint-TV shrink (unires/_update.py:166-173) and the
objective (unires/_update.py:417-425): per ADMM iteration one SUM all-reduce
of an (X,Y,Z) float32 field for the prox, one for the prior energy and one
float64 scalar for the data term.  Everything here is backend-agnostic
(`torch.distributed` with nccl on the GPU box, gloo in the CPU tests); the
per-rank arithmetic is injected as callables.
"""
import os

import torch


def channel_shard(n_channels, world_size, rank):
    """Channels owned by `rank`: round-robin c -> c mod world_size."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError('bad rank/world_size')
    if world_size > n_channels:
        # a rank without channels could not take part in the kernels between the all-reduces;
        # every rank evaluates this with the same arguments, so all of them raise together
        raise ValueError('channel_shard: %d ranks for %d channels (use at most one rank per '
                         'channel)' % (world_size, n_channels))
    return [c for c in range(n_channels) if c % world_size == rank]


def owner(channel, world_size):
    return channel % world_size


def env_rank():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')),
            int(os.environ.get('WORLD_SIZE', '1')))


def _dist():
    import torch.distributed as dist
    return dist


def is_distributed(group=None):
    dist = _dist()
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def all_reduce_sum(t, group=None):
    if is_distributed(group):
        _dist().all_reduce(t, group=group)
    return t


def coupled_prox(field, norm2_local, apply_local, group=None):
    """JTV prox across ranks: field <- sum over ALL channels of |u_c|^2, then apply.

    norm2_local(field): overwrite `field` with this rank's sum_c |u_c|^2
    apply_local(field): z, w update of this rank's channels given the global field"""
    norm2_local(field)
    all_reduce_sum(field, group)
    apply_local(field)
    return field


def coupled_objective(row, field, data_and_prior_local, sqrt_sum, group=None):
    """Objective across ranks.  data_and_prior_local(row, field) writes this rank's data term
    into row[1] and its prior energy field into `field`; sqrt_sum(field) -> sum sqrt(field)."""
    data_and_prior_local(row, field)
    all_reduce_sum(field, group)
    all_reduce_sum(row[1:2], group)
    row[2] = sqrt_sum(field)
    row[0] = row[1] + row[2]
    return row


def coupled_objective_and_prox(row, fields, data_and_prior_local, norm2_local, sqrt_sum,
                               apply_local, group=None):
    """Objective and JTV prox of one ADMM iteration with ONE field all-reduce.

    fields: (2, X, Y, Z).  fields[0] receives this rank's prior-energy field
    (data_and_prior_local(row, fields[0]) also writes the local data term into row[1]) and
    fields[1] its sum_c |u_c|^2 (norm2_local(fields[1])); both depend only on the freshly solved
    y and on z, w, so they are formed back to back and summed across ranks in a single
    all-reduce of 2 N floats (one launch / ring set-up instead of two, better bus utilisation),
    followed by the float64 scalar.  Then sqrt_sum(fields[0]) -> row[2] and apply_local(fields[1])."""
    data_and_prior_local(row, fields[0])
    norm2_local(fields[1])
    all_reduce_sum(fields, group)
    all_reduce_sum(row[1:2], group)
    row[2] = sqrt_sum(fields[0])
    row[0] = row[1] + row[2]
    apply_local(fields[1])
    return row
