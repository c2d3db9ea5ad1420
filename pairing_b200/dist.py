"""Multi-GPU: one process per GPU (torch.distributed), batches sharded contiguously.

Independent batches (pairing, wNAF, normalisation) need no collective: every rank works on
`shard_range(n, rank, world)` and the results are gathered by the caller if wanted.

The ONE exchange step of the path is the sharded multi-pairing (`Engine::miller_loop` over many
pairs, bls12_381/mod.rs:80-95): each rank reduces its shard to one Fq12 partial product (576 B),
the partials are all-gathered (NCCL on GPUs; 576 B per rank, latency-bound), every rank multiplies
them in rank order and a single final exponentiation follows.  The product is a field value, so
any partition and order gives the reference's bits.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of rank `rank` among `world` ranks."""
    return n * rank // world, n * (rank + 1) // world


def all_gather_partials(partial, group=None):
    """partial: (1, 72) int64 tensor (CUDA for nccl, CPU for gloo) -> (world, 72)."""
    world = dist.get_world_size(group)
    out = torch.empty((world, partial.shape[1]), dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(out, partial.contiguous(), group=group)
    return out


def pairing_product_sharded(eng, p_shard, q_shard, group=None):
    """final_exponentiation(miller_loop(ALL pairs of all ranks)), BASELINE configs[2] in its stated (strong-scaling) form.
    eng: DeviceEngine.  Per rank: k_pair_multi_miller + k_pair_product_tail -> one 576-byte partial; all-gather; then the
    tail kernel once more over the `world` partials with the single final exponentiation (every rank computes it: it is one
    warp's work and saves a broadcast).  -> ((1, 72), is_some)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return eng.pairing_product(p_shard, q_shard)
    part = eng.multi_miller_loop(p_shard, q_shard)
    return eng.fq12_product_tail(all_gather_partials(part, group), final_exp=True)


def multi_miller_loop_sharded(local_product, merge, p_shard, q_shard, group=None):
    """local_product(p, q) -> (1,72) partial of this rank's shard; merge(partials (world,72)) -> (1,72).
    With a DeviceEngine: local_product = eng.multi_miller_loop, merge = eng.fq12_product."""
    part = local_product(p_shard, q_shard)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return part
    return merge(all_gather_partials(part, group))
