"""Host-side data-parallel logic of the train step: one process per GPU, contiguous shards of the
global minibatch, ONE all-reduce (sum) over the flat gradient buffer per step.

Loss normalisation decides what is summed: mean-type terms (reconstruction_loss, forward_loss, inverse_loss:
losses/losses.py:126,181) are normalised by the GLOBAL element count on every rank; sum-type terms (generation_loss,
kl_loss: losses/losses.py:210-211,253-254) are left alone; then ncclSum.  BatchNorm statistics stay per-rank
(the parity target for W > 1 is the oracle evaluated per shard with shared weights), running statistics stay
rank-local (rank 0's are the ones saved, DDP semantics)."""
import torch

N_PIX = 3 * 224 * 224


def shard_slice(global_batch, rank, world):
    """contiguous pairs [lo, hi) of the global minibatch owned by `rank`"""
    if global_batch % world != 0:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


def mse_coef(kind, weight, global_batch):
    """d(total)/d(decoded) = coef * (decoded - target)   (AE/DAE: mean over global elements; VAE: plain sum)"""
    if kind == "vae":
        return 2.0 * weight
    return 2.0 * weight / (global_batch * N_PIX)


def recon_scale(kind, global_batch):
    """factor turning this rank's squared-error SUM into its share of the logged loss value"""
    return 1.0 if kind == "vae" else 1.0 / (global_batch * N_PIX)


def allreduce_flat(flat, world, group=None):
    """the single collective of the step: sum over ranks of the flat gradient (+ per-loss scalar tail) buffer"""
    if world > 1:
        torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM, group=group)
    return flat
