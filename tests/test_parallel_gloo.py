"""N>1 host logic on CPU (gloo, world_size 2): sharding + loss normalisation + the single flat all-reduce reproduce the
defined parity target: the oracle evaluated per shard with shared weights, gradients summed."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, kind, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import srl_oracle as O
    from srl_zoo_b200 import parallel
    gbs = 4
    obs, nobs, actions = O.synthetic_batch(gbs, seed=1234)
    g = torch.Generator().manual_seed(7)
    eps = (torch.randn(gbs, 200, generator=g), torch.randn(gbs, 200, generator=g))
    lo, hi = parallel.shard_slice(gbs, rank, world)
    sd = O.build_state("vae" if kind == "vae" else "ae", 200, 6, seed=1)
    P, B = O.split_state(sd)
    # per-shard step with the mean-type terms normalised by the GLOBAL count: weight scaled by local/global
    frac = (hi - lo) / gbs
    weights = {"autoencoder": 1.0 * frac} if kind == "ae" else None   # mean over global elements = frac * local mean
    r = O.train_step(kind, P, B, obs[lo:hi], nobs[lo:hi], actions[lo:hi], eps[0][lo:hi], eps[1][lo:hi], weights=weights)
    names = list(P.keys())
    flat = torch.cat([(P[k].grad if P[k].grad is not None else torch.zeros_like(P[k])).reshape(-1) for k in names])
    # engine-side coefficient agrees with the oracle-side weighting
    coef = parallel.mse_coef(kind, 1.0 if kind == "ae" else 0.5e-6, gbs)
    local_coef = (2.0 * weights["autoencoder"] / ((hi - lo) * parallel.N_PIX)) if kind == "ae" else 2.0 * 0.5e-6
    assert abs(coef - local_coef) <= 1e-12 * abs(coef)
    tail = torch.tensor([sum(r["losses"].values())])
    buf = torch.cat([flat, tail])
    parallel.allreduce_flat(buf, world)
    if rank == 0:
        ret["flat"] = buf.clone()
        ret["names"] = names
    dist.destroy_process_group()


def _reference(kind):
    """single process: per-shard BN statistics (shared weights), gradients summed with the same normalisation"""
    sys.path.insert(0, ROOT)
    from oracle import srl_oracle as O
    gbs, world = 4, 2
    obs, nobs, actions = O.synthetic_batch(gbs, seed=1234)
    g = torch.Generator().manual_seed(7)
    eps = (torch.randn(gbs, 200, generator=g), torch.randn(gbs, 200, generator=g))
    total = None
    for rank in range(world):
        lo, hi = rank * 2, rank * 2 + 2
        sd = O.build_state("vae" if kind == "vae" else "ae", 200, 6, seed=1)
        P, B = O.split_state(sd)
        weights = {"autoencoder": 0.5} if kind == "ae" else None
        O.train_step(kind, P, B, obs[lo:hi], nobs[lo:hi], actions[lo:hi], eps[0][lo:hi], eps[1][lo:hi], weights=weights)
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in P.values()])
        total = flat if total is None else total + flat
    return total


def _run(kind, port):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, kind, ret), nprocs=2, join=True)
    got = ret["flat"][:-1]
    want = _reference(kind)
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err < 1e-5, err  # fp32 summation order (thread-split reductions differ between 1 and 2 processes)


def test_shard_slice_and_coefficients():
    sys.path.insert(0, ROOT)
    from srl_zoo_b200 import parallel
    assert parallel.shard_slice(1024, 3, 8) == (384, 512)
    assert abs(parallel.mse_coef("ae", 1.0, 256) * 256 * parallel.N_PIX - 2.0) < 1e-12
    assert parallel.mse_coef("vae", 0.5e-6, 1024) == 1e-6
    assert parallel.recon_scale("vae", 64) == 1.0


def test_two_rank_allreduce_matches_per_shard_oracle_ae():
    _run("ae", 29541)


def test_two_rank_allreduce_matches_per_shard_oracle_vae():
    _run("vae", 29542)
