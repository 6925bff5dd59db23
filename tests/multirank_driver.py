"""One rank of the 2-GPU parity check (launched by tests/test_gpu_multirank.py through torch.distributed.run, NCCL):
TrainStep(world_size=2) on this rank's contiguous shard of a global minibatch; rank 0 then checks the all-reduced flat gradient,
the per-loss scalars and the post-Adam parameters against the defined parity target: the oracle evaluated per shard
with shared weights, mean-type losses normalised by the GLOBAL count, gradients summed."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    kind, losses = sys.argv[1], sys.argv[2].split(",")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import helpers as H
    import srl_zoo_b200
    from oracle import srl_oracle as O
    from srl_zoo_b200 import parallel
    H.tf32_off()
    gbs = 6
    per = gbs // world
    mod, _, _ = H.make_pair(kind, losses, device=dev)
    cpu, _ = H.inputs(gbs, device=dev)
    lo, hi = parallel.shard_slice(gbs, rank, world)
    sl = lambda t: t[lo:hi].contiguous().to(dev)
    eng = srl_zoo_b200.TrainStep(mod, per, lr=1e-6, world_size=world)
    t = eng.step(sl(cpu["obs"]), sl(cpu["nobs"]), sl(cpu["actions"]), sl(cpu["eps"][0]), sl(cpu["eps"][1]),
                 torch.from_numpy(cpu["rects"][0][lo:hi]).to(dev), torch.from_numpy(cpu["rects"][1][lo:hi]).to(dev))
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().clone() for n, p in mod.named_parameters()}
    # every rank holds the same all-reduced gradient and therefore the same parameters
    flat = eng.flat_g.detach().clone()
    other = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(other, flat)
    assert all(torch.equal(o, other[0]) for o in other)
    ok = True
    if rank == 0:
        total, losses_sum = None, {}
        for r in range(world):
            a, b = parallel.shard_slice(gbs, r, world)
            P, B = H.oracle_state(kind, device=dev)
            frac = (b - a) / gbs
            # mean-type terms normalised by the global count = frac * the shard's own mean; sum-type terms (VAE) left alone
            w = {"autoencoder": frac, "dae": frac, "forward": frac, "inverse": 2.0 * frac}
            x = dict(obs=cpu["obs"][a:b].to(dev), nobs=cpu["nobs"][a:b].to(dev), actions=cpu["actions"][a:b].to(dev),
                     eps=(cpu["eps"][0][a:b].to(dev), cpu["eps"][1][a:b].to(dev)), rects=(cpu["rects"][0][a:b], cpu["rects"][1][a:b]))
            res = O.train_step(kind, P, B, x["obs"], x["nobs"], x["actions"], x["eps"][0], x["eps"][1], x["rects"][0], x["rects"][1],
                               use_forward="forward" in losses, use_inverse="inverse" in losses, weights=w)
            for n, v in res["losses"].items():
                mean_type = n in ("reconstruction_loss", "forward_loss", "inverse_loss")
                losses_sum[n] = losses_sum.get(n, 0.0) + v * (frac if mean_type else 1.0)
            g = {k: (p.grad.detach() if p.grad is not None else torch.zeros_like(p)) for k, p in P.items()}
            total = g if total is None else {k: total[k] + g[k] for k in g}
        for i, n in enumerate(eng.loss_names()):
            if n:
                e = abs(t[i].item() - losses_sum[n]) / abs(losses_sum[n])
                print("loss %-20s engine %.6f  per-shard oracle %.6f  rel %.1e" % (n, t[i].item(), losses_sum[n], e))
                ok &= e < 1e-5
        noise = ("model.decoder_conv.0.bias", "model.decoder_conv.3.bias", "model.decoder_conv.6.bias", "model.decoder_conv.9.bias")
        worst = 0.0
        for k, g in total.items():
            if k in noise or g.abs().max().item() == 0.0:
                continue
            cos, err = H.cosine(grads[k], g), H.rel_err(grads[k], g)
            worst = max(worst, err)
            if not (cos > 0.9999 and err < 5e-2):
                print("GRAD MISMATCH", k, cos, err)
                ok = False
        print("worst gradient rel err vs the per-shard oracle sum: %.2e" % worst)
        print("MULTIRANK_OK" if ok else "MULTIRANK_FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
