#!/usr/bin/env python
"""
bench.py -- images/sec of the conv-AE / VAE train step (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5            # this repo's CUDA path (hand-written sm_100a kernels)
    python bench.py --impl reference --steps 5 --warmup 1     # the reference's arithmetic on the host CPU (oracle port)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU, weak scaling

One "step" = one training minibatch of SRL4robotics.learn (models/learner.py:373-497): zero-grad, forward(obs),
forward(next_obs), losses, backward, (one all-reduce of the flat gradient buffer), Adam, per-loss scalars.
An "image" is one 224x224x3 observation through encoder+decoder forward and backward; a minibatch of bs pairs is
2*bs images.  value = 2 * bs_global / step_time, inputs resident in HBM; e2e = the same through TrainStep.step_host
(pinned host buffers -> H2D every step, loss scalars D2H every step).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG, S, A = 224, 200, 6
# exact per-image MAC counts of the instantiated reference modules, x2 = FLOP
FWD_MACS = {"enc0": 118013952, "enc4": 115605504, "enc8": 7225344, "dec0": 1327104, "dec3": 6230016, "dec6": 26873856,
            "dec9": 111513600, "dec12": 37850112}
CONV_TRAIN_FLOP_PER_IMAGE = 2311809024   # fwd + wgrad (all) + dgrad (all but enc0), conv/convT layers
ALG_BYTES_PER_IMAGE_FP32 = 44.7e6        # perfect-fusion lower bound, fp32 activations
# elements moved per image by the elementwise / reduction call sites (read + write), fp32
EW_ELEMS = {"bn_relu_pool.fwd": 802816 + 200704 + 200704 + 46656 + 12544 + 2304,
            "pool.bwd_stats": 2 * (200704 + 46656 + 2304),                                    # pooled-side sums: dpool, a in
            "pool.bwd": 2 * (802816 + 200704 + 12544) + (200704 + 46656 + 2304),              # one full-size pass: y, dpool in; dy out
            "bn.bwd": 3 * (10816 + 46656 + 193600 + 788544)}   # decoder stages: dz, y in; dy out

# algorithmic (minimum) HBM bytes per image of the conv call sites: every operand read once, every result written once, fp32
_T = lambda e, c=64: e * e * c * 4
_X, _Y1, _A1, _A2, _Y3 = _T(224, 3), _T(112), _T(56), _T(27), _T(14)
_D0, _Y4, _Y5, _Y6, _Y7 = _T(6), _T(13), _T(27), _T(55), _T(111)
SITE_BYTES = {
    "enc0.fwd": _X + _Y1, "enc0.wgrad": _X + _Y1,
    "enc4.fwd": _A1 + _A1, "enc4.dgrad": _A1 + _A1, "enc4.wgrad": _A1 + _A1,
    "enc8.fwd": _A2 + _Y3, "enc8.dgrad": _A2 + _Y3, "enc8.wgrad": _A2 + _Y3,
    "dec0.fwd": _D0 + _Y4, "dec0.dgrad": _D0 + _Y4, "dec0.wgrad": _D0 + _Y4,
    "dec3.fwd": _Y4 + _Y5, "dec3.wgrad": _Y4 + _Y5, "dec3.dgrad": _Y5 + 2 * _Y4,   # dgrad epilogue: + pre-BN activation (ReLU mask, BN sums)
    "dec6.fwd": _Y5 + _Y6, "dec6.wgrad": _Y5 + _Y6, "dec6.dgrad": _Y6 + 2 * _Y5,
    "dec9.fwd": _Y6 + _Y7, "dec9.wgrad": _Y6 + _Y7, "dec9.dgrad": _Y7 + 2 * _Y6,
    "dec12.fwd": _Y7 + 2 * _X, "dec12.wgrad": _Y7 + 2 * _X, "dec12.dgrad": 2 * _X + 2 * _Y7,   # decoded + target; y7 (mask) + dz7
}


def site_roofline(site, cnt, tot_ms, bs, steps, pk, traffic_tab):
    """Roofline entry of one call site from its CUDA-event time: achieved = algorithmic units of one launch / average launch
    duration.  Conv sites are measured against both roofs (2*MACs against the bf16 tensor peak, minimum bytes against the HBM
    peak) and `bound` names the one they sit closer to; elementwise sites only have the HBM roof."""
    avg_ms = tot_ms / cnt
    layer, _, what = site.partition(".")
    traffic = traffic_tab.get(site, {}).get("dram_bytes_per_launch") if bs == 256 else None
    if layer in FWD_MACS and what in ("fwd", "dgrad", "wgrad", "bwd"):
        mult = 2 if what == "bwd" else 1  # dec12.bwd (SIMT scaffold) = dgrad + wgrad
        tf = 2.0 * FWD_MACS[layer] * mult * bs / (avg_ms * 1e-3) / 1e12
        t_roof = {"bound": "tensor", "achieved": tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": tf / pk["bf16_sustained"]}
        if site not in SITE_BYTES:
            best, other = t_roof, None
        else:
            gb = SITE_BYTES[site] * bs / (avg_ms * 1e-3) / 1e9
            h_roof = {"bound": "hbm", "achieved": gb, "peak": pk["hbm"], "unit": "GB/s", "frac": gb / pk["hbm"]}
            best, other = (h_roof, t_roof) if h_roof["frac"] >= t_roof["frac"] else (t_roof, h_roof)
        out = {"kernel": site}
        out.update(best)
        out.update({"traffic": traffic, "avg_launch_ms": avg_ms, "other_roof": other})
        return out
    if site not in EW_ELEMS:
        return None
    # an elementwise call site covers several launches of different sizes per model call: EW_ELEMS is the per-image
    # total over all of them, so bytes per launch = per-step bytes / launches per step (same ratio as bytes / time)
    elems = EW_ELEMS[site] * 2 * bs / (cnt / steps)
    ach = elems * 4 / (avg_ms * 1e-3) / 1e9
    return {"kernel": site, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
            "traffic": traffic, "avg_launch_ms": avg_ms, "other_roof": None}


CONFIGS = {
    "ae": dict(losses=["autoencoder"], bs=256, name="conv autoencoder (models/autoencoders.py), 224x224x3, state-dim 200, bs=256/GPU"),
    "vae": dict(losses=["vae"], bs=128, name="beta-VAE (models/vae.py, beta=1), 224x224x3, state-dim 200, bs=128/GPU"),
    "dae": dict(losses=["dae"], bs=256, name="denoising autoencoder (dae) with zero-pixel mask, 224x224x3, bs=256/GPU"),
    "ae_fwd_inv": dict(losses=["autoencoder", "inverse", "forward"], bs=128, name="autoencoder+inverse+forward, 224x224x3, bs=128/GPU"),
}


def kind_of(losses):
    return "vae" if "vae" in losses else ("dae" if "dae" in losses else "ae")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic(bs, seed, device=None, pin=False):
    """uint8 U{0..255} -> /255, ImageNet mean/std, (B,3,224,224) fp32; actions U{0..5}."""
    import torch
    g = torch.Generator().manual_seed(seed)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)

    def one():
        u8 = torch.randint(0, 256, (bs, 3, IMG, IMG), generator=g, dtype=torch.uint8)
        t = ((u8.float() / 255.0) - mean) / std
        return t.pin_memory() if pin else t

    obs, nobs = one(), one()
    actions = torch.randint(0, A, (bs, 1), generator=g, dtype=torch.int64)
    return obs, nobs, actions


def cpu_oracle_rate(losses, bs, steps, warmup, threads=0):
    """images/s of the reference's arithmetic (oracle port: same torch CPU ops as the reference modules) on the host.
    threads=0: probe {8,16,32,64,all cores} with one step each and keep the fastest (torch's CPU convs stop scaling
    long before 128 threads at these batch sizes); the count actually used is reported as `cores`."""
    import numpy as np
    import torch
    from oracle import srl_oracle as O
    ncpu = os.cpu_count() or 1
    kind = kind_of(losses)
    obs, nobs, actions = O.synthetic_batch(bs, seed=1234)
    sd = O.build_state("vae" if kind == "vae" else "ae", S, A, seed=1)
    P, B = O.split_state(sd)
    opt = O.Adam(P, lr=0.005)
    rng = np.random.RandomState(1)
    rects = (O.sample_rects(bs, rng=rng), O.sample_rects(bs, rng=rng))

    def one():
        t0 = time.perf_counter()
        O.train_step(kind, P, B, obs, nobs, actions, None, None, rects[0], rects[1], use_forward="forward" in losses,
                     use_inverse="inverse" in losses, optimizer=opt)
        return time.perf_counter() - t0

    if threads <= 0:
        best = (None, 1e30)
        for c in sorted({min(c, ncpu) for c in (8, 16, 32, 64, ncpu)}):
            torch.set_num_threads(c)
            one()
            t = one()
            if t < best[1]:
                best = (c, t)
        threads = best[0]
    torch.set_num_threads(threads)
    for _ in range(warmup):
        one()
    times = [one() for _ in range(steps)]
    total = sum(times)
    return 2 * bs * len(times) / total, total / len(times), threads


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    bs = args.ref_bs
    rate, sec, cores = cpu_oracle_rate(cfg["losses"], bs, args.steps, args.warmup, args.ref_threads)
    sample = "bs=%d pairs (%d images) per step of the same train step, torch CPU fp32, %d threads" % (bs, 2 * bs, cores)
    line = {"impl": "reference", "metric": "images/sec (conv-AE/VAE train step)", "value": rate, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "sample_pairs_per_step": bs},
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def build_line(args, cfg, bs, world, ms, ms_e2e, prof, top, launches, clocks, loss_last, h2d_bytes, d2h_bytes, cpu):
    """The bench JSON line from the measured quantities (pure: unit-tested on CPU).  ms / ms_e2e: device time of the K timed
    steps (max over ranks); prof: {call site: (instances, total ms)} of the timed region; top: the dominant call site;
    launches: libsrlz kernels launched per step (srlz_launch_count difference over the timed region / steps)."""
    losses = cfg["losses"]
    images_per_step = 2 * bs * world
    value = images_per_step / (ms / args.steps) * 1e3
    e2e = images_per_step / (ms_e2e / args.steps) * 1e3
    pk = peaks()
    # ---- roofline of the dominant call site (measured live with CUDA events in the timed region) ----
    traffic_tab = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tpath):
        traffic_tab = json.load(open(tpath)).get("sites", {})

    sr = lambda site: site_roofline(site, prof[site][0], prof[site][1], bs, args.steps, pk, traffic_tab)
    roof = sr(top) or {"kernel": top, "bound": "hbm", "achieved": None, "peak": pk["hbm"], "unit": "GB/s", "frac": None, "traffic": None}
    roof["peak_source"] = pk["source"] + (", bf16 dense sustained" if roof["bound"] == "tensor" else "")
    roof["note"] = ("conv sites: tensor roof = algorithmic fp32-equivalent FLOPs (2*MACs; the tcgen05 kernels issue 1-3 bf16 MMAs per "
                    "product for the hi/lo split), hbm roof = minimum fp32 bytes; `bound` is the roof the site sits closer to, "
                    "`other_roof` the other one")
    ranked = sorted(prof.items(), key=lambda kv: -kv[1][1])
    roof["sites"] = [r for r in (sr(k) for k, _ in ranked[:12]) if r is not None]
    step_ms = ms / args.steps
    roof["time_share_of_step"] = {k: round(v[1] / args.steps / step_ms, 4) for k, v in ranked[:8]}
    if args.prof_out:
        with open(args.prof_out, "w") as f:
            for k, v in ranked:
                f.write("%-14s calls/step %5.1f  ms/step %7.3f  share %5.1f%%\n" % (k, v[0] / args.steps, v[1] / args.steps, 100 * v[1] / args.steps / step_ms))
    whole = {"conv_tflops": value * CONV_TRAIN_FLOP_PER_IMAGE / world / 1e12, "frac_of_bf16_peak": value * CONV_TRAIN_FLOP_PER_IMAGE / world / 1e12 / pk["bf16_sustained"],
             "alg_gbs_fp32": value * ALG_BYTES_PER_IMAGE_FP32 / world / 1e9, "frac_of_hbm_peak": value * ALG_BYTES_PER_IMAGE_FP32 / world / 1e9 / pk["hbm"]}
    return {"metric": "images/sec (conv-AE/VAE train step)", "value": value, "unit": "images/s", "pairs_per_s": value / 2, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "losses": losses, "pairs_per_gpu": bs, "global_pairs": bs * world, "state_dim": S,
                       "parallelism": "dp%d" % world, "l2": "inputs (2 x %.0f MB per rank) exceed the 126 MB L2" % (bs * 3 * IMG * IMG * 4 / 1e6),
                       "loss_last": loss_last},
            "e2e": {"value": e2e, "unit": "images/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "api": "srl_zoo_b200.TrainStep.step_host (pinned host buffers; H2D on a copy stream, next minibatch prefetched one step ahead)"},
            "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches), "clocks": clocks, "roofline": roof,
            "whole_step": whole, "cpu_baseline": cpu}


def run_b200(args, cfg):
    import torch
    import torch.distributed as dist
    import srl_zoo_b200
    from srl_zoo_b200 import _lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl b200) needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bs = args.bs or cfg["bs"]
    losses = cfg["losses"]
    kind = kind_of(losses)
    torch.manual_seed(1)  # train.py:27 ; identical replicas on every rank
    mod = srl_zoo_b200.B200SRLModules(S, A, True, "custom_cnn", losses).to(dev)
    eng = srl_zoo_b200.TrainStep(mod, bs, lr=0.005, beta=1.0, world_size=world)
    obs_h, nobs_h, act_h = synthetic(bs, 1234 + rank, pin=True)
    obs, nobs, act = obs_h.to(dev), nobs_h.to(dev), act_h.to(dev)
    kw = {}
    if kind == "dae":
        import numpy as np
        from srl_zoo_b200.occlusion import sample_rects
        rng = np.random.RandomState(1 + rank)
        kw = dict(rects=torch.from_numpy(sample_rects(bs, rng=rng)).to(dev), next_rects=torch.from_numpy(sample_rects(bs, rng=rng)).to(dev))
    use_act = "forward" in losses or "inverse" in losses

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    dev_step = lambda: eng.step(obs, nobs, act if use_act else None, **kw)
    # two pinned host minibatches alternate, as a loader queue would hand them over: every step copies ITS inputs host ->
    # device inside the timed region (issued one step ahead through `prefetch`, overlapping the previous step's kernels)
    obs_h2, nobs_h2, act_h2 = synthetic(bs, 4321 + rank, pin=True)
    host_batches = [(obs_h, nobs_h, act_h), (obs_h2, nobs_h2, act_h2)]
    host_i = [0]

    def host_step():
        cur, nxt = host_batches[host_i[0] & 1], host_batches[(host_i[0] + 1) & 1]
        host_i[0] += 1
        return eng.step_host(cur[0], cur[1], cur[2] if use_act else None, prefetch=(nxt[0], nxt[1], nxt[2] if use_act else None), **kw)
    for _ in range(args.warmup):
        dev_step()
    # which call site dominates? (one profiled step, outside the timed region)
    _lib.prof_enable(True)
    dev_step()
    prof1 = _lib.prof_report()
    top = max(prof1, key=lambda k: prof1[k][1])
    l0 = _lib.lib.srlz_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.prof_enable(True)  # per-call-site CUDA events ride along in the timed region (2 event records per call site)
    ms = timed(dev_step, args.steps)
    prof = _lib.prof_report()
    _lib.prof_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    launches = (_lib.lib.srlz_launch_count() - l0) // max(args.steps, 1)
    last = eng.step(obs, nobs, act if use_act else None, training=False, **kw).tolist()
    for _ in range(2):
        host_step()
    ms_e2e = timed(host_step, args.steps)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_rate, cpu_sec, cores = cpu_oracle_rate(losses, args.ref_bs, 2, 1, args.ref_threads)
        cpu = {"value": cpu_rate, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "2 timed steps of bs=%d pairs (%d images) of the same train step on the host CPU, torch fp32, %d threads" % (args.ref_bs, 2 * args.ref_bs, cores)}
    line = build_line(args, cfg, bs, world, ms, ms_e2e, prof, top, launches, clocks, dict(zip([n for n in eng.loss_names()], last[:4])),
                      eng.h2d_bytes_per_step(use_act) * world, eng.d2h_bytes_per_step() * world, cpu)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="ae", choices=sorted(CONFIGS))
    ap.add_argument("--bs", type=int, default=0, help="pairs per GPU (default: the config's)")
    ap.add_argument("--ref-bs", type=int, default=8, help="pairs per step of the CPU sample (reference arm / cpu_baseline)")
    ap.add_argument("--ref-threads", type=int, default=0, help="CPU threads of the reference arm (0: probe and keep the fastest)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prof-out", default="", help="write the per-call-site table (ms per step) of the timed region to this file")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_b200(args, cfg)


if __name__ == "__main__":
    main()
