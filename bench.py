#!/usr/bin/env python
"""
bench.py -- images/sec of the conv-AE / VAE train step (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5            # this repo's CUDA path (hand-written sm_100a kernels)
    python bench.py --impl reference --steps 5 --warmup 1     # the reference's own modules (oracle/_ref) on the host CPU
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU, weak scaling

One "step" = one training minibatch of SRL4robotics.learn (models/learner.py:373-497): zero-grad, forward(obs),
forward(next_obs), losses, backward, (one all-reduce of the flat gradient buffer), Adam, per-loss scalars.
An "image" is one 224x224x3 observation through encoder+decoder forward and backward; a minibatch of bs pairs is
2*bs images.  value = 2 * bs_global / step_time, inputs resident in HBM; e2e = the same through TrainStep.step_host
(pinned host uint8 frames -> H2D every step, normalised on the device, loss scalars D2H every step).  The line also carries
`configs` (the other BASELINE configs at this N), `roofline` (separate profiled pass), `cpu_baseline` and `dropin`.
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
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG, S, A = 224, 200, 6
# exact per-image MAC counts of the instantiated reference modules, x2 = FLOP
FWD_MACS = {"enc0": 118013952, "enc4": 115605504, "enc8": 7225344, "dec0": 1327104, "dec3": 6230016, "dec6": 26873856,
            "dec9": 111513600, "dec12": 37850112}
CONV_TRAIN_FLOP_PER_IMAGE = 2311809024   # fwd + wgrad (all) + dgrad (all but enc0), conv/convT layers
ALG_BYTES_PER_IMAGE_FP32 = 44.7e6        # perfect-fusion lower bound, fp32 activations
# elements moved per image by the elementwise / reduction call sites (read + write), fp32
EW_ELEMS = {"bn_relu_pool.fwd": 802816 + 200704 + 12544 + 1.25 * (200704 + 46656 + 2304),   # y in; pooled out + 1-byte argmax
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


def synthetic_u8(bs, seed):
    """uint8 U{0..255} RGB frames in the loader's native (B, H, W, 3) order, pinned; actions U{0..5}.
    Normalisation (/255, ImageNet mean / std, preprocessing/utils.py:20-32) and the (C, W, H) transpose of
    preprocessing/data_loader.py:255 happen on the device (srlz_preprocess_u8)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    obs = torch.randint(0, 256, (bs, IMG, IMG, 3), generator=g, dtype=torch.uint8).pin_memory()
    nobs = torch.randint(0, 256, (bs, IMG, IMG, 3), generator=g, dtype=torch.uint8).pin_memory()
    actions = torch.randint(0, A, (bs, 1), generator=g, dtype=torch.int64).pin_memory()
    return obs, nobs, actions


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference_rate(losses, bs, steps, warmup, threads=0):
    """images/s of the reference's CPU implementation of the path on the host cores: the reference's OWN modules (oracle/_ref,
    the unmodified files vendored by build()) driven through the exact call sequence of models/learner.py:373-497
    (oracle/ref_loader.RefStep: SRLModules + LossManager + loss functions + th.optim.Adam, device = cpu, i.e. `--no-cuda`
    semantics); when no reference copy is present the oracle port (same torch CPU ops) stands in and `kind` says so.
    threads=0: one probe step at {16,32,64,all cores} picks the thread count (torch's CPU convs stop scaling well before
    128 threads at bs=32), which then stays fixed; the value is the MEDIAN step of the timed ones.
    -> (images/s, seconds per step, threads, kind)"""
    import numpy as np
    import torch
    from oracle import ref_loader, srl_oracle as O
    ncpu = os.cpu_count() or 1
    kind = kind_of(losses)
    obs, nobs, actions = O.synthetic_batch(bs, seed=1234)
    rng = np.random.RandomState(1)
    rects = (O.sample_rects(bs, rng=rng), O.sample_rects(bs, rng=rng))
    ref = ref_loader.load()
    if ref is not None:
        ns = types.SimpleNamespace(SRLModules=ref.modules.SRLModules, LossManager=ref.losses.LossManager,
                                   **{n: getattr(ref.losses, n) for n in ("autoEncoderLoss", "generationLoss", "kullbackLeiblerLoss",
                                                                          "forwardModelLoss", "inverseModelLoss")})
        drv = ref_loader.RefStep(ns, kind, "forward" in losses, "inverse" in losses, device="cpu", state_dim=S, action_dim=A)
        noisy = (O.apply_occlusion(obs, rects[0]), O.apply_occlusion(nobs, rects[1])) if kind == "dae" else (None, None)

        def one():
            t0 = time.perf_counter()
            drv.step(obs, nobs, actions, noisy[0], noisy[1])
            return time.perf_counter() - t0
        which = "reference"
    else:
        sd = O.build_state("vae" if kind == "vae" else "ae", S, A, seed=1)
        P, B = O.split_state(sd)
        opt = O.Adam(P, lr=0.005)

        def one():
            t0 = time.perf_counter()
            O.train_step(kind, P, B, obs, nobs, actions, None, None, rects[0], rects[1], use_forward="forward" in losses,
                         use_inverse="inverse" in losses, optimizer=opt)
            return time.perf_counter() - t0
        which = "port"
    if threads <= 0:
        best = (None, 1e30)
        for c in sorted({min(c, ncpu) for c in (16, 32, 64, ncpu)}):
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
    sec = statistics.median(times)
    return 2 * bs / sec, sec, threads, which


def train_py_rate(bs=32, n_frames=161):
    """The literal `python train.py --no-cuda` leg (train.py:23-212, timer at models/learner.py:528): the reference's unchanged
    train.py on a synthetic JPEG dataset folder, one epoch, CPU; images/s = 2*bs*minibatches / seconds of the epoch loop
    (includes the reference's loader: cv2 JPEG decode on its worker threads).  -> dict or None when no reference copy."""
    import tempfile
    from oracle import ref_loader, synth_dataset
    if ref_loader.find_root() is None:
        return None
    work = tempfile.mkdtemp(prefix="srlz_trainpy_")
    synth_dataset.make_dataset(work, n_frames=n_frames)
    stamps = {}
    ref = ref_loader.load()
    real_learn = ref.learner.SRL4robotics.learn

    def timed_learn(self, *a, **k):
        stamps["t0"] = time.perf_counter()
        out = real_learn(self, *a, **k)
        stamps["t1"] = time.perf_counter()
        return out
    ref.learner.SRL4robotics.learn = timed_learn   # a stopwatch around learn(); the function itself is untouched
    try:
        g = ref_loader.run_train_py(work, ["--no-cuda", "--no-display-plots", "--epochs", "1", "--losses", "autoencoder", "--model-type",
                                           "custom_cnn", "--state-dim", str(S), "-bs", str(bs), "--data-folder", "synth", "--log-folder",
                                           os.path.join(work, "logs", "run")])
    finally:
        ref.learner.SRL4robotics.learn = real_learn
    n_mb = (n_frames - 1) // bs
    sec = stamps["t1"] - stamps["t0"]
    return {"value": 2 * bs * n_mb / sec, "unit": "images/s", "seconds_learn": sec, "minibatches": n_mb, "pairs_per_minibatch": bs,
            "what": "reference train.py --no-cuda --epochs 1 --losses autoencoder --model-type custom_cnn --state-dim %d -bs %d on a "
                    "synthetic %d-frame JPEG dataset: learn() wall time incl. the loader and the final state prediction pass" % (S, bs, n_frames),
            "loss_history_keys": sorted(g["loss_history"].keys())}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    bs = args.ref_bs
    rate, sec, cores, which = cpu_reference_rate(cfg["losses"], bs, args.steps, args.warmup, args.ref_threads)
    sample = "median of %d steps of bs=%d pairs (%d images) of the same train step, %s on the host CPU (%s, %d logical cores), torch fp32, %d threads" % (
        args.steps, bs, 2 * bs, "the reference's own modules (oracle/_ref) through the learner.py:373-497 call sequence" if which == "reference"
        else "oracle port (no reference copy present)", cpu_model(), os.cpu_count() or 1, cores)
    line = {"impl": "reference", "metric": "images/sec (conv-AE/VAE train step)", "value": rate, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "sample_pairs_per_step": bs},
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": which, "sample": sample, "cpu_model": cpu_model()},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.no_train_py:
        # its own process: the reference's loader forks a worker (preprocessing/data_loader.py:120-126), which must not
        # inherit this process's warmed-up OpenMP pool
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "train_py"], stdout=subprocess.PIPE,
                               stderr=subprocess.DEVNULL, text=True, timeout=900)
            line["train_py"] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:  # the literal leg is a reported extra: never lose the line over it
            line["train_py"] = {"error": "%s: %s" % (type(e).__name__, e)}
    print(json.dumps(line), flush=True)


def build_line(args, cfg, bs, world, ms, ms_e2e, prof, prof_steps, top, launches, clocks, loss_last, h2d_bytes, d2h_bytes, cpu,
               configs=None, dropin=None):
    """The bench JSON line from the measured quantities (pure: unit-tested on CPU).  ms / ms_e2e: device time of the K timed
    steps (max over ranks); prof: {call site: (instances, total ms)} of a SEPARATE profiled pass of prof_steps steps (the
    per-call-site CUDA events do not ride in the headline region); top: the dominant call site; launches: libsrlz kernels
    launched per step (srlz_launch_count difference over the timed region / steps)."""
    losses = cfg["losses"]
    images_per_step = 2 * bs * world
    value = images_per_step / (ms / args.steps) * 1e3
    e2e = images_per_step / (ms_e2e / args.steps) * 1e3
    pk = peaks()
    # ---- roofline of the dominant call site (measured live with CUDA events, profiled pass) ----
    traffic_tab = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tpath):
        traffic_tab = json.load(open(tpath)).get("sites", {})

    sr = lambda site: site_roofline(site, prof[site][0], prof[site][1], bs, prof_steps, pk, traffic_tab)
    roof = sr(top) or {"kernel": top, "bound": "hbm", "achieved": None, "peak": pk["hbm"], "unit": "GB/s", "frac": None, "traffic": None}
    roof["peak_source"] = pk["source"] + (", bf16 dense sustained" if roof["bound"] == "tensor" else "")
    roof["note"] = ("conv sites: tensor roof = algorithmic fp32-equivalent FLOPs (2*MACs; the tcgen05 kernels issue 1-3 bf16 MMAs per "
                    "product for the hi/lo split), hbm roof = minimum fp32 bytes; `bound` is the roof the site sits closer to, "
                    "`other_roof` the other one; timed in a separate profiled pass of %d steps" % prof_steps)
    ranked = sorted(prof.items(), key=lambda kv: -kv[1][1])
    roof["sites"] = [r for r in (sr(k) for k, _ in ranked[:12]) if r is not None]
    step_ms = ms / args.steps
    prof_step_ms = sum(v[1] for v in prof.values()) / prof_steps
    roof["time_share_of_step"] = {k: round(v[1] / prof_steps / prof_step_ms, 4) for k, v in ranked[:8]}
    if args.prof_out:
        with open(args.prof_out, "w") as f:
            for k, v in ranked:
                f.write("%-14s calls/step %5.1f  ms/step %7.3f  share %5.1f%%\n" % (k, v[0] / prof_steps, v[1] / prof_steps, 100 * v[1] / prof_steps / prof_step_ms))
    whole = {"conv_tflops": value * CONV_TRAIN_FLOP_PER_IMAGE / world / 1e12, "frac_of_bf16_peak": value * CONV_TRAIN_FLOP_PER_IMAGE / world / 1e12 / pk["bf16_sustained"],
             "alg_gbs_fp32": value * ALG_BYTES_PER_IMAGE_FP32 / world / 1e9, "frac_of_hbm_peak": value * ALG_BYTES_PER_IMAGE_FP32 / world / 1e9 / pk["hbm"]}
    line = {"metric": "images/sec (conv-AE/VAE train step)", "value": value, "unit": "images/s", "pairs_per_s": value / 2, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32/bf16x3", "data": "synthetic",
            "config": {"workload": cfg["name"], "losses": losses, "pairs_per_gpu": bs, "global_pairs": bs * world, "state_dim": S,
                       "parallelism": "dp%d" % world, "l2": "inputs (2 x %.0f MB per rank) exceed the 126 MB L2" % (bs * 3 * IMG * IMG * 4 / 1e6),
                       "dtype_note": "fp32 storage and accumulation; conv products as bf16 hi/lo split tensor-core MMAs (bf16x3, ~2^-17 per operand)",
                       "loss_last": loss_last},
            "e2e": {"value": e2e, "unit": "images/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "api": "srl_zoo_b200.TrainStep.step_host (pinned host uint8 HWC frames, the loader's native order; H2D on a copy stream, next minibatch prefetched one step ahead; normalisation + transpose on the device)"},
            "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches), "clocks": clocks, "roofline": roof,
            "whole_step": whole, "cpu_baseline": cpu}
    if configs:
        line["configs"] = configs
    if dropin:
        line["dropin"] = dropin
    return line


def measure(args, name, world, rank, dev, profile):
    """One BASELINE config on this rank's GPU: K timed steps with device-resident inputs, K timed steps through step_host
    with pinned host uint8 frames, and (profile=True) a separate per-call-site pass.  -> dict of raw measurements."""
    import torch
    import torch.distributed as dist
    import srl_zoo_b200
    from srl_zoo_b200 import _lib
    cfg = CONFIGS[name]
    bs = (args.bs if name == args.config else 0) or cfg["bs"]
    losses = cfg["losses"]
    kind = kind_of(losses)
    torch.manual_seed(1)  # train.py:27 ; identical replicas on every rank
    mod = srl_zoo_b200.B200SRLModules(S, A, True, "custom_cnn", losses).to(dev)
    eng = srl_zoo_b200.TrainStep(mod, bs, lr=0.005, beta=1.0, world_size=world)
    u8 = [synthetic_u8(bs, s + rank) for s in (1234, 4321)]
    obs, nobs = eng.preprocess(u8[0][0].to(dev)), eng.preprocess(u8[0][1].to(dev))
    act = u8[0][2].to(dev)
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
    # two pinned host minibatches of uint8 HWC frames alternate, as a loader queue would hand them over: every step copies ITS
    # inputs host -> device inside the timed region (issued one step ahead through `prefetch`, overlapping the previous step)
    host_i = [0]

    def host_step():
        cur, nxt = u8[host_i[0] & 1], u8[(host_i[0] + 1) & 1]
        host_i[0] += 1
        return eng.step_host(cur[0], cur[1], cur[2] if use_act else None, prefetch=(nxt[0], nxt[1], nxt[2] if use_act else None), **kw)
    for _ in range(args.warmup):
        dev_step()
    l0 = _lib.lib.srlz_launch_count()
    ms = timed(dev_step, args.steps)
    launches = (_lib.lib.srlz_launch_count() - l0) // max(args.steps, 1)
    out = dict(name=name, cfg=cfg, bs=bs, ms=ms, launches=launches, eng=eng)
    if profile:   # per-call-site CUDA events in their own pass: the headline region above carries none
        _lib.prof_enable(True)
        timed(dev_step, args.steps)
        out["prof"] = _lib.prof_report()
        _lib.prof_enable(False)
        out["top"] = max(out["prof"], key=lambda k: out["prof"][k][1])
    out["loss_last"] = dict(zip(eng.loss_names(), eng.step(obs, nobs, act if use_act else None, training=False, **kw).tolist()[:4]))
    for _ in range(3):
        host_step()
    out["ms_e2e"] = timed(host_step, args.steps)
    out["h2d"] = eng.h2d_bytes_per_step(use_act) * world
    out["d2h"] = eng.d2h_bytes_per_step() * world
    return out


def dropin_rate(args, dev, bs=256):
    """images/s of the reference's unchanged minibatch body (models/learner.py:373-497: model(obs), loss functions,
    loss.backward(), th.optim.Adam) with srl_zoo_b200.install() applied to the reference's `models.learner`: the drop-in path
    a user of train.py gets without touching learn().  Device-resident inputs (the reference's loader is out of scope)."""
    import torch
    import srl_zoo_b200
    from oracle import ref_loader
    ref = ref_loader.load()
    if ref is None:
        return None
    srl_zoo_b200.install(ref.learner, ref.modules)
    drv = ref_loader.RefStep(ref.learner, "ae", device=str(dev), state_dim=S, action_dim=A)
    assert isinstance(drv.model, srl_zoo_b200.B200SRLModules)
    g = torch.Generator().manual_seed(5)
    obs = torch.randn(bs, 3, IMG, IMG, generator=g).to(dev)
    nobs = torch.randn(bs, 3, IMG, IMG, generator=g).to(dev)
    for _ in range(3):
        drv.step(obs, nobs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        drv.step(obs, nobs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    return {"value": 2 * bs / ms * 1e3, "unit": "images/s", "ms_per_step": ms, "pairs_per_gpu": bs,
            "api": "reference learner minibatch body on install()ed names: B200SRLModules through torch.autograd + srl_zoo_b200.losses + torch.optim.Adam"}


def run_b200(args, cfg):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl b200) needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()   # runs across every timed region of this process
    head = measure(args, args.config, world, rank, dev, profile=True)
    head.pop("eng")
    torch.cuda.empty_cache()
    configs = {}
    if not args.no_other_configs:
        for name in CONFIGS:   # the other BASELINE configs at this N (X1): same K / W, their own per-GPU batch
            if name == args.config:
                continue
            m = measure(args, name, world, rank, dev, profile=False)
            m.pop("eng")
            torch.cuda.empty_cache()
            per_step = 2 * m["bs"] * world
            configs[name] = {"workload": m["cfg"]["name"], "pairs_per_gpu": m["bs"], "value": per_step / (m["ms"] / args.steps) * 1e3,
                             "ms_per_step": m["ms"] / args.steps, "e2e": per_step / (m["ms_e2e"] / args.steps) * 1e3,
                             "e2e_ms_per_step": m["ms_e2e"] / args.steps, "unit": "images/s", "gpu_launches_per_step": int(m["launches"])}
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu, dropin = None, None
    if world == 1 and not args.no_dropin:
        try:
            dropin = dropin_rate(args, dev)
        except Exception as e:
            dropin = {"error": "%s: %s" % (type(e).__name__, e)}
    if world == 1 and not args.no_cpu_baseline:
        cpu_rate, cpu_sec, cores, which = cpu_reference_rate(cfg["losses"], args.ref_bs, 5, 1, args.ref_threads)
        cpu = {"value": cpu_rate, "unit": "images/s", "cores": cores, "kind": which, "cpu_model": cpu_model(),
               "sample": "median of 5 timed steps of bs=%d pairs (%d images) of the same train step on the host CPU (%s), torch fp32, %d threads" % (
                   args.ref_bs, 2 * args.ref_bs, "the reference's own modules from oracle/_ref, learner.py:373-497 call sequence" if which == "reference" else "oracle port", cores)}
    line = build_line(args, head["cfg"], head["bs"], world, head["ms"], head["ms_e2e"], head["prof"], args.steps, head["top"], head["launches"],
                      clocks, head["loss_last"], head["h2d"], head["d2h"], cpu, configs, dropin)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "train_py"])
    ap.add_argument("--config", default="ae", choices=sorted(CONFIGS))
    ap.add_argument("--bs", type=int, default=0, help="pairs per GPU (default: the config's)")
    ap.add_argument("--ref-bs", type=int, default=32, help="pairs per step of the CPU sample (reference arm / cpu_baseline; SURVEY 8d: 32)")
    ap.add_argument("--ref-threads", type=int, default=0, help="CPU threads of the reference arm (0: one probe, then fixed)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="only the headline config (skip the `configs` dict)")
    ap.add_argument("--no-dropin", action="store_true", help="skip the install()ed learner-body rate")
    ap.add_argument("--no-train-py", action="store_true", help="reference arm: skip the literal train.py --no-cuda leg")
    ap.add_argument("--prof-out", default="", help="write the per-call-site table (ms per step) of the profiled pass to this file")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    cfg = CONFIGS[args.config]
    if args.impl == "train_py":   # internal: the literal train.py leg of the reference arm, in its own process
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):
            out = train_py_rate()
        print(json.dumps(out), flush=True)
    elif args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_b200(args, cfg)


if __name__ == "__main__":
    main()
