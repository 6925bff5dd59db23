"""bench.py's per-site roofline model on CPU: field contract of the JSON line's `roofline` object, the choice of the binding
roof, and the algorithmic byte / MAC tables against the layer geometry."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

PK = dict(hbm=6547.2, bf16=1682.4, bf16_sustained=1396.2, source="test")


def test_conv_site_reports_the_closer_roof_and_keeps_the_other():
    # dec12.fwd at 0.227 ms per launch of 256 images (profiles/r1_callsite_ms_per_step.txt): HBM-bound
    r = bench.site_roofline("dec12.fwd", 20, 20 * 0.227, 256, 10, PK, {"dec12.fwd": {"dram_bytes_per_launch": 1108500000}})
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] == PK["hbm"]
    alg = (111 * 111 * 64 + 2 * 3 * 224 * 224) * 4 * 256
    assert abs(r["achieved"] - alg / 0.227e-3 / 1e9) < 1e-6 * r["achieved"]
    assert abs(r["frac"] - r["achieved"] / PK["hbm"]) < 1e-12
    assert r["traffic"] == 1108500000 and r["traffic"] / alg < 1.05          # measured DRAM traffic ~ algorithmic bytes
    o = r["other_roof"]
    assert o["bound"] == "tensor" and abs(o["achieved"] - 2 * bench.FWD_MACS["dec12"] * 256 / 0.227e-3 / 1e12) < 1e-9
    # a compute-heavy hypothetical: the same site 100x faster than HBM allows flips to the tensor roof only if that is closer
    r2 = bench.site_roofline("enc4.fwd", 1, 0.001, 256, 1, PK, {})
    assert r2["bound"] in ("hbm", "tensor") and r2["frac"] >= r2["other_roof"]["frac"]


def test_elementwise_site_uses_bytes_per_launch():
    # bn.bwd: 8 launches per step (4 decoder stages x 2 model calls), 3 tensors of each stage move
    r = bench.site_roofline("bn.bwd", 80, 10 * 1.104, 256, 10, PK, {})
    per_step = 3 * (13 * 13 + 27 * 27 + 55 * 55 + 111 * 111) * 64 * 4 * 2 * 256
    assert abs(r["achieved"] - per_step / 1.104e-3 / 1e9) < 1e-6 * r["achieved"]
    assert r["bound"] == "hbm" and r["other_roof"] is None
    assert bench.site_roofline("fc.bwd", 1, 1.0, 256, 1, PK, {}) is None       # no model for the small dense sites


def test_tables_match_the_layer_geometry():
    assert bench.FWD_MACS["enc0"] == 112 * 112 * 64 * 3 * 49
    assert bench.FWD_MACS["enc4"] == 56 * 56 * 64 * 64 * 9
    assert bench.FWD_MACS["dec12"] == 111 * 111 * 64 * 3 * 16
    assert bench.SITE_BYTES["enc0.fwd"] == (3 * 224 * 224 + 112 * 112 * 64) * 4
    assert bench.SITE_BYTES["dec9.dgrad"] == (111 * 111 + 2 * 55 * 55) * 64 * 4
    for site in bench.SITE_BYTES:
        assert site.split(".")[0] in bench.FWD_MACS


def test_committed_traffic_table_is_close_to_algorithmic_bytes():
    tab = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["sites"]
    for site, rec in tab.items():
        if site in bench.SITE_BYTES:
            ratio = rec["dram_bytes_per_launch"] / (bench.SITE_BYTES[site] * 256)
            assert 0.85 < ratio < 1.15, (site, ratio)    # no wasted HBM re-reads on any conv site


def _committed_profile():
    prof = {}
    for ln in open(os.path.join(ROOT, "profiles", "r1_callsite_ms_per_step.txt")):
        f = ln.split()
        prof[f[0]] = (int(round(float(f[2]) * 10)), float(f[4]) * 10)      # 10 timed steps
    return prof


def test_build_line_contract(tmp_path):
    """The JSON line bench.py prints, assembled from the committed per-call-site profile: every key of the driver's contract is
    present, the values are consistent with each other, and the line survives a JSON round trip."""
    import argparse
    prof = _committed_profile()
    top = max(prof, key=lambda k: prof[k][1])
    args = argparse.Namespace(steps=10, warmup=3, prof_out=str(tmp_path / "prof.txt"))
    cfg = bench.CONFIGS["ae"]
    clocks = {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 3}
    cpu = {"value": 232.0, "unit": "images/s", "cores": 16, "kind": "port", "sample": "test"}
    configs = {"vae": {"workload": bench.CONFIGS["vae"]["name"], "pairs_per_gpu": 128, "value": 31340.0, "ms_per_step": 8.17, "e2e": 33281.0,
                       "e2e_ms_per_step": 7.69, "unit": "images/s", "gpu_launches_per_step": 199}}
    line = bench.build_line(args, cfg, 256, 1, 135.5, 131.9, prof, 10, top, 181, clocks, {"reconstruction_loss": 6.7}, 77070336, 32, cpu,
                            configs, {"value": 30000.0, "unit": "images/s"})
    line = json.loads(json.dumps(line))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in line, k
    assert abs(line["value"] - 512 / 13.55 * 1e3) < 1.0 and abs(line["e2e"]["value"] - 512 / 13.19 * 1e3) < 1.0
    assert line["config"]["workload"].startswith("conv autoencoder") and "model" not in line["config"]
    assert line["vs_baseline"] is None and line["scaling"] == "weak" and line["dtype"] == "f32/bf16x3"
    assert line["configs"]["vae"]["pairs_per_gpu"] == 128 and line["dropin"]["value"] == 30000.0
    assert line["gpu_launches"] == 181 * 10 and line["gpu_launches_per_step"] == 181 and line["pairs_per_s"] * 2 == line["value"]
    r = line["roofline"]
    for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic", "sites", "time_share_of_step"):
        assert k in r, k
    assert r["kernel"] == top and r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["traffic"] is not None and len(r["sites"]) >= 8
    assert os.path.exists(args.prof_out)
    # two ranks: whole-job throughput doubles for the same step time
    line2 = bench.build_line(args, cfg, 256, 2, 135.5, 131.9, prof, 10, top, 181, clocks, {}, 2 * 77070336, 64, None)
    assert "configs" not in line2 and "dropin" not in line2
    assert abs(line2["value"] - 2 * line["value"]) < 1e-6 * line["value"] and line2["config"]["parallelism"] == "dp2"
