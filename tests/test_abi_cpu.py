"""CPU-only: the C-ABI library loads and exports every symbol include/srlz.h declares (no compute calls)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "srlz.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(srlz_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from srl_zoo_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(_lib.lib, s), "libsrlz.so does not export %s" % s
    assert _lib.lib.srlz_version() == 1
    assert set(_lib.EXPORTED) == set(syms), set(_lib.EXPORTED) ^ set(syms)


def test_block_sizes_are_consistent():
    from srl_zoo_b200 import _lib
    lib = _lib.lib
    assert lib.srlz_pack_floats(0, 200) > 12 * 9 * 4096 + 2 * 200 * 2304   # twelve bf16 hi/lo conv images + the two permuted FC matrices
    assert lib.srlz_pack_floats(1, 200) == lib.srlz_pack_floats(0, 200) + 200 * 2304
    s1, s2 = lib.srlz_saved_bytes(1, 200, 0), lib.srlz_saved_bytes(2, 200, 0)
    assert s1 > 9_000_000 and 1.9 < s2 / s1 < 2.1
    assert lib.srlz_workspace_bytes(4, 200, 0) > 4 * 112 * 112 * 64 * 4


def test_module_mirrors_reference_state_dict():
    """same key set / shapes / init as the oracle's restatement of models/modules.py:37-49 (bit-equal with seed)."""
    import torch
    from oracle import srl_oracle as O
    import srl_zoo_b200
    for kind, losses in (("ae", ["autoencoder"]), ("vae", ["vae"]), ("ae", ["dae", "forward", "inverse"])):
        torch.manual_seed(1)
        mod = srl_zoo_b200.B200SRLModules(200, 6, True, "custom_cnn", losses)
        sd = O.build_state(kind, 200, 6, seed=1)
        msd = mod.state_dict()
        assert list(msd.keys()) == list(sd.keys())
        for k in sd:
            assert torch.equal(msd[k], sd[k]), k


def test_hot_path_has_no_cpu_fallback():
    import pytest
    import torch
    import srl_zoo_b200
    mod = srl_zoo_b200.B200SRLModules(200, 6, True, "custom_cnn", ["autoencoder"])
    with pytest.raises(RuntimeError):
        mod(torch.zeros(1, 3, 224, 224))  # CPU tensor: must fail loudly, never fall back
    with pytest.raises(ValueError):
        srl_zoo_b200.B200SRLModules(200, 6, True, "mlp", ["autoencoder"])


def test_ctypes_signatures_match_the_header():
    """ABI drift guard: for every entry point whose ctypes argtypes are declared in _lib.py, the argument count equals the C
    prototype's in include/srlz.h (ctypes would otherwise pass garbage silently)."""
    from srl_zoo_b200 import _lib
    src = open(os.path.join(ROOT, "include", "srlz.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(srlz_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    assert len(protos) >= 30
    checked = 0
    for name, args in protos:
        args = " ".join(args.split())
        n = 0 if args in ("", "void") else len(args.split(","))
        fn = getattr(_lib.lib, name)
        if fn.argtypes is not None:
            assert len(fn.argtypes) == n, (name, len(fn.argtypes), n)
            checked += 1
    assert checked >= 20
