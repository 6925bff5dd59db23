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


def test_product_never_imports_the_oracle_and_bench_only_in_its_cpu_arm():
    """oracle/ is test infrastructure: no product module may import it (a product path through the oracle would void every parity
    claim), and bench.py may touch it only inside the functions of the CPU arm (cpu_baseline / --impl reference / train_py)."""
    import ast
    import glob
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def oracle_imports(path):
        out = []
        tree = ast.parse(open(path).read())
        for fn in ast.walk(tree):
            scope = fn.name if isinstance(fn, (ast.FunctionDef, ast.AsyncFunctionDef)) else None
            for n in ast.iter_child_nodes(fn) if scope is None else ast.walk(fn):
                if isinstance(n, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in n.names):
                    out.append((scope, n.lineno))
                if isinstance(n, ast.ImportFrom) and (n.module or "").split(".")[0] == "oracle":
                    out.append((scope, n.lineno))
        return out

    for f in glob.glob(os.path.join(root, "srl_zoo_b200", "*.py")):
        assert oracle_imports(f) == [], f
    for f in glob.glob(os.path.join(root, "srl_zoo_b200", "csrc", "*")):
        if f.endswith((".cu", ".cuh", ".h")):
            assert "oracle" not in open(f).read(), f
    scopes = {s for s, _ in oracle_imports(os.path.join(root, "bench.py"))}
    # cpu_reference_rate / train_py_rate: the CPU arm.  dropin_rate: loads the REFERENCE's own learner body (oracle/_ref through
    # oracle/ref_loader) as the CALLER of the installed B200 modules -- the reference-facing API being timed, not the checker.
    assert scopes == {"cpu_reference_rate", "train_py_rate", "dropin_rate"}, scopes
    src = open(os.path.join(root, "bench.py")).read()
    body = src[src.index("def dropin_rate"):src.index("def run_b200")]
    assert "srl_oracle" not in body and "ref_loader" in body      # the restatement itself never runs in the b200 arm


def test_batch_limit_is_checked_before_any_cuda_call():
    """SRLZ_MAX_BATCH (include/srlz.h): the whole-model entry points reject larger calls with SRLZ_E_ARG up front (no launch, so this
    runs without a GPU); the Python mirror of the constant matches the header; the eval-mode wrapper splits instead (exact: no batch
    statistics in eval mode), training-mode calls surface the error."""
    import ctypes as C
    import os
    import re
    from srl_zoo_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "srlz.h")).read()
    assert int(re.search(r"#define\s+SRLZ_MAX_BATCH\s+(\d+)", hdr).group(1)) == _lib.MAX_BATCH == 2048
    net = _lib.SrlzNet()
    net.state_dim = 200
    fake = C.c_void_p(4096)   # never dereferenced: the size check comes first
    rc = _lib.lib.srlz_forward(C.byref(net), fake, fake, None, None, _lib.MAX_BATCH + 1, 1, fake, None, fake, fake, fake, fake, fake, None)
    assert rc == 1001 and b"SRLZ_MAX_BATCH" in _lib.lib.srlz_last_error()
    rc = _lib.lib.srlz_encode_eval(C.byref(net), fake, fake, None, _lib.MAX_BATCH + 1, fake, fake, None)
    assert rc == 1001 and b"SRLZ_MAX_BATCH" in _lib.lib.srlz_last_error()
    rc = _lib.lib.srlz_forward(C.byref(net), fake, fake, None, None, 0, 1, fake, None, fake, fake, fake, fake, fake, None)
    assert rc == 1001 and b"B <= 0" in _lib.lib.srlz_last_error()
