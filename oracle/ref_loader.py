"""
ORACLE -- TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Nothing on the product path may import this file.

Imports the UNMODIFIED reference (araffin/srl-zoo) from `oracle/_ref/` (vendored copy that travels to the GPU box, see
oracle/vendor_reference.py) or from the read-only checkout `/root/reference` (build container), with the two import stubs
the reference needs under this image (`termcolor` is not installed; `plotting/__init__.py:4` shells out
to `xset` and needs matplotlib / seaborn).  On top of the imported modules:

  * RefStep      -- the minibatch body of SRL4robotics.learn (models/learner.py:373-497) driven with the reference's own
                    SRLModules + LossManager + loss functions + th.optim.Adam on given tensors (what `bench.py --impl reference`
                    times on the host cores, and what the oracle restatement is validated against).
  * run_train_py -- the reference's literal `train.py` (train.py:23-212) through runpy on a dataset folder.
"""
import contextlib
import os
import runpy
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.path.join(HERE, "_ref"), os.environ.get("SRL_REFERENCE", "/root/reference"))
# models/learner.py:204-207
DEFAULT_WEIGHTS = {"forward": 1.0, "inverse": 2.0, "autoencoder": 1.0, "vae": 0.5e-6, "dae": 1.0}


def find_root():
    for c in CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "models", "learner.py")):
            return c
    return None


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    if "termcolor" not in sys.modules:
        try:
            import termcolor  # noqa: F401
        except ImportError:
            _stub("termcolor", colored=lambda s, *a, **k: s)                      # utils.py:10
    if "plotting" not in sys.modules:
        p = _stub("plotting")
        p.__path__ = []
        _stub("plotting.representation_plot", plotRepresentation=lambda *a, **k: None, plotImage=lambda *a, **k: None, plt=None,
              INTERACTIVE_PLOT=False)
        _stub("plotting.losses_plot", plotLosses=lambda *a, **k: None)


def load():
    """-> namespace(root, learner, modules, losses) of the reference's own modules, or None when no copy is present"""
    root = find_root()
    if root is None:
        return None
    sys.dont_write_bytecode = True
    install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    import models.learner as learner
    import models.modules as modules
    import losses.losses as losses
    return types.SimpleNamespace(root=root, learner=learner, modules=modules, losses=losses)


class RefStep:
    """One training minibatch exactly as models/learner.py:373-497 does it, with the reference's own objects.
    kind: "ae" | "dae" | "vae"; the module / loss names are looked up on `ns` at call time, so the same driver runs the stock
    reference (ns = the reference's modules) and the reference under srl_zoo_b200.install() (ns = models.learner)."""

    def __init__(self, ns, kind, use_forward=False, use_inverse=False, device="cpu", state_dim=200, action_dim=6, lr=0.005,
                 beta=1.0, seed=1, inverse_model_type="linear"):
        import torch
        self.th, self.ns, self.kind, self.beta = torch, ns, kind, beta
        self.use_forward, self.use_inverse = use_forward, use_inverse
        losses = [{"ae": "autoencoder", "dae": "dae", "vae": "vae"}[kind]] + (["forward"] if use_forward else []) + \
            (["inverse"] if use_inverse else [])
        self.losses = losses
        torch.manual_seed(seed)                                                       # learner.py:59
        self.device = torch.device(device)
        self.model = ns.SRLModules(state_dim=state_dim, action_dim=action_dim, model_type="custom_cnn", cuda=self.device.type == "cuda",
                                   losses=losses, inverse_model_type=inverse_model_type).to(self.device)   # learner.py:179-192
        params = [p for p in self.model.parameters() if p.requires_grad]              # learner.py:194-199
        self.optimizer = torch.optim.Adam(params, lr=lr)
        self.loss_history = {}
        self.w = dict(DEFAULT_WEIGHTS)

    def step(self, obs, next_obs, actions=None, noisy_obs=None, next_noisy_obs=None, training=True):
        ns, th, model = self.ns, self.th, self.model
        from collections import defaultdict
        history = defaultdict(list)
        lm = ns.LossManager(model, history)
        model.train(training)                                                          # learner.py:362-366
        self.optimizer.zero_grad()                                                     # learner.py:373
        lm.resetLosses()
        if self.kind == "vae":                                                         # learner.py:399-402
            (decoded_obs, mu, logvar), (next_decoded_obs, next_mu, next_logvar) = model(obs), model(next_obs)
            states, next_states = model.getStates(obs), model.getStates(next_obs)
        elif self.kind == "dae":                                                       # learner.py:395-397
            (states, decoded_obs), (next_states, next_decoded_obs) = model(noisy_obs), model(next_noisy_obs)
        else:                                                                          # learner.py:392-393
            (states, decoded_obs), (next_states, next_decoded_obs) = model(obs), model(next_obs)
        if self.use_forward:                                                           # learner.py:432-436
            next_states_pred = model.forwardModel(states, actions)
            ns.forwardModelLoss(next_states_pred, next_states, weight=self.w["forward"], loss_manager=lm)
        if self.use_inverse:                                                           # learner.py:438-441
            actions_pred = model.inverseModel(states, next_states)
            ns.inverseModelLoss(actions_pred, actions, weight=self.w["inverse"], loss_manager=lm)
        if self.kind in ("ae", "dae"):                                                 # learner.py:452-455
            ns.autoEncoderLoss(obs, decoded_obs, next_obs, next_decoded_obs, weight=self.w["dae" if self.kind == "dae" else "autoencoder"],
                               loss_manager=lm)
        else:                                                                          # learner.py:457-468
            ns.kullbackLeiblerLoss(mu, next_mu, logvar, next_logvar, loss_manager=lm, beta=self.beta)
            ns.generationLoss(decoded_obs, next_decoded_obs, obs, next_obs, weight=self.w["vae"], loss_manager=lm)
        lm.updateLossHistory()                                                         # learner.py:484
        loss = lm.computeTotalLoss()
        loss.backward()                                                                # learner.py:489
        if training:
            self.optimizer.step()                                                      # learner.py:495
        total = loss.item()
        return {"total": total, "losses": {n: float(v.detach()) for n, v in zip(lm.names, lm.losses)}, "states": states.detach(),
                "decoded": decoded_obs.detach(), "history": dict(history)}


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def run_train_py(workdir, argv, before_main=None):
    """Runs the reference's unchanged train.py (train.py:23-212) with cwd = workdir (which holds `data/<dataset>/`) and the
    given argument list; `correlationCall` (pipeline.py: spawns `python -m evaluation.knn_images`, out of scope) is a no-op.
    before_main(learner_module): hook run after the reference is imported and before train.py executes (install() goes here).
    Returns the dict of train.py's globals (args, loss_history, learned_states, ...)."""
    ref = load()
    if ref is None:
        raise RuntimeError("no reference copy (oracle/_ref or /root/reference)")
    import pipeline
    saved = pipeline.correlationCall
    pipeline.correlationCall = lambda *a, **k: None
    if before_main is not None:
        before_main(ref.learner)
    old_argv = sys.argv
    sys.argv = ["train.py"] + list(argv)
    try:
        with _cwd(workdir):
            if "--log-folder" in argv:   # train.py only creates the dated default folder (train.py:149-157)
                os.makedirs(argv[list(argv).index("--log-folder") + 1], exist_ok=True)
            return runpy.run_path(os.path.join(ref.root, "train.py"), run_name="__main__")
    finally:
        sys.argv = old_argv
        pipeline.correlationCall = saved
