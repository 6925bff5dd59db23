"""Drop-in installation behind the reference's unchanged train.py + models/learner.py.

The reference has no plugin registry: `models.learner` imports `SRLModules` and the loss functions by name
(models/learner.py:15-17,24) and constructs the model at models/learner.py:179.  `install()` rebinds exactly those
names, so train.py and learner.py stay byte-identical:

    import models.learner, models.modules           # the reference, on sys.path
    import srl_zoo_b200; srl_zoo_b200.install(models.learner, models.modules)
"""
from . import losses as _losses
from .modules import B200SRLModules, B200SRLModulesSplit

_LOSS_NAMES = ("LossManager", "autoEncoderLoss", "generationLoss", "kullbackLeiblerLoss", "forwardModelLoss",
               "inverseModelLoss", "rewardModelLoss")
_COLD_LOSSES = ("triplet", "priors", "episode-prior", "reward-prior", "random", "supervised")   # other model classes / pair mining


def _hot(seen, state_dim, cuda, model_type, losses, inverse_model_type):
    """configurations the B200 module covers: custom_cnn with an autoencoder-family loss on a CUDA device, optional forward /
    inverse ('linear' or 'mlp') / reward heads; perceptual configs (learner.py:317-321,404-412) differentiate a frozen DAE
    w.r.t. its INPUT, so both the VAE and the denoiser built after it (learner.py:319, losses=["dae"]) stay on the reference class"""
    if losses is not None and "perceptual" in losses:
        seen["perceptual"] = True
    return (not seen["perceptual"] and model_type == "custom_cnn" and losses is not None and cuda and state_dim % 4 == 0
            and any(k in losses for k in ("autoencoder", "dae", "vae")) and not any(k in losses for k in _COLD_LOSSES)
            and inverse_model_type in ("linear", "mlp"))


def _make_dispatch(reference_cls, reference_split_cls=None):
    """Class factories with the constructor signatures of SRLModules (models/modules.py:18-19) and SRLModulesSplit
    (models/modules.py:104-105): hot-path configurations get the B200 modules, everything else falls through to the reference classes."""
    seen = {"perceptual": False}

    def SRLModules(state_dim=2, action_dim=6, cuda=False, model_type="custom_cnn", losses=None, inverse_model_type="linear"):
        if _hot(seen, state_dim, cuda, model_type, losses, inverse_model_type):
            return B200SRLModules(state_dim, action_dim, cuda, model_type, losses, inverse_model_type)
        if reference_cls is None:
            raise ValueError("configuration outside the B200 hot path and no reference class to fall back to")
        return reference_cls(state_dim=state_dim, action_dim=action_dim, cuda=cuda, model_type=model_type, losses=losses,
                             inverse_model_type=inverse_model_type)

    def SRLModulesSplit(state_dim=2, action_dim=6, cuda=False, model_type="custom_cnn", losses=None, split_dimensions=None,
                        n_hidden_reward=16, inverse_model_type="linear"):
        if _hot(seen, state_dim, cuda, model_type, losses, inverse_model_type) and n_hidden_reward == 16:
            return B200SRLModulesSplit(state_dim, action_dim, cuda, model_type, losses, split_dimensions, n_hidden_reward, inverse_model_type)
        if reference_split_cls is None:
            raise ValueError("configuration outside the B200 hot path and no reference class to fall back to")
        return reference_split_cls(state_dim=state_dim, action_dim=action_dim, cuda=cuda, model_type=model_type, losses=losses,
                                   split_dimensions=split_dimensions, n_hidden_reward=n_hidden_reward, inverse_model_type=inverse_model_type)

    SRLModules.split = SRLModulesSplit
    return SRLModules


def install(learner_module, modules_module=None):
    """Rebind `SRLModules` and the hot-path loss functions inside the reference's `models.learner` (and
    `models.modules` for external importers, models/modules.py:8-14).  Returns the dict of replaced objects."""
    replaced = {}
    ref_cls = getattr(learner_module, "SRLModules", None)
    ref_split = getattr(learner_module, "SRLModulesSplit", None)
    dispatch = _make_dispatch(ref_cls, ref_split)
    replaced["SRLModules"], replaced["SRLModulesSplit"] = ref_cls, ref_split
    learner_module.SRLModules = dispatch
    if ref_split is not None:
        learner_module.SRLModulesSplit = dispatch.split
    if modules_module is not None:
        modules_module.B200SRLModules = B200SRLModules
        modules_module.B200SRLModulesSplit = B200SRLModulesSplit
    for name in _LOSS_NAMES:
        ref_fn = getattr(learner_module, name, None)
        replaced[name] = ref_fn
        setattr(learner_module, name, _route(getattr(_losses, name), ref_fn))
    return replaced


def _route(b200_fn, ref_fn):
    """CUDA float32 tensors of matching shapes -> libsrlz kernels; anything else (CPU plumbing configs, broadcasting or
    mixed-dtype calls) -> the reference implementation."""
    if isinstance(b200_fn, type):  # LossManager: pure host bookkeeping, identical semantics
        return b200_fn

    def fn(*args, **kwargs):
        if ref_fn is not None and not _losses.supports(b200_fn.__name__, list(args) + list(kwargs.values())):
            return ref_fn(*args, **kwargs)
        return b200_fn(*args, **kwargs)

    fn.__name__ = b200_fn.__name__
    fn.__doc__ = b200_fn.__doc__
    return fn
