"""TEST INFRASTRUCTURE ONLY.  Import the *unmodified* reference (/root/reference) in this
container so the oracle can be pinned against it and golden vectors can be generated.

The reference's `modules/mage_model.py` imports three packages that are not installed
here (`pytorch_transformers`, `omegaconf`, `ldm`).  None of them is touched
on the VQ sampling path, so three empty shims in `sys.modules` are enough to run
`MAGE.autoregressive_generate` as shipped.

/root/reference does not exist on the GPU box.  `oracle/build_ref.py` (run by `__graft_entry__.build()` in the authoring
container) copies the three reference source files of this path (modules/mage_model.py, modules/vqvae_model.py, utils/util.py), unmodified, into the git-ignored `oracle/_ref/`, which travels
to the GPU box with the repo snapshot: there it serves ONLY bench.py's reference arms (`--impl reference`, `cpu_baseline`,
`eager_gpu_baseline`) -- the thing being compared against, never the thing measured or shipped.  Callers: oracle/make_golden.py,
the `needs_reference` CPU tests, bench.py's reference arms.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    env = os.environ.get("MAGE_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isfile(os.path.join(cand, "modules", "mage_model.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "mage_model.py"))


class _DictConfig(dict):
    """Hashable dict: utils/util.py:53 puts configs in a *set literal* before merging."""
    __hash__ = object.__hash__  # type: ignore[assignment]


def _install_shims() -> None:
    if "pytorch_transformers" not in sys.modules:
        sys.modules["pytorch_transformers"] = types.ModuleType("pytorch_transformers")
    if "omegaconf" not in sys.modules:
        m = types.ModuleType("omegaconf")

        class OmegaConf:  # noqa: D401 - only .merge is used (utils/util.py:53)
            @staticmethod
            def merge(*cfgs):
                out = {}
                # key sets never overlap in the reference's callers (mage_model.py:475-477)
                for c in cfgs:
                    out.update(dict(c))
                return out

        m.OmegaConf = OmegaConf
        m.DictConfig = _DictConfig
        sys.modules["omegaconf"] = m
    if "ldm" not in sys.modules:
        ldm = types.ModuleType("ldm")
        models = types.ModuleType("ldm.models")
        ae = types.ModuleType("ldm.models.autoencoder")

        class DiagonalGaussianDistribution:  # only used in an isinstance test (mage_model.py:543)
            pass

        ae.DiagonalGaussianDistribution = DiagonalGaussianDistribution
        ldm.models = models
        models.autoencoder = ae
        sys.modules["ldm"] = ldm
        sys.modules["ldm.models"] = models
        sys.modules["ldm.models.autoencoder"] = ae


_REF_NAMES = ("modules", "modules.vqvae_model", "modules.mage_model", "utils", "utils.util")
_ref_cache = {}
_ref_cache_ln = {}   # the reference with its documented MAGE+ edit applied in memory (mage_model.py:92-93)

_LINE92 = "x = q + self.dropout(self.attention(q, k, v)) #NOTE"
_LINE93 = "# x = q + self.dropout(self.attention(self.ln_q(q), self.ln_kv(k), self.ln_kv(v), key_mask)) #NOTE"


def _mage_plus_edit(src: str) -> str:
    """The manual edit the reference's own comments prescribe for MAGE+ ("Kindly comment out this line when employing MAGE+" /
    "Please uncomment this line when employing MAGE+", mage_model.py:92-93), applied to the source text IN MEMORY: line 92 is
    commented out, line 93 is uncommented.  Nothing is written anywhere."""
    assert src.count(_LINE92) == 1 and src.count(_LINE93) == 1, "reference source differs from the surveyed revision"
    src = src.replace(_LINE93, _LINE93[2:])
    return src.replace("        " + _LINE92, "        # " + _LINE92)


def _exec_reference_modules(cache=None, mage_plus_edit: bool = False):
    """Execute the reference's source files under their own import names (`modules.*`, `utils.*`).
    This repo has same-named drop-in packages (regular packages beat the reference's namespace
    packages on sys.path), so the files are loaded explicitly by path and the names are swapped in
    only while the reference code runs its imports."""
    import importlib.util

    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ("modules", "utils") or k.startswith("modules.") or k.startswith("utils.")}
    try:
        for pkg in ("modules", "utils"):
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REFERENCE_ROOT, pkg)]
            sys.modules[pkg] = m
        cache = _ref_cache if cache is None else cache
        for name in ("utils.util", "modules.vqvae_model", "modules.mage_model"):
            path = os.path.join(REFERENCE_ROOT, *name.split(".")) + ".py"
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            if mage_plus_edit and name == "modules.mage_model":
                exec(compile(_mage_plus_edit(open(path).read()), path, "exec"), mod.__dict__)
            else:
                spec.loader.exec_module(mod)
        for k in _REF_NAMES:
            cache[k] = sys.modules[k]
    finally:
        for k in list(sys.modules):
            if k in ("modules", "utils") or k.startswith("modules.") or k.startswith("utils."):
                del sys.modules[k]
        sys.modules.update(saved)


class _reference_names:
    """Context manager: `modules.*` / `utils.*` resolve to the reference while active
    (its instantiate_from_config imports targets by dotted path, utils/util.py:58-63)."""

    def __init__(self, cache=None):
        self.cache = _ref_cache if cache is None else cache

    def __enter__(self):
        self.saved = {k: sys.modules.get(k) for k in _REF_NAMES}
        sys.modules.update(self.cache)

    def __exit__(self, *exc):
        for k, v in self.saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load_reference(mage_plus_edit: bool = False):
    """Returns (mage_model module, vqvae_model module) of the real reference; `mage_plus_edit` applies the MAGE+ source edit the
    reference documents at mage_model.py:92-93 (in memory)."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    _install_shims()
    cache = _ref_cache_ln if mage_plus_edit else _ref_cache
    if not cache:
        _exec_reference_modules(cache, mage_plus_edit)
    return cache["modules.mage_model"], cache["modules.vqvae_model"]


def to_dictconfig(obj):
    """Nested plain dict -> nested hashable DictConfig (what the shimmed util expects)."""
    if isinstance(obj, dict):
        return _DictConfig({k: to_dictconfig(v) for k, v in obj.items()})
    return obj


def build_reference_mage(params: dict, state_dict: dict, mage_plus_edit: bool = False):
    """Instantiate the reference's MAGE with `params`, load the synthetic sampling subset
    (strict=False: the train-only conv3d/conv_mu2/conv_var2 keep their default init)."""
    import torch

    mm, _ = load_reference(mage_plus_edit)
    with _reference_names(_ref_cache_ln if mage_plus_edit else _ref_cache):
        torch.manual_seed(0)
        model = mm.MAGE(**to_dictconfig(params))
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    bad = [k for k in missing if not (k.startswith("conv3d.") or k.startswith("conv_mu2") or k.startswith("conv_var2"))]
    assert not bad, f"synthetic checkpoint misses sampling-path keys: {bad}"
    return model.eval()


def build_reference_vqvae(fs_params: dict, state_dict: dict):
    _, vq = load_reference()
    p = {k: v for k, v in fs_params.items() if k != "ckpt_path"}
    model = vq.VectorQuantizedVAE(**p)
    model.load_state_dict(state_dict, strict=True)
    return model.eval()
