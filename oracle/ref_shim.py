"""Import shim for the UNMODIFIED reference (TobyBoyne/bark) hot path.

TEST INFRASTRUCTURE ONLY.  This module exists so that, in the build container
(where `/root/reference` is mounted read-only), the oracle restatement in
`oracle/bark_oracle.py` can be validated against the reference's own numba
functions and golden vectors can be generated (`oracle/make_golden.py`).
Nothing on the product path, in `-m gpu` tests or in `smoke()` imports it.
`bench.py`'s CPU legs (`--impl reference`, `cpu_baseline`) use it to time the
reference's own sampler: on the GPU box `/root/reference` does not exist, so the
shim then resolves to the byte-identical copy staged by `oracle/build_ref.py`
under `baseline/_ref/src` (git-ignored, shipped by gpurun).

What is shimmed (the reference itself is untouched):
  * `bofire.*`  -- imported by `src/bark/fitting/bark_sampler.py:3` and
    `src/bofire_mixed/domain.py:6-13`; absent here -> empty stub modules.
  * `gpytorch`  -- imported by `src/bark/tree_kernels/tree_gps.py:3`; stub.
  * `scipy.special.gammaln` inside njit
    (`src/bark/fitting/noise_scale_proposals.py:26,37`) needs numba-scipy;
    overloaded with `math.lgamma`.
"""
from __future__ import annotations

import math
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "src")


def _pick_src() -> str:
    """The mounted reference (build container) or, on the GPU box, the byte-identical copy of its hot-path modules
    staged by `oracle/build_ref.py` under baseline/_ref/ (git-ignored, shipped by gpurun)."""
    env = os.environ.get("BARK_REFERENCE_SRC")
    if env:
        return env
    if os.path.isdir("/root/reference/src/bark"):
        return "/root/reference/src"
    return _STAGED


REFERENCE_SRC = _pick_src()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "bark"))


def _stub(name: str, **attrs):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


_installed = False


def install():
    """Make `import bark...` resolve to the reference sources."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")

    class _Any:  # placeholder type for annotations / isinstance checks
        pass

    def mk(n):
        return type(n, (_Any,), {})

    _stub("bofire")
    _stub("bofire.data_models")
    _stub("bofire.data_models.domain")
    _stub("bofire.data_models.domain.api", Domain=mk("Domain"), Features=mk("Features"),
          Inputs=mk("Inputs"), Outputs=mk("Outputs"))
    _stub("bofire.data_models.features")
    _stub("bofire.data_models.features.api", AnyFeature=mk("AnyFeature"),
          CategoricalInput=mk("CategoricalInput"), ContinuousInput=mk("ContinuousInput"),
          DiscreteInput=mk("DiscreteInput"))

    gpy = _stub("gpytorch")
    gpy.models = _stub("gpytorch.models", ExactGP=mk("ExactGP"))
    gpy.kernels = _stub("gpytorch.kernels", Kernel=mk("Kernel"), ScaleKernel=mk("ScaleKernel"),
                        IndexKernel=mk("IndexKernel"))
    gpy.means = _stub("gpytorch.means", ZeroMean=mk("ZeroMean"))
    gpy.distributions = _stub("gpytorch.distributions", MultivariateNormal=mk("MultivariateNormal"))
    gpy.constraints = _stub("gpytorch.constraints", Positive=mk("Positive"), Interval=mk("Interval"))
    gpy.likelihoods = _stub("gpytorch.likelihoods", Likelihood=mk("Likelihood"),
                            GaussianLikelihood=mk("GaussianLikelihood"))
    gpy.priors = _stub("gpytorch.priors", Prior=mk("Prior"))

    import numba
    import scipy.special
    from numba.extending import overload

    @overload(scipy.special.gammaln)
    def _gammaln(x):  # noqa: ANN001
        def impl(x):
            return math.lgamma(x)
        return impl

    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    _installed = True


def empty_forest(m: int, node_limit: int = 100):
    """numpy>=2-safe version of `create_empty_forest` (`src/bark/forest.py:114-117`):
    `-1 -> uint32` raises OverflowError on numpy 2, so write 0xFFFFFFFF."""
    install()
    import numpy as np
    from bark.forest import NODE_RECORD_DTYPE
    f = np.zeros((m, node_limit), dtype=NODE_RECORD_DTYPE)
    f[:, 0] = (1, 0, 0, 0, 0, 0xFFFFFFFF, 0, 1)
    return f
