"""geomjax_b200: B200-native batched-chain engine for geomjax's Riemannian samplers.

Drop-in for the reference's hot path only (geomjax/__init__.py:17-25 names):
``rmhmc``, ``lmc``, ``lmcmonge`` (+ ``rhat``, ``ess``, ``step_size_adaptation``,
``window_adaptation``, ``dual_averaging``), operating on a leading chain axis and executing as
hand-written CUDA for sm_100a behind the C ABI of ``include/geomb200.h``.
"""
from . import integrators, random, targets  # noqa: F401
from .adaptation import dual_averaging, step_size_adaptation, window_adaptation  # noqa: F401
from .base import AdaptationAlgorithm, AdaptationResults, SamplingAlgorithm  # noqa: F401
from .samplers import (DynamicLMCState, DynamicRMHMCState, LMCInfo, LMCState, Proposal, RMHMCInfo,  # noqa: F401
                       RMHMCState, dynamic_lmc, dynamic_lmcmonge, dynamic_rmhmc, lmc, lmcmonge, rmhmc, run_fused)
from .chees import chees_adaptation as chees_adaptation_riemanian  # noqa: F401
from . import chees  # noqa: F401
from .plan import LockstepPlan  # noqa: F401
from .diagnostics import effective_sample_size as ess  # noqa: F401
from .diagnostics import potential_scale_reduction as rhat  # noqa: F401
from .diagnostics import StreamingDiagnostics, sample_streaming  # noqa: F401
from .targets import TargetDescriptor, banana, gaussian, logistic_regression, neal_funnel, softabs  # noqa: F401

__version__ = "0.1.0"
