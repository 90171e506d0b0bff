"""API protocol types (mirror of geomjax/base.py:88-113,146 and geomjax/types.py)."""
from typing import Callable, NamedTuple


class SamplingAlgorithm(NamedTuple):
    """geomjax/base.py:88-113: ``init(position) -> State``; ``step(rng_key, state) -> (State, Info)``.
    Both operate on a leading chain axis (== ``jax.vmap`` of the reference functions)."""
    init: Callable
    step: Callable


class AdaptationAlgorithm(NamedTuple):
    """geomjax/base.py:146: ``run(rng_key, position, num_steps) -> (AdaptationResults, info)``."""
    run: Callable


class AdaptationResults(NamedTuple):
    """geomjax/adaptation/base.py:22-27."""
    state: NamedTuple
    parameters: dict
