"""resolve2d_b200 — B200-native (sm_100a) implementation of resolve2d's `Solver.process` hot path.

Python here is only the host-side mirror of the reference's `Solver`/`EntityFactory` API over the C ABI of
`libr2d_b200.so` (include/r2d_abi.h).  All per-step work runs in hand-written CUDA kernels; there is no CPU fallback.
"""
from ._abi import (MODE_FAST, MODE_PARITY, MODE_REFERENCE_ORDER, OPT_SLEEP_CALLS, OPT_SLEEPING, OPT_WARM_START, SHAPE_DISC, SHAPE_RECT, BodyState, R2DError, StepStats, body_desc_dtype,
                   load_library, manifold_dtype)
from .solver import (Batch, BodyHandle, BodyOptions, DiscOptions, EntityFactory, Parameters, RectangleOptions, ShardedBatch,
                     Solver)

__all__ = [
    "Batch", "BodyHandle", "BodyOptions", "BodyState", "DiscOptions", "EntityFactory", "MODE_FAST", "MODE_PARITY", "MODE_REFERENCE_ORDER", "OPT_SLEEPING", "OPT_SLEEP_CALLS", "OPT_WARM_START",
    "Parameters", "R2DError", "RectangleOptions", "SHAPE_DISC", "SHAPE_RECT", "ShardedBatch", "Solver", "StepStats",
    "body_desc_dtype", "load_library", "manifold_dtype",
]
