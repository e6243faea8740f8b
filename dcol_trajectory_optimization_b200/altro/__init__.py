"""AL-iLQR caller with batched collision constraints + the reference's three scenarios."""
from .problems import PROBLEMS, Problem, cone_through_wall, piano_mover, quadrotor  # noqa: F401
from .solver import AltroResult, EngineEvaluator, altro_solve  # noqa: F401
