"""`from src.model import TransitionModel` -> conclave_sim_b200.model."""
from conclave_sim_b200.model import TransitionModel  # noqa: F401
