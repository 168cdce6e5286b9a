"""Alias package so that the reference's own command lines keep working unchanged
(`python -m src.simulate ...`, `from src.model import TransitionModel`): every name resolves to the
B200-native implementation in conclave_sim_b200."""
