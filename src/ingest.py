"""`from src.ingest import ElectorDataIngester` -> conclave_sim_b200.ingest."""
from conclave_sim_b200.ingest import ElectorDataIngester, roster_arrays  # noqa: F401
