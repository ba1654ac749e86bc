"""ORACLE helper (test infrastructure): the seeded synthetic inputs every parity test, golden file and benchmark
uses. Defined once in diffvits_b200.synthetic (the product may not import oracle/, the oracle may import the
product's workload definition) and re-exported here."""
from diffvits_b200.synthetic import betas_linear, make_inputs, make_prompt, ragged_mask  # noqa: F401
