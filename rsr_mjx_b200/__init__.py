"""rsr_mjx_b200 — B200-native batched physics-and-environment stepper for the
RSR-MJX Airbot tasks."""
__version__ = "0.1.0"
