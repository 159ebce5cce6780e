"""File names BASELINE.json's north_star asks for (`models/fusion.py`, `models/encoders.py`); the semantics are the
reference's late-fusion model (notebooks/extended_metrics.py:338-367), not its legacy models/ package."""
