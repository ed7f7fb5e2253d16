"""Measured baselines that stand beside the engine's numbers (bench.py) -- not part of the product package."""
