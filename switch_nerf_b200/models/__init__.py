"""Import surface of `switch_nerf.models` for the hot path (reference models/nerf_moe.py, models/nerf.py)."""
