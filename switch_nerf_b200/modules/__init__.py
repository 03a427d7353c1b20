"""Import surface of `switch_nerf.modules` for the hot path."""
