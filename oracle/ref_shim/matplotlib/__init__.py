"""Stub: the reference imports matplotlib.pyplot only for a debug viewer (show_all_obs)."""
