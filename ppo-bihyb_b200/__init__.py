"""B200-native lower-level GED solver (IPFP + LSAPE/Hungarian projection + edit-cost scoring) for PPO-BiHyb.

Drop-in for the reference's ``utils/ged_env.py`` solver call sites; see DESIGN.md and INTEGRATION.md.
"""
__version__ = "0.1.0"
