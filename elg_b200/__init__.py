"""elg_b200 — B200-native hot path of gaocrr/ELG (rollout of the ensemble local+global policy)."""
__version__ = "0.1.0"
