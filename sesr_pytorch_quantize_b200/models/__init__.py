"""Model definitions of the hot path: collapsible SESR x4 / x2 and NRDM_3, float and `_sim` variants
(mirror of the reference's models/ package; same class and attribute names so checkpoints load unchanged)."""
