from anovos_b200.drift_stability.drift_detector import statistics  # noqa: F401
