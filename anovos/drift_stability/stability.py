from anovos_b200.drift_stability.stability import stability_index_computation  # noqa: F401
