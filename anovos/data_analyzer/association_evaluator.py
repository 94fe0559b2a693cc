from anovos_b200.data_analyzer.association_evaluator import IV_calculation, IG_calculation  # noqa: F401
