from anovos_b200.data_analyzer.quality_checker import nullColumns_detection, IDness_detection, biasedness_detection  # noqa: F401
