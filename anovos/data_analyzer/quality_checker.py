from anovos_b200.data_analyzer.quality_checker import (  # noqa: F401
    nullColumns_detection, outlier_detection, IDness_detection, biasedness_detection)
