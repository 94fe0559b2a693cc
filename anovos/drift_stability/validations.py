from anovos_b200.drift_stability.validations import (check_list_of_columns, check_distance_method, compute_score,  # noqa: F401
    compute_si, check_metric_weightages, check_threshold)
