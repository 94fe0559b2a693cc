from anovos_b200.drift_stability.validations import check_list_of_columns, check_distance_method  # noqa: F401
