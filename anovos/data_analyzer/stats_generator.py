from anovos_b200.data_analyzer.stats_generator import *  # noqa: F401,F403
from anovos_b200.data_analyzer.stats_generator import (global_summary, missingCount_computation,  # noqa: F401
    nonzeroCount_computation, measures_of_counts, mode_computation, measures_of_centralTendency,
    uniqueCount_computation, measures_of_cardinality, measures_of_dispersion, measures_of_percentiles,
    measures_of_shape)
