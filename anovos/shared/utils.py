from anovos_b200.shared.utils import (  # noqa: F401
    attributeType_segregation, get_dtype, ends_with, flatten_dataframe, transpose_dataframe)
