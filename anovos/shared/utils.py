from anovos_b200.shared.utils import attributeType_segregation, get_dtype, ends_with  # noqa: F401
