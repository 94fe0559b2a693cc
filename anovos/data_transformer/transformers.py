from anovos_b200.data_transformer.transformers import attribute_binning  # noqa: F401
