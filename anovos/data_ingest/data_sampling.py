from anovos_b200.data_ingest.data_sampling import data_sample  # noqa: F401
