from anovos_b200.data_ingest.data_ingest import read_dataset, write_dataset  # noqa: F401
