from anovos_b200.data_report.report_preprocessing import save_stats  # noqa: F401
