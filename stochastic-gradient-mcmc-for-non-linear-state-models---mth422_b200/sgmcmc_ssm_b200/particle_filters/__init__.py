from .buffered_smoother import buffered_pf_wrapper, pf_wrapper, average_statistic, batched_pf  # noqa: F401
