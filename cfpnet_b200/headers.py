"""Constants of include/cfp.h the host side needs (checked against the header by tests/test_host.py)."""
SUMSQ_FLOATS = 1026
SILOG_SCRATCH_DOUBLES = 1544
METRICS_SCRATCH_DOUBLES = 5634
