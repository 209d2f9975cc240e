"""pairs_b200: a B200 (sm_100a) execution backend for the hot path of the P4IRS/"pairs" particle DSL."""
__version__ = "0.1"
