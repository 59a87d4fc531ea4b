"""Import shim (test infrastructure): matplotlib is absent in the build container and only used by the reference's
plotting helpers, never by the functions the oracle is pinned against."""
