"""Import shim (test infrastructure): seaborn is absent in the build container (ovo/utils/eval_utils.py:7, plots only)."""
