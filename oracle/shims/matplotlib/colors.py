class LogNorm:  # ovo/utils/eval_utils.py:3 (plots only)
    pass
