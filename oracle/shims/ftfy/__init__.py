"""Shim for the un-vendored `ftfy` dependency (tokenizer.py:14). Identity is exact for ASCII."""


def fix_text(x):
    return x
