"""Test-only shim: the reference logger imports termcolor (utils/logger.py:6)."""


def colored(text, *args, **kwargs):
    return text
