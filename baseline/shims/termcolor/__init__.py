"""Dependency stand-in for `termcolor` (used by the reference's logger only)."""


def colored(text, *args, **kwargs):
    return text
