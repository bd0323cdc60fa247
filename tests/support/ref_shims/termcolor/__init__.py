"""TEST SHIM: stands in for the absent `termcolor` package (reference logger only)."""


def colored(text, *args, **kwargs):
    return text
