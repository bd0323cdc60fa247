"""Dependency stand-in: the reference imports `yacs.config.CfgNode`; yacs is not installable offline, so the reference arm of
bench.py (`--impl reference`, running the UNMODIFIED reference installed under baseline/_ref) gets this package's yacs-compatible
CfgNode instead.  Not used by the product path."""
