"""Minimal stand-in for the `typus` package (absent offline): only what linnaeus/inference/postprocessing.py and artifacts.py touch.
Test infrastructure: lets the UNMODIFIED reference functions run in the build container to pin the oracle."""
