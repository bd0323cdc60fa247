"""TEST SHIM: stands in for the absent `yacs` package so the read-only reference
under /root/reference can be imported in the build container."""
