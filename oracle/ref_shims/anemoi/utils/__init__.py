"""TEST INFRASTRUCTURE ONLY -- stand-in for the part of `anemoi-utils` the reference's model file imports (not in this image).
`anemoi` is a namespace package, so this directory merges with /root/reference/src/anemoi when both are on sys.path."""
