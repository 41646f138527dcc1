"""Import shim (test infrastructure only) for the two timm symbols the reference imports."""
