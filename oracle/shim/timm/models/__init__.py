"""TEST INFRASTRUCTURE ONLY (see oracle/shim/timm/__init__.py)."""
