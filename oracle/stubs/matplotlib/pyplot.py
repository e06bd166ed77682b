"""Empty stub, see ../gymnasium/__init__.py."""
