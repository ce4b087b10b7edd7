"""The reference's package name, re-exporting ptina_b200 (see ptina_b200/compat/__init__.py).  Unlike the reference's
ptina/__init__.py this does not register a Blender add-on."""
