"""Drop-in replacements for the reference's ``model`` package (inference path)."""
