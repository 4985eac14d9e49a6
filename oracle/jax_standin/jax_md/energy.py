"""Placeholder: imported by name at the top of the reference's MD files; not restated."""
