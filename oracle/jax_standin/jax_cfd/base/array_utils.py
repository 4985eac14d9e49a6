"""Placeholder: the reference's diffusion.py imports the name; nothing on the explicit IB step calls into it."""
