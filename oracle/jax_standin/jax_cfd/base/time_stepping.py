"""Placeholder: the reference's time_stepping.py imports the name and defines its own Runge-Kutta driver."""
