"""Import-path shims that let unmodified reference modules bind to libdwb (see cauchy_mult.py)."""
import os

PATH = os.path.dirname(os.path.abspath(__file__))
