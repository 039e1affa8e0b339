"""Drop-in for the reference's core/const.py."""
limit = 1e-5
g = -9.80  # Gravity
