"""radzero_b200 -- B200-native VL-CABS similarity path behind RadZero's own surface."""
__version__ = "0.1.0"
