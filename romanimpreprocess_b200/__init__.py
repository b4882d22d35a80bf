"""B200-native L1 -> L2 calibration / forward-ramp hot path of romanimpreprocess."""

__version__ = "0.2.0"
