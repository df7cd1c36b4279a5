"""Circuit library (stand-in for circomlib / circom-ecdsa, which are not in the reference tree)."""
from .basic import multiplier2, is_zero, num2bits, bits2num, less_than, multiplier_n, all_ops  # noqa: F401
