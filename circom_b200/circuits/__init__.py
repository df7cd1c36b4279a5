"""Circuit library (stand-in for circomlib / circom-ecdsa, which are not in the reference tree)."""
from .basic import multiplier2, is_zero, num2bits, bits2num, less_than, multiplier_n, all_ops, mixed_array, table_lookup, logging  # noqa: F401
from .poseidon import poseidon, poseidon_hash  # noqa: F401
from .sha256 import sha256_compression, sha256  # noqa: F401
from .bigint import big_mult_mod_p, ecdsa_scale, ecdsa_scale_expected, SECP256K1_P  # noqa: F401
from .functions import fn_bit_length, fn_divmod, fn_divmod_array, fn_gcd, gcd_circuit, int_div, int_div_array  # noqa: F401
