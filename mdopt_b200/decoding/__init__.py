from mdopt_b200.decoding.decoding import (css_code_stabilisers, decode_css, decode_css_batch,  # noqa: F401
                                          decode_custom, decode_custom_batch, multiply_pauli_strings)
