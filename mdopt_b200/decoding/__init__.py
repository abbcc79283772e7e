from mdopt_b200.decoding.decoding import (decode_css, decode_css_batch, decode_custom,  # noqa: F401
                                          decode_custom_batch)
