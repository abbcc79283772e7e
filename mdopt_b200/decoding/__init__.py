from mdopt_b200.decoding.decoding import decode_css, decode_css_batch  # noqa: F401
