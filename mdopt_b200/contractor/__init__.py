from mdopt_b200.contractor.contractor import apply_one_site_operator, mps_mpo_contract  # noqa: F401
