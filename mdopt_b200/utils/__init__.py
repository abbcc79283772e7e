from mdopt_b200.utils.utils import split_two_site_tensor, svd, svd_batch  # noqa: F401
