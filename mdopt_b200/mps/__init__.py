from mdopt_b200.mps.canonical import CanonicalMPS  # noqa: F401
from mdopt_b200.mps.explicit import ExplicitMPS  # noqa: F401
