from .. import shift_invar_ops  # noqa: F401
