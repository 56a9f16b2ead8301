"""See __init__.py: importable placeholders; the golden generator never calls them (it replaces
`rotate_sh` before running the reference's GaussianAdapter)."""


def matrix_to_angles(*args, **kwargs):
    raise NotImplementedError("e3nn is not available offline; rotate_sh cannot be run from the reference")


def wigner_D(*args, **kwargs):
    raise NotImplementedError("e3nn is not available offline; rotate_sh cannot be run from the reference")
