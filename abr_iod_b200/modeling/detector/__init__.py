from .soften import select_soften_proposals  # noqa: F401
